// Shim for boost::math::gamma_distribution<> -- only the type is needed (a data member of gridpp::Gamma,
// include/gridpp.h:2453); transform.cpp, which calls cdf/quantile, is NOT part of the hot-path build.
#ifndef ORACLE_SHIM_BOOST_MATH_GAMMA_HPP
#define ORACLE_SHIM_BOOST_MATH_GAMMA_HPP
namespace boost { namespace math {
template <class T = double> class gamma_distribution {
public:
    gamma_distribution(T shape = 1, T scale = 1) : m_shape(shape), m_scale(scale) {}
    T shape() const { return m_shape; }
    T scale() const { return m_scale; }
private:
    T m_shape, m_scale;
};
}}
#endif
