// Shim for boost::math::normal -- only the type is needed (include/gridpp.h:2454).
#ifndef ORACLE_SHIM_BOOST_MATH_NORMAL_HPP
#define ORACLE_SHIM_BOOST_MATH_NORMAL_HPP
namespace boost { namespace math {
template <class T = double> class normal_distribution {
public:
    normal_distribution(T mean = 0, T sd = 1) : m_mean(mean), m_sd(sd) {}
    T mean() const { return m_mean; }
    T standard_deviation() const { return m_sd; }
private:
    T m_mean, m_sd;
};
typedef normal_distribution<double> normal;
}}
#endif
