// Minimal stand-in for the parts of Boost.Geometry that include/gridpp.h and src/api/kdtree.cpp use.
//
// TEST INFRASTRUCTURE ONLY. Boost is not installed in this image, so the reference's hot-path sources are
// compiled (unmodified, from /root/reference/src/api) against this header to obtain oracle/_ref/. Nothing in
// the product (gridpp_b200/) includes this file.
//
// Surface provided (all that gridpp.h:1846-1849 and kdtree.cpp:6-16,39-106 touch):
//   boost::geometry::model::point<float,3,cs::cartesian>   get<I>(), set<I>()
//   boost::geometry::model::box<point>
//   boost::geometry::index::rtree<value, quadratic<16>>    insert(), size(), query(pred, out)
//   predicates: within(box) [STRICT interior, as Boost's within()], satisfies(f), nearest(p, k), operator&&
//
// The spatial index behind rtree<> is a uniform-cell bucket grid (not brute force) so that the CPU baseline
// timed through oracle/_ref is not penalised by the shim. Result ORDER of a within() query is ascending
// insertion index; for <=16 values (one Boost leaf) that is also Boost's order, which is what the
// reference's own tests pin (tests/test_kdtree.py:12-13). nearest(p,k) ties resolve to the lowest index and
// results come back sorted by (distance, index); Boost leaves ties unspecified (tests/test_points.py:99).
#ifndef ORACLE_SHIM_BOOST_GEOMETRY_HPP
#define ORACLE_SHIM_BOOST_GEOMETRY_HPP

// The real Boost headers pull these in transitively and the reference relies on that. <math.h> and
// <stdlib.h> matter for NUMERICS: Boost (boost/math/special_functions/fpclassify.hpp, reached from
// boost/math/distributions/gamma.hpp in gridpp.h:9) includes the C++ wrapper <math.h>, which puts the
// float overloads of log/exp/sqrt/abs into the global namespace. structure.cpp calls them unqualified on
// float arguments (structure.cpp:79,281,429-430), so with Boost they evaluate in float; without the wrapper
// they would silently evaluate in double (and abs(float) would truncate to int).
#include <math.h>
#include <stdlib.h>
#include <cassert>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

namespace boost { namespace geometry {

namespace cs { struct cartesian {}; }

namespace model {
template <class T, std::size_t D, class CS>
class point {
public:
    point() { for(std::size_t i = 0; i < D; i++) m_v[i] = T(); }
    point(T a, T b = T(), T c = T()) {
        T tmp[3] = {a, b, c};
        for(std::size_t i = 0; i < D && i < 3; i++) m_v[i] = tmp[i];
    }
    template <std::size_t I> T get() const { return m_v[I]; }
    template <std::size_t I> void set(T value) { m_v[I] = value; }
private:
    T m_v[D];
};

template <class P>
class box {
public:
    box() {}
    box(P const& lo, P const& hi) : m_lo(lo), m_hi(hi) {}
    P const& min_corner() const { return m_lo; }
    P const& max_corner() const { return m_hi; }
private:
    P m_lo, m_hi;
};
}  // namespace model

namespace index {

template <std::size_t Max, std::size_t Min = 0> struct quadratic {};

namespace detail {
template <class Box> struct within_pred { Box b; };
template <class F> struct satisfies_pred { F f; };
template <class P> struct nearest_pred { P p; unsigned k; };
template <class A, class B> struct and_pred { A a; B b; };

template <class Box, class F>
and_pred<within_pred<Box>, satisfies_pred<F> > operator&&(within_pred<Box> const& a, satisfies_pred<F> const& b) {
    and_pred<within_pred<Box>, satisfies_pred<F> > r = {a, b};
    return r;
}
template <class P, class F>
and_pred<nearest_pred<P>, satisfies_pred<F> > operator&&(nearest_pred<P> const& a, satisfies_pred<F> const& b) {
    and_pred<nearest_pred<P>, satisfies_pred<F> > r = {a, b};
    return r;
}
struct always_true { template <class V> bool operator()(V const&) const { return true; } };
inline std::mutex& build_mutex() { static std::mutex m; return m; }
}  // namespace detail

template <class Box> detail::within_pred<Box> within(Box const& b) { detail::within_pred<Box> r = {b}; return r; }
template <class F> detail::satisfies_pred<F> satisfies(F const& f) { detail::satisfies_pred<F> r = {f}; return r; }
template <class P> detail::nearest_pred<P> nearest(P const& p, unsigned k) { detail::nearest_pred<P> r = {p, k}; return r; }

// Value is std::pair<point<float,3>, unsigned>
template <class Value, class Params>
class rtree {
    typedef typename Value::first_type point_t;
    struct cells_t {
        float lo[3];
        double inv[3];     // cells per unit length
        double size[3];    // cell edge
        int n[3];
        std::vector<int> start;   // n0*n1*n2 + 1 offsets into order
        std::vector<int> order;   // value indices sorted by cell, insertion order inside a cell
        int cell_of(int d, float c) const {
            int i = (int) std::floor((double(c) - double(lo[d])) * inv[d]);
            if(i < 0) i = 0;
            if(i > n[d] - 1) i = n[d] - 1;
            return i;
        }
    };
public:
    rtree() {}
    rtree(rtree const& other) : m_values(other.m_values), m_cells(std::atomic_load(&other.m_cells)) {}
    rtree& operator=(rtree const& other) {
        if(this != &other) {
            m_values = other.m_values;
            std::atomic_store(&m_cells, std::atomic_load(&other.m_cells));
        }
        return *this;
    }
    void insert(Value const& v) {
        m_values.push_back(v);
        std::atomic_store(&m_cells, std::shared_ptr<const cells_t>());
    }
    std::size_t size() const { return m_values.size(); }
    bool empty() const { return m_values.empty(); }

    template <class Box, class F, class Out>
    std::size_t query(detail::and_pred<detail::within_pred<Box>, detail::satisfies_pred<F> > const& p, Out out) const {
        return query_within(p.a.b, p.b.f, out);
    }
    template <class Box, class Out>
    std::size_t query(detail::within_pred<Box> const& p, Out out) const {
        return query_within(p.b, detail::always_true(), out);
    }
    template <class P, class F, class Out>
    std::size_t query(detail::and_pred<detail::nearest_pred<P>, detail::satisfies_pred<F> > const& p, Out out) const {
        return query_nearest(p.a.p, p.a.k, p.b.f, out);
    }
    template <class P, class Out>
    std::size_t query(detail::nearest_pred<P> const& p, Out out) const {
        return query_nearest(p.p, p.k, detail::always_true(), out);
    }

private:
    static float coord(point_t const& p, int d) {
        return d == 0 ? p.template get<0>() : (d == 1 ? p.template get<1>() : p.template get<2>());
    }
    std::shared_ptr<const cells_t> cells() const {
        std::shared_ptr<const cells_t> c = std::atomic_load(&m_cells);
        if(!c) {
            std::lock_guard<std::mutex> lock(detail::build_mutex());
            c = std::atomic_load(&m_cells);
            if(!c) {
                c = build();
                std::atomic_store(&m_cells, c);
            }
        }
        return c;
    }
    std::shared_ptr<const cells_t> build() const {
        std::shared_ptr<cells_t> c(new cells_t);
        const int N = (int) m_values.size();
        float hi[3];
        for(int d = 0; d < 3; d++) {
            c->lo[d] = std::numeric_limits<float>::infinity();
            hi[d] = -std::numeric_limits<float>::infinity();
        }
        for(int i = 0; i < N; i++)
            for(int d = 0; d < 3; d++) {
                float v = coord(m_values[i].first, d);
                c->lo[d] = std::min(c->lo[d], v);
                hi[d] = std::max(hi[d], v);
            }
        double ext[3];
        int nd = 0;
        double vol = 1;
        for(int d = 0; d < 3; d++) {
            ext[d] = N > 0 ? double(hi[d]) - double(c->lo[d]) : 0;
            if(ext[d] > 0) { nd++; vol *= ext[d]; }
        }
        // about 2 values per occupied cell; a single cell for tiny sets (keeps insertion order)
        double edge = 0;
        if(N > 16 && nd > 0) edge = std::pow(vol / (N / 2.0), 1.0 / nd);
        long long total = 1;
        for(int d = 0; d < 3; d++) {
            int n = 1;
            if(edge > 0 && ext[d] > 0) n = (int) std::min(1024.0, std::max(1.0, std::ceil(ext[d] / edge)));
            c->n[d] = n;
            total *= n;
        }
        while(total > (1LL << 23)) {   // cap memory: halve the largest dimension
            int dmax = 0;
            for(int d = 1; d < 3; d++) if(c->n[d] > c->n[dmax]) dmax = d;
            total /= c->n[dmax];
            c->n[dmax] = (c->n[dmax] + 1) / 2;
            total *= c->n[dmax];
        }
        for(int d = 0; d < 3; d++) {
            c->size[d] = ext[d] > 0 ? ext[d] / c->n[d] : 1.0;
            c->inv[d] = ext[d] > 0 ? c->n[d] / ext[d] : 0.0;
        }
        c->start.assign((std::size_t) total + 1, 0);
        std::vector<int> cell(N);
        for(int i = 0; i < N; i++) {
            int id = (c->cell_of(2, coord(m_values[i].first, 2)) * c->n[1] + c->cell_of(1, coord(m_values[i].first, 1))) * c->n[0]
                     + c->cell_of(0, coord(m_values[i].first, 0));
            cell[i] = id;
            c->start[id + 1]++;
        }
        for(std::size_t i = 0; i < (std::size_t) total; i++) c->start[i + 1] += c->start[i];
        c->order.resize(N);
        std::vector<int> fill(c->start.begin(), c->start.end() - 1);
        for(int i = 0; i < N; i++) c->order[fill[cell[i]]++] = i;
        return c;
    }

    template <class Box, class F, class Out>
    std::size_t query_within(Box const& b, F const& f, Out out) const {
        if(m_values.empty()) return 0;
        std::shared_ptr<const cells_t> cp = cells();
        const cells_t& c = *cp;
        float lo[3], hi[3];
        int c0[3], c1[3];
        for(int d = 0; d < 3; d++) {
            lo[d] = coord(b.min_corner(), d);
            hi[d] = coord(b.max_corner(), d);
            if(!(lo[d] < hi[d])) return 0;   // empty interior (also rejects NaN)
            c0[d] = c.cell_of(d, lo[d]);
            c1[d] = c.cell_of(d, hi[d]);
        }
        // Matches are reported in ascending insertion index. Boost's own order is its tree-traversal order,
        // which nothing in the reference specifies; a fixed order makes the one order-dependent statement of the
        // hot path (the linear index lY[e] in oi_ensi.cpp:523-524) reproducible.
        std::vector<int> hits;
        for(int cz = c0[2]; cz <= c1[2]; cz++)
            for(int cy = c0[1]; cy <= c1[1]; cy++) {
                int base = (cz * c.n[1] + cy) * c.n[0];
                int s = c.start[base + c0[0]], e = c.start[base + c1[0] + 1];
                for(int k = s; k < e; k++) {
                    Value const& v = m_values[c.order[k]];
                    float x = v.first.template get<0>(), y = v.first.template get<1>(), z = v.first.template get<2>();
                    // Boost's within(): strictly inside the box
                    if(x > lo[0] && x < hi[0] && y > lo[1] && y < hi[1] && z > lo[2] && z < hi[2]) {
                        if(f(v)) hits.push_back(c.order[k]);
                    }
                }
            }
        std::sort(hits.begin(), hits.end());
        for(std::size_t i = 0; i < hits.size(); i++) *out++ = m_values[hits[i]];
        return hits.size();
    }

    template <class P, class F, class Out>
    std::size_t query_nearest(P const& p, unsigned k, F const& f, Out out) const {
        if(m_values.empty() || k == 0) return 0;
        std::shared_ptr<const cells_t> cp = cells();
        const cells_t& c = *cp;
        double q[3];
        int cc[3];
        for(int d = 0; d < 3; d++) {
            q[d] = coord(p, d);
            cc[d] = c.cell_of(d, coord(p, d));
        }
        typedef std::pair<double, unsigned> cand;   // (squared distance, insertion index)
        std::vector<cand> best;                      // max-heap on (dist, index), size <= k
        int maxr = std::max(c.n[0], std::max(c.n[1], c.n[2]));
        for(int r = 0; r <= maxr; r++) {
            int b0[3], b1[3];
            for(int d = 0; d < 3; d++) {
                b0[d] = std::max(0, cc[d] - r);
                b1[d] = std::min(c.n[d] - 1, cc[d] + r);
            }
            for(int cz = b0[2]; cz <= b1[2]; cz++)
                for(int cy = b0[1]; cy <= b1[1]; cy++) {
                    // Only the shell at Chebyshev distance r from the centre cell: a full x-run when this
                    // (z, y) row is itself on the shell, otherwise just the two end cells.
                    bool row_on_shell = std::abs(cz - cc[2]) == r || std::abs(cy - cc[1]) == r;
                    int xs[2] = {cc[0] - r, cc[0] + r};
                    int nx = row_on_shell ? (b1[0] - b0[0] + 1) : (r == 0 ? 1 : 2);
                    for(int ix = 0; ix < nx; ix++) {
                        int cx = row_on_shell ? b0[0] + ix : xs[ix];
                        if(cx < 0 || cx > c.n[0] - 1) continue;
                        int id = (cz * c.n[1] + cy) * c.n[0] + cx;
                        for(int kk = c.start[id]; kk < c.start[id + 1]; kk++) {
                            unsigned idx = (unsigned) c.order[kk];
                            Value const& v = m_values[idx];
                            if(!f(v)) continue;
                            double ddx = q[0] - double(v.first.template get<0>());
                            double ddy = q[1] - double(v.first.template get<1>());
                            double ddz = q[2] - double(v.first.template get<2>());
                            cand cd(ddx * ddx + ddy * ddy + ddz * ddz, idx);
                            if(best.size() < k) {
                                best.push_back(cd);
                                std::push_heap(best.begin(), best.end());
                            }
                            else if(cd < best.front()) {
                                std::pop_heap(best.begin(), best.end());
                                best.back() = cd;
                                std::push_heap(best.begin(), best.end());
                            }
                        }
                    }
                }
            // All cells visited?
            bool all = true;
            for(int d = 0; d < 3; d++) if(b0[d] > 0 || b1[d] < c.n[d] - 1) all = false;
            if(all) break;
            if(best.size() == k) {
                // smallest possible distance to anything outside the visited block
                double bound = std::numeric_limits<double>::infinity();
                for(int d = 0; d < 3; d++) {
                    if(b0[d] > 0) bound = std::min(bound, q[d] - (double(c.lo[d]) + b0[d] * c.size[d]));
                    if(b1[d] < c.n[d] - 1) bound = std::min(bound, (double(c.lo[d]) + (b1[d] + 1) * c.size[d]) - q[d]);
                }
                // one cell of slack absorbs rounding in the cell assignment
                bound -= 1e-3 * std::max(c.size[0], std::max(c.size[1], c.size[2]));
                if(bound > 0 && best.front().first < bound * bound) break;
            }
        }
        std::sort(best.begin(), best.end());
        for(std::size_t i = 0; i < best.size(); i++) *out++ = m_values[best[i].second];
        return best.size();
    }

    std::vector<Value> m_values;
    mutable std::shared_ptr<const cells_t> m_cells;
};

}  // namespace index
}}  // namespace boost::geometry
#endif
