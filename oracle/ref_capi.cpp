// Flat C entry points over the UNMODIFIED reference (metno/gridpp src/api/*.cpp compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libgridpp_ref.so).
//
// TEST INFRASTRUCTURE ONLY: used by tests/ (to pin oracle/gridpp_oracle.c and to check the CUDA path), by
// tests/golden/make_golden.py (to generate the committed fixtures) and by bench.py's cpu_baseline /
// --impl reference leg. The product never links or loads this.
//
// Every function returns 0 on success, 1 for std::invalid_argument, 2 for any other exception; the message is
// available from ref_last_error(). Functions that time the reference call (only the gridpp:: call itself, not
// the construction of Points/Grid or the vector<vector<>> marshalling) return the seconds in *seconds.
#include "gridpp.h"
#include "gridpp_b200.h"

#include <chrono>
#include <cstring>
#include <memory>
#include <string>

namespace {
thread_local std::string g_error;

typedef std::chrono::steady_clock clk;
double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

gridpp::vec to_vec(const float* p, int n) { return p ? gridpp::vec(p, p + n) : gridpp::vec(); }
gridpp::vec2 to_vec2(const float* p, int ny, int nx) {
    gridpp::vec2 v(ny);
    for(int y = 0; y < ny; y++) v[y].assign(p + (size_t) y * nx, p + (size_t) (y + 1) * nx);
    return v;
}
void from_vec2(const gridpp::vec2& v, float* out, int ny, int nx) {
    for(int y = 0; y < ny; y++) std::memcpy(out + (size_t) y * nx, v[y].data(), sizeof(float) * nx);
}
gridpp::Points make_points(const float* lats, const float* lons, const float* elevs, const float* lafs, int n, int type) {
    return gridpp::Points(to_vec(lats, n), to_vec(lons, n), to_vec(elevs, n), to_vec(lafs, n), (gridpp::CoordinateType) type);
}

gridpp::StructureFunctionPtr make_term(const gpp_structure_term& t) {
    // hmax is encoded through min_rho: the spatial constructors take min_rho directly and collapse to the
    // non-spatial case for 1x1 scale arrays (structure.cpp:168-184), which is also what clone() does.
    gridpp::vec2 h(1, gridpp::vec(1, t.h)), v(1, gridpp::vec(1, t.v)), w(1, gridpp::vec(1, t.w));
    gridpp::Grid g;
    switch(t.type) {
        case GPP_STRUCT_BARNES: return std::make_shared<gridpp::BarnesStructure>(g, h, v, w, t.min_rho);
        case GPP_STRUCT_CRESSMAN: return std::make_shared<gridpp::CressmanStructure>(t.h, t.v, t.w);
        case GPP_STRUCT_SOAR: return std::make_shared<gridpp::SoarStructure>(g, h, v, w, t.min_rho);
        case GPP_STRUCT_TOAR: return std::make_shared<gridpp::ToarStructure>(g, h, v, w, t.min_rho);
        case GPP_STRUCT_POWERLAW: return std::make_shared<gridpp::PowerlawStructure>(g, h, v, w, t.min_rho);
        case GPP_STRUCT_LINEAR: return std::make_shared<gridpp::LinearStructure>(g, h, v, w, t.min_rho);
    }
    throw std::invalid_argument("unknown structure type");
}
gridpp::StructureFunctionPtr make_structure(const gpp_structure* s) {
    gridpp::StructureFunctionPtr base;
    if(s->n_terms == 3) {
        gridpp::StructureFunctionPtr a = make_term(s->term[0]), b = make_term(s->term[1]), c = make_term(s->term[2]);
        base = std::make_shared<gridpp::MultipleStructure>(*a, *b, *c);
    }
    else
        base = make_term(s->term[0]);
    if(s->has_cv) return std::make_shared<gridpp::CrossValidation>(*base, s->cv_dist);
    return base;
}
}  // namespace

#define REF_TRY try {
#define REF_CATCH                                                        \
    }                                                                    \
    catch(const std::invalid_argument& e) { g_error = e.what(); return 1; } \
    catch(const std::exception& e) { g_error = e.what(); return 2; }     \
    catch(...) { g_error = "unknown exception"; return 2; }              \
    return 0;

extern "C" {

const char* ref_last_error() { return g_error.c_str(); }
const char* ref_version() { static std::string v = gridpp::version(); return v.c_str(); }
void ref_set_omp_threads(int n) { gridpp::set_omp_threads(n); }
int ref_get_omp_threads() { return gridpp::get_omp_threads(); }

// gridpp::convert_coordinates, util.cpp:583-615
int ref_convert_coordinates(const float* lats, const float* lons, int n, int type, float* x, float* y, float* z) {
    REF_TRY
    for(int i = 0; i < n; i++) gridpp::convert_coordinates(lats[i], lons[i], (gridpp::CoordinateType) type, x[i], y[i], z[i]);
    REF_CATCH
}

// Structure function constructors with hmax (structure.cpp:143-167 etc.): fills min_rho/loc_dist of a
// descriptor from the reference object, so the product's host-side formulas can be checked against it.
int ref_structure_describe(int type, float h, float v, float w, float hmax, float* loc_dist) {
    REF_TRY
    gridpp::StructureFunctionPtr s;
    switch(type) {
        case GPP_STRUCT_BARNES: s = std::make_shared<gridpp::BarnesStructure>(h, v, w, hmax); break;
        case GPP_STRUCT_CRESSMAN: s = std::make_shared<gridpp::CressmanStructure>(h, v, w); break;
        case GPP_STRUCT_SOAR: s = std::make_shared<gridpp::SoarStructure>(h, v, w, hmax); break;
        case GPP_STRUCT_TOAR: s = std::make_shared<gridpp::ToarStructure>(h, v, w, hmax); break;
        case GPP_STRUCT_POWERLAW: s = std::make_shared<gridpp::PowerlawStructure>(h, v, w, hmax); break;
        case GPP_STRUCT_LINEAR: s = std::make_shared<gridpp::LinearStructure>(h, v, w, hmax); break;
        default: throw std::invalid_argument("unknown structure type");
    }
    *loc_dist = s->localization_distance(gridpp::Point(0, 0, 0, 0, gridpp::Cartesian));
    REF_CATCH
}

// StructureFunction::corr / corr_background for n point pairs given as (x,y,z,elev,laf)
int ref_structure_corr(const gpp_structure* sd, const float* p1, const float* p2, int n, int background, float* out) {
    REF_TRY
    gridpp::StructureFunctionPtr s = make_structure(sd);
    for(int i = 0; i < n; i++) {
        const float* a = p1 + 5 * i;
        const float* b = p2 + 5 * i;
        gridpp::Point q1(a[1], a[0], a[3], a[4], gridpp::Cartesian, a[0], a[1], a[2]);
        gridpp::Point q2(b[1], b[0], b[3], b[4], gridpp::Cartesian, b[0], b[1], b[2]);
        out[i] = background ? s->corr_background(q1, q2) : s->corr(q1, q2);
    }
    REF_CATCH
}
int ref_structure_localization_distance(const gpp_structure* sd, float* out) {
    REF_TRY
    *out = make_structure(sd)->localization_distance(gridpp::Point(0, 0, 0, 0, gridpp::Cartesian));
    REF_CATCH
}

// ---- KDTree queries (kdtree.cpp:18-106) --------------------------------------------------------------
int ref_points_nearest(const float* lats, const float* lons, int n, int type, const float* qlats, const float* qlons,
                       int nq, int include_match, int* out_index, double* seconds) {
    REF_TRY
    gridpp::Points p = make_points(lats, lons, NULL, NULL, n, type);
    if(n > 0) p.get_nearest_neighbour(qlats[0], qlons[0]);   // build the index outside the timed region
    clk::time_point t0 = clk::now();
    #pragma omp parallel for
    for(int q = 0; q < nq; q++) out_index[q] = p.get_nearest_neighbour(qlats[q], qlons[q], include_match);
    if(seconds) *seconds = since(t0);
    REF_CATCH
}
int ref_points_neighbours(const float* lats, const float* lons, int n, int type, const float* qlats, const float* qlons,
                          const float* radii, int nq, int include_match, int capacity, int* out_index, float* out_dist,
                          int* out_count) {
    REF_TRY
    gridpp::Points p = make_points(lats, lons, NULL, NULL, n, type);
    for(int q = 0; q < nq; q++) {
        gridpp::vec dist;
        gridpp::ivec I = p.get_neighbours_with_distance(qlats[q], qlons[q], radii[q], dist, include_match);
        out_count[q] = I.size();
        // canonical order for comparison: ascending index
        std::vector<std::pair<int, float> > pairs(I.size());
        for(size_t i = 0; i < I.size(); i++) pairs[i] = std::make_pair(I[i], dist[i]);
        std::sort(pairs.begin(), pairs.end());
        for(int i = 0; i < (int) pairs.size() && i < capacity; i++) {
            if(out_index) out_index[(size_t) q * capacity + i] = pairs[i].first;
            if(out_dist) out_dist[(size_t) q * capacity + i] = pairs[i].second;
        }
    }
    REF_CATCH
}
// raw (unsorted) order of a single query, to check the small-tree insertion-order contract
int ref_points_neighbours_raw(const float* lats, const float* lons, int n, int type, float qlat, float qlon, float radius,
                              int include_match, int capacity, int* out_index, int* out_count) {
    REF_TRY
    gridpp::Points p = make_points(lats, lons, NULL, NULL, n, type);
    gridpp::ivec I = p.get_neighbours(qlat, qlon, radius, include_match);
    *out_count = I.size();
    for(int i = 0; i < (int) I.size() && i < capacity; i++) out_index[i] = I[i];
    REF_CATCH
}
int ref_points_closest(const float* lats, const float* lons, int n, int type, const float* qlats, const float* qlons, int nq,
                       int num, int include_match, int* out_index) {
    REF_TRY
    gridpp::Points p = make_points(lats, lons, NULL, NULL, n, type);
    for(int q = 0; q < nq; q++) {
        gridpp::ivec I = p.get_closest_neighbours(qlats[q], qlons[q], num, include_match);
        for(int i = 0; i < num; i++) out_index[(size_t) q * num + i] = i < (int) I.size() ? I[i] : -1;
    }
    REF_CATCH
}
float ref_calc_distance(float lat1, float lon1, float lat2, float lon2, int type) {
    return gridpp::KDTree::calc_distance(lat1, lon1, lat2, lon2, (gridpp::CoordinateType) type);
}

// gridpp::nearest(Grid, Points, vec2) nearest.cpp:124-144 and (Points, Points, vec) :177-197 share the per-point
// lookup; the Grid index arithmetic (grid.cpp:108-114) is flat = y*nx + x, so a flattened field is equivalent.
int ref_nearest(const float* ilats, const float* ilons, int n_in, int type, const float* qlats, const float* qlons, int nq,
                const float* ivalues, int n_fields, float* out, double* seconds) {
    REF_TRY
    gridpp::Points ip = make_points(ilats, ilons, NULL, NULL, n_in, type);
    gridpp::Points op = make_points(qlats, qlons, NULL, NULL, nq, type);
    if(n_in > 0 && nq > 0) ip.get_nearest_neighbour(qlats[0], qlons[0]);
    double total = 0;
    if(n_fields == 1) {
        gridpp::vec iv = to_vec(ivalues, n_in);
        clk::time_point t0 = clk::now();
        gridpp::vec o = gridpp::nearest(ip, op, iv);
        total = since(t0);
        std::memcpy(out, o.data(), sizeof(float) * nq);
    }
    else {
        gridpp::vec2 iv = to_vec2(ivalues, n_fields, n_in);
        clk::time_point t0 = clk::now();
        gridpp::vec2 o = gridpp::nearest(ip, op, iv);
        total = since(t0);
        from_vec2(o, out, n_fields, nq);
    }
    if(seconds) *seconds = total;
    REF_CATCH
}

// ---- optimal interpolation ----------------------------------------------------------------------------
// gridpp::optimal_interpolation_full(Points...) oi.cpp:138-341. bvariance / bvariance_at_points NULL = 1.
int ref_optimal_interpolation(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB,
                              const float* background, const float* bvariance, const float* plats, const float* plons,
                              const float* pelevs, const float* plafs, int nS, int type, const float* pobs,
                              const float* obs_variance, const float* pbackground, const float* bvariance_at_points,
                              const gpp_structure* sd, int max_points, int allow_extrapolation, float* analysis,
                              float* analysis_variance, double* seconds) {
    REF_TRY
    gridpp::Points bp = make_points(blats, blons, belevs, blafs, nB, type);
    gridpp::Points op = make_points(plats, plons, pelevs, plafs, nS, type);
    gridpp::StructureFunctionPtr s = make_structure(sd);
    gridpp::vec bg = to_vec(background, nB), bvar = bvariance ? to_vec(bvariance, nB) : gridpp::vec(nB, 1.0f);
    gridpp::vec obs = to_vec(pobs, nS), ovar = to_vec(obs_variance, nS), pbg = to_vec(pbackground, nS);
    gridpp::vec pbvar = bvariance_at_points ? to_vec(bvariance_at_points, nS) : gridpp::vec(nS, 1.0f);
    if(nS > 0) op.get_neighbours(plats[0], plons[0], 1.0f);   // build the observation index outside the timed region
    gridpp::vec avar;
    clk::time_point t0 = clk::now();
    gridpp::vec out = gridpp::optimal_interpolation_full(bp, bg, bvar, op, obs, ovar, pbg, pbvar, *s, max_points, avar,
                                                         allow_extrapolation != 0);
    if(seconds) *seconds = since(t0);
    std::memcpy(analysis, out.data(), sizeof(float) * out.size());
    if(analysis_variance && avar.size() == (size_t) nB) std::memcpy(analysis_variance, avar.data(), sizeof(float) * nB);
    else if(analysis_variance) std::memcpy(analysis_variance, bvar.data(), sizeof(float) * nB);
    REF_CATCH
}

// gridpp::optimal_interpolation(Grid...) oi.cpp:26-87: the overload a user of the README example calls, with its
// to_points() copy of the grid and the two full-field vec2 <-> vec copies inside the timed region (the Grid itself is
// built outside it). Used by bench.py --impl reference for one full pass over the 4000 x 4000 grid of config 3.
int ref_optimal_interpolation_grid(const float* blats, const float* blons, int ny, int nx, const float* background,
                                   const float* plats, const float* plons, int nS, int type, const float* pobs,
                                   const float* pratios, const float* pbackground, const gpp_structure* sd, int max_points,
                                   int allow_extrapolation, float* analysis, double* seconds) {
    REF_TRY
    gridpp::Grid grid(to_vec2(blats, ny, nx), to_vec2(blons, ny, nx), gridpp::vec2(), gridpp::vec2(), (gridpp::CoordinateType) type);
    gridpp::Points op = make_points(plats, plons, nullptr, nullptr, nS, type);
    gridpp::StructureFunctionPtr s = make_structure(sd);
    gridpp::vec2 bg = to_vec2(background, ny, nx);
    gridpp::vec obs = to_vec(pobs, nS), ratios = to_vec(pratios, nS), pbg = to_vec(pbackground, nS);
    if(nS > 0) op.get_neighbours(plats[0], plons[0], 1.0f);   // build the observation index outside the timed region
    clk::time_point t0 = clk::now();
    gridpp::vec2 out = gridpp::optimal_interpolation(grid, bg, op, obs, ratios, pbg, *s, max_points, allow_extrapolation != 0);
    if(seconds) *seconds = since(t0);
    from_vec2(out, analysis, ny, nx);
    REF_CATCH
}

// gridpp::optimal_interpolation_full(Points...) with a SPATIALLY VARYING structure function: <Family>Structure(Grid, vec2 h,
// vec2 v, vec2 w, min_rho), structure.cpp:168-184 (Barnes), :342 (Soar), :492 (Toar), :643 (Powerlaw), :790 (Linear).
// The scale grid is gny x gnx (glats/glons/h/v/w row-major).
int ref_optimal_interpolation_spatial(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB,
                                      const float* background, const float* plats, const float* plons, const float* pelevs,
                                      const float* plafs, int nS, int type, const float* pobs, const float* obs_variance,
                                      const float* pbackground, int stype, const float* glats, const float* glons, int gny, int gnx,
                                      const float* h, const float* v, const float* w, float min_rho, int max_points,
                                      int allow_extrapolation, float* analysis, float* analysis_variance) {
    REF_TRY
    gridpp::Points bp = make_points(blats, blons, belevs, blafs, nB, type);
    gridpp::Points op = make_points(plats, plons, pelevs, plafs, nS, type);
    gridpp::Grid sg(to_vec2(glats, gny, gnx), to_vec2(glons, gny, gnx), gridpp::vec2(), gridpp::vec2(), (gridpp::CoordinateType) type);
    gridpp::vec2 h2 = to_vec2(h, gny, gnx), v2 = to_vec2(v, gny, gnx), w2 = to_vec2(w, gny, gnx);
    gridpp::StructureFunctionPtr s;
    switch(stype) {
        case GPP_STRUCT_BARNES: s = std::make_shared<gridpp::BarnesStructure>(sg, h2, v2, w2, min_rho); break;
        case GPP_STRUCT_SOAR: s = std::make_shared<gridpp::SoarStructure>(sg, h2, v2, w2, min_rho); break;
        case GPP_STRUCT_TOAR: s = std::make_shared<gridpp::ToarStructure>(sg, h2, v2, w2, min_rho); break;
        case GPP_STRUCT_POWERLAW: s = std::make_shared<gridpp::PowerlawStructure>(sg, h2, v2, w2, min_rho); break;
        case GPP_STRUCT_LINEAR: s = std::make_shared<gridpp::LinearStructure>(sg, h2, v2, w2, min_rho); break;
        default: throw std::invalid_argument("unknown structure type");
    }
    gridpp::vec bg = to_vec(background, nB), bvar(nB, 1.0f);
    gridpp::vec obs = to_vec(pobs, nS), ovar = to_vec(obs_variance, nS), pbg = to_vec(pbackground, nS), pbvar(nS, 1.0f);
    gridpp::vec avar;
    gridpp::vec out = gridpp::optimal_interpolation_full(bp, bg, bvar, op, obs, ovar, pbg, pbvar, *s, max_points, avar, allow_extrapolation != 0);
    std::memcpy(analysis, out.data(), sizeof(float) * out.size());
    if(analysis_variance && avar.size() == (size_t) nB) std::memcpy(analysis_variance, avar.data(), sizeof(float) * nB);
    else if(analysis_variance) std::memcpy(analysis_variance, bvar.data(), sizeof(float) * nB);
    REF_CATCH
}

// gridpp::optimal_interpolation_ensi(Points...) oi_ensi.cpp:114-568. background is nB x nE, pbackground nS x nE.
int ref_optimal_interpolation_ensi(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB,
                                   const float* background, int nE, const float* plats, const float* plons,
                                   const float* pelevs, const float* plafs, int nS, int type, const float* pobs,
                                   const float* psigmas, const float* pbackground, const gpp_structure* sd, int max_points,
                                   int allow_extrapolation, float* analysis, double* seconds) {
    REF_TRY
    gridpp::Points bp = make_points(blats, blons, belevs, blafs, nB, type);
    gridpp::Points op = make_points(plats, plons, pelevs, plafs, nS, type);
    gridpp::StructureFunctionPtr s = make_structure(sd);
    gridpp::vec2 bg = to_vec2(background, nB, nE);
    gridpp::vec2 pbg = to_vec2(pbackground, nS, nE);
    gridpp::vec obs = to_vec(pobs, nS), sig = to_vec(psigmas, nS);
    if(nS > 0) op.get_neighbours(plats[0], plons[0], 1.0f);
    clk::time_point t0 = clk::now();
    gridpp::vec2 out = gridpp::optimal_interpolation_ensi(bp, bg, op, obs, sig, pbg, *s, max_points, allow_extrapolation != 0);
    if(seconds) *seconds = since(t0);
    from_vec2(out, analysis, nB, nE);
    REF_CATCH
}

// ---- neighbourhood ------------------------------------------------------------------------------------
// gridpp::neighbourhood(vec2, halfwidth, statistic) neighbourhood.cpp:28-242
int ref_neighbourhood(const float* input, int ny, int nx, int halfwidth, int statistic, float* output, double* seconds) {
    REF_TRY
    gridpp::vec2 in = to_vec2(input, ny, nx);
    clk::time_point t0 = clk::now();
    gridpp::vec2 out = gridpp::neighbourhood(in, halfwidth, (gridpp::Statistic) statistic);
    if(seconds) *seconds = since(t0);
    if(out.size() == (size_t) ny) from_vec2(out, output, ny, nx);
    REF_CATCH
}
// gridpp::neighbourhood_brute_force(vec2, ...) neighbourhood.cpp:528-530 -- the reference's own cross-check
int ref_neighbourhood_brute_force(const float* input, int ny, int nx, int halfwidth, int statistic, float* output) {
    REF_TRY
    gridpp::vec2 in = to_vec2(input, ny, nx);
    gridpp::vec2 out = gridpp::neighbourhood_brute_force(in, halfwidth, (gridpp::Statistic) statistic);
    if(out.size() == (size_t) ny) from_vec2(out, output, ny, nx);
    REF_CATCH
}
// gridpp::neighbourhood_brute_force(vec2 | vec3, ...) and gridpp::neighbourhood_quantile(vec2 | vec3, ...) (statistic ==
// Quantile), neighbourhood.cpp:528-539. input is ny x nx x ne (ne == 1: the vec2 overloads).
int ref_neighbourhood_window_ens(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float quantile, float* output) {
    REF_TRY
    gridpp::vec2 out;
    if(ne == 1) {
        gridpp::vec2 in = to_vec2(input, ny, nx);
        out = statistic == gridpp::Quantile ? gridpp::neighbourhood_quantile(in, quantile, halfwidth)
                                            : gridpp::neighbourhood_brute_force(in, halfwidth, (gridpp::Statistic) statistic);
    }
    else {
        gridpp::vec3 in(ny, gridpp::vec2(nx));
        for(int y = 0; y < ny; y++)
            for(int x = 0; x < nx; x++) in[y][x].assign(input + ((size_t) y * nx + x) * ne, input + ((size_t) y * nx + x + 1) * ne);
        out = statistic == gridpp::Quantile ? gridpp::neighbourhood_quantile(in, quantile, halfwidth)
                                            : gridpp::neighbourhood_brute_force(in, halfwidth, (gridpp::Statistic) statistic);
    }
    if(out.size() == (size_t) ny) from_vec2(out, output, ny, nx);
    REF_CATCH
}
// gridpp::neighbourhood_quantile_fast(vec2, float | vec2, halfwidth, thresholds) neighbourhood.cpp:296-409
int ref_neighbourhood_quantile_fast(const float* input, int ny, int nx, float quantile, const float* quantile_field,
                                    int halfwidth, const float* thresholds, int num_thresholds, float* output,
                                    double* seconds) {
    REF_TRY
    gridpp::vec2 in = to_vec2(input, ny, nx);
    gridpp::vec thr = to_vec(thresholds, num_thresholds);
    gridpp::vec2 out;
    clk::time_point t0 = clk::now();
    if(quantile_field) {
        gridpp::vec2 q = to_vec2(quantile_field, ny, nx);
        t0 = clk::now();
        out = gridpp::neighbourhood_quantile_fast(in, q, halfwidth, thr);
    }
    else
        out = gridpp::neighbourhood_quantile_fast(in, quantile, halfwidth, thr);
    if(seconds) *seconds = since(t0);
    if(out.size() == (size_t) ny) from_vec2(out, output, ny, nx);
    REF_CATCH
}
// gridpp::neighbourhood(vec3, ...) neighbourhood.cpp:12-27 and neighbourhood_quantile_fast(vec3, ...) :411-527
static gridpp::vec3 to_vec3(const float* in, int ny, int nx, int ne) {
    gridpp::vec3 out(ny);
    for(int y = 0; y < ny; y++) {
        out[y].resize(nx);
        for(int x = 0; x < nx; x++) out[y][x].assign(in + ((size_t) y * nx + x) * ne, in + ((size_t) y * nx + x + 1) * ne);
    }
    return out;
}
int ref_neighbourhood_ens(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float* output) {
    REF_TRY
    gridpp::vec2 out = gridpp::neighbourhood(to_vec3(input, ny, nx, ne), halfwidth, (gridpp::Statistic) statistic);
    if(out.size() == (size_t) ny) from_vec2(out, output, ny, nx);
    REF_CATCH
}
int ref_neighbourhood_quantile_fast_ens(const float* input, int ny, int nx, int ne, float quantile, const float* quantile_field,
                                        int halfwidth, const float* thresholds, int num_thresholds, float* output) {
    REF_TRY
    gridpp::vec3 in = to_vec3(input, ny, nx, ne);
    gridpp::vec thr = to_vec(thresholds, num_thresholds);
    gridpp::vec2 out;
    if(quantile_field) out = gridpp::neighbourhood_quantile_fast(in, to_vec2(quantile_field, ny, nx), halfwidth, thr);
    else out = gridpp::neighbourhood_quantile_fast(in, quantile, halfwidth, thr);
    if(out.size() == (size_t) ny) from_vec2(out, output, ny, nx);
    REF_CATCH
}
// gridpp::get_neighbourhood_thresholds(vec2, num) neighbourhood.cpp:243-266
int ref_get_neighbourhood_thresholds(const float* input, int ny, int nx, int num, float* out, int* out_n) {
    REF_TRY
    gridpp::vec t = gridpp::get_neighbourhood_thresholds(to_vec2(input, ny, nx), num);
    *out_n = t.size();
    for(size_t i = 0; i < t.size() && (int) i < num; i++) out[i] = t[i];
    REF_CATCH
}
// gridpp::interpolate(float, vec, vec) util.cpp:377-414
int ref_interpolate(float x, const float* ix, const float* iy, int n, float* out) {
    REF_TRY
    *out = gridpp::interpolate(x, to_vec(ix, n), to_vec(iy, n));
    REF_CATCH
}
// gridpp::calc_statistic / calc_quantile util.cpp:19-178
int ref_calc_statistic(const float* a, int n, int statistic, float* out) {
    REF_TRY
    *out = gridpp::calc_statistic(to_vec(a, n), (gridpp::Statistic) statistic);
    REF_CATCH
}
int ref_calc_quantile(const float* a, int n, float q, float* out) {
    REF_TRY
    *out = gridpp::calc_quantile(to_vec(a, n), q);
    REF_CATCH
}

// ---- consumers of the point index: gridding.cpp, count.cpp, distance.cpp, fill.cpp, doping.cpp -----------------------
// A location set is given as flat lat / lon arrays plus (ny, nx): nx > 0 -> a gridpp::Grid of ny x nx, nx == 0 -> gridpp::Points
// of ny locations. Results are flat in the same order.
namespace {
gridpp::Grid make_grid(const float* lats, const float* lons, const float* elevs, int ny, int nx, int type) {
    return gridpp::Grid(to_vec2(lats, ny, nx), to_vec2(lons, ny, nx), elevs ? to_vec2(elevs, ny, nx) : gridpp::vec2(), gridpp::vec2(),
                        (gridpp::CoordinateType) type);
}
void from_any(const gridpp::vec2& v, float* out, int ny, int nx) { from_vec2(v, out, ny, nx); }
}  // namespace

int ref_gridding(const float* olats, const float* olons, int ony, int onx, const float* ilats, const float* ilons, int nI, int type,
                 const float* values, float radius, int min_num, int statistic, float* output) {
    REF_TRY
    gridpp::Points ip = make_points(ilats, ilons, nullptr, nullptr, nI, type);
    gridpp::vec v = to_vec(values, nI);
    if(onx > 0) from_any(gridpp::gridding(make_grid(olats, olons, nullptr, ony, onx, type), ip, v, radius, min_num, (gridpp::Statistic) statistic), output, ony, onx);
    else {
        gridpp::vec out = gridpp::gridding(make_points(olats, olons, nullptr, nullptr, ony, type), ip, v, radius, min_num, (gridpp::Statistic) statistic);
        std::memcpy(output, out.data(), sizeof(float) * out.size());
    }
    REF_CATCH
}
int ref_gridding_nearest(const float* olats, const float* olons, int ony, int onx, const float* ilats, const float* ilons, int nI, int type,
                         const float* values, int min_num, int statistic, float* output) {
    REF_TRY
    gridpp::Points ip = make_points(ilats, ilons, nullptr, nullptr, nI, type);
    gridpp::vec v = to_vec(values, nI);
    if(onx > 0) from_any(gridpp::gridding_nearest(make_grid(olats, olons, nullptr, ony, onx, type), ip, v, min_num, (gridpp::Statistic) statistic), output, ony, onx);
    else {
        gridpp::vec out = gridpp::gridding_nearest(make_points(olats, olons, nullptr, nullptr, ony, type), ip, v, min_num, (gridpp::Statistic) statistic);
        std::memcpy(output, out.data(), sizeof(float) * out.size());
    }
    REF_CATCH
}
// count / distance: (input set, output set), each a Grid (nx > 0) or Points
int ref_count(const float* ilats, const float* ilons, int iny, int inx, const float* olats, const float* olons, int ony, int onx, int type,
              float radius, float* output) {
    REF_TRY
    if(inx > 0 && onx > 0) from_any(gridpp::count(make_grid(ilats, ilons, nullptr, iny, inx, type), make_grid(olats, olons, nullptr, ony, onx, type), radius), output, ony, onx);
    else if(inx > 0) { gridpp::vec o = gridpp::count(make_grid(ilats, ilons, nullptr, iny, inx, type), make_points(olats, olons, nullptr, nullptr, ony, type), radius); std::memcpy(output, o.data(), sizeof(float) * o.size()); }
    else if(onx > 0) from_any(gridpp::count(make_points(ilats, ilons, nullptr, nullptr, iny, type), make_grid(olats, olons, nullptr, ony, onx, type), radius), output, ony, onx);
    else { gridpp::vec o = gridpp::count(make_points(ilats, ilons, nullptr, nullptr, iny, type), make_points(olats, olons, nullptr, nullptr, ony, type), radius); std::memcpy(output, o.data(), sizeof(float) * o.size()); }
    REF_CATCH
}
int ref_distance(const float* ilats, const float* ilons, int iny, int inx, const float* olats, const float* olons, int ony, int onx, int type,
                 int num, float* output) {
    REF_TRY
    if(inx > 0 && onx > 0) from_any(gridpp::distance(make_grid(ilats, ilons, nullptr, iny, inx, type), make_grid(olats, olons, nullptr, ony, onx, type), num), output, ony, onx);
    else if(inx > 0) { gridpp::vec o = gridpp::distance(make_grid(ilats, ilons, nullptr, iny, inx, type), make_points(olats, olons, nullptr, nullptr, ony, type), num); std::memcpy(output, o.data(), sizeof(float) * o.size()); }
    else if(onx > 0) from_any(gridpp::distance(make_points(ilats, ilons, nullptr, nullptr, iny, type), make_grid(olats, olons, nullptr, ony, onx, type), num), output, ony, onx);
    else { gridpp::vec o = gridpp::distance(make_points(ilats, ilons, nullptr, nullptr, iny, type), make_points(olats, olons, nullptr, nullptr, ony, type), num); std::memcpy(output, o.data(), sizeof(float) * o.size()); }
    REF_CATCH
}
int ref_fill(const float* glats, const float* glons, int ny, int nx, const float* input, const float* plats, const float* plons, int nP, int type,
             const float* radii, float value, int outside, float* output) {
    REF_TRY
    from_any(gridpp::fill(make_grid(glats, glons, nullptr, ny, nx, type), to_vec2(input, ny, nx), make_points(plats, plons, nullptr, nullptr, nP, type),
                          to_vec(radii, nP), value, outside != 0), output, ny, nx);
    REF_CATCH
}
int ref_fill_missing(const float* values, int ny, int nx, float* output) {
    REF_TRY
    from_any(gridpp::fill_missing(to_vec2(values, ny, nx)), output, ny, nx);
    REF_CATCH
}
int ref_doping_square(const float* glats, const float* glons, const float* gelevs, int ny, int nx, const float* background, const float* plats,
                      const float* plons, const float* pelevs, int nP, int type, const float* obs, const int* halfwidth, float max_elev_diff,
                      float* output) {
    REF_TRY
    from_any(gridpp::doping_square(make_grid(glats, glons, gelevs, ny, nx, type), to_vec2(background, ny, nx),
                                   make_points(plats, plons, pelevs, nullptr, nP, type), to_vec(obs, nP), gridpp::ivec(halfwidth, halfwidth + nP),
                                   max_elev_diff), output, ny, nx);
    REF_CATCH
}
int ref_doping_circle(const float* glats, const float* glons, const float* gelevs, int ny, int nx, const float* background, const float* plats,
                      const float* plons, const float* pelevs, int nP, int type, const float* obs, const float* radii, float max_elev_diff,
                      float* output) {
    REF_TRY
    from_any(gridpp::doping_circle(make_grid(glats, glons, gelevs, ny, nx, type), to_vec2(background, ny, nx),
                                   make_points(plats, plons, pelevs, nullptr, nP, type), to_vec(obs, nP), to_vec(radii, nP), max_elev_diff), output, ny, nx);
    REF_CATCH
}

}  // extern "C"

// ---- ensi_multi / staticcorr_points -------------------------------------------------------------------
extern "C" {
// gridpp::staticcorr_points, corr_points.cpp:26-131. output nY x nS
int ref_staticcorr_points(const float* lats, const float* lons, const float* elevs, const float* lafs, int nY, const float* klats,
                          const float* klons, const float* kelevs, const float* klafs, int nS, int type, const gpp_structure* sd,
                          int max_points, float* output) {
    REF_TRY
    gridpp::Points p = make_points(lats, lons, elevs, lafs, nY, type);
    gridpp::Points k = make_points(klats, klons, kelevs, klafs, nS, type);
    gridpp::StructureFunctionPtr s = make_structure(sd);
    gridpp::vec2 out = gridpp::staticcorr_points(p, k, *s, max_points);
    from_vec2(out, output, nY, nS);
    REF_CATCH
}
// gridpp::optimal_interpolation_ensi_multi_ebe / _ebesc (Points overloads), oi_ensi_multi.cpp:329-628 / :630-859
int ref_ensi_multi_ebe(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB, const float* bratios,
                       const float* background, const float* background_corr, int nE, const float* plats, const float* plons,
                       const float* pelevs, const float* plafs, int nS, int type, const float* pobs, const float* pratios,
                       const float* pbackground, const float* pbackground_corr, const gpp_structure* sd, int max_points,
                       int allow_extrapolation, int with_ens, float* analysis) {
    REF_TRY
    gridpp::Points bp = make_points(blats, blons, belevs, blafs, nB, type);
    gridpp::Points op = make_points(plats, plons, pelevs, plafs, nS, type);
    gridpp::StructureFunctionPtr s = make_structure(sd);
    gridpp::vec2 bg = to_vec2(background, nB, nE), pbg = to_vec2(pbackground, nS, nE), obs = to_vec2(pobs, nS, nE);
    gridpp::vec br = to_vec(bratios, nB), pr = to_vec(pratios, nS);
    gridpp::vec2 out;
    if(with_ens) {
        gridpp::vec2 bgc = to_vec2(background_corr, nB, nE), pbgc = to_vec2(pbackground_corr, nS, nE);
        out = gridpp::optimal_interpolation_ensi_multi_ebe(bp, br, bg, bgc, op, obs, pr, pbg, pbgc, *s, max_points, allow_extrapolation != 0);
    }
    else out = gridpp::optimal_interpolation_ensi_multi_ebesc(bp, br, bg, op, obs, pr, pbg, *s, max_points, allow_extrapolation != 0);
    from_vec2(out, analysis, nB, nE);
    REF_CATCH
}
// gridpp::optimal_interpolation_ensi_multi_utem (Points overload), oi_ensi_multi.cpp:862-1311
int ref_ensi_multi_utem(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB, const float* bratios,
                        const float* background, const float* background_corr, int nE, const float* plats, const float* plons,
                        const float* pelevs, const float* plafs, int nS, int type, const float* pobs, const float* pratios,
                        const float* pbackground, const float* pbackground_corr, const gpp_structure* sd, int max_points,
                        int allow_extrapolation, float* analysis) {
    REF_TRY
    gridpp::Points bp = make_points(blats, blons, belevs, blafs, nB, type);
    gridpp::Points op = make_points(plats, plons, pelevs, plafs, nS, type);
    gridpp::StructureFunctionPtr s = make_structure(sd);
    gridpp::vec2 bg = to_vec2(background, nB, nE), pbg = to_vec2(pbackground, nS, nE);
    gridpp::vec2 bgc = to_vec2(background_corr, nB, nE), pbgc = to_vec2(pbackground_corr, nS, nE);
    gridpp::vec br = to_vec(bratios, nB), pr = to_vec(pratios, nS), obs = to_vec(pobs, nS);
    gridpp::vec2 out = gridpp::optimal_interpolation_ensi_multi_utem(bp, br, bg, bgc, op, obs, pr, pbg, pbgc, *s, max_points, allow_extrapolation != 0);
    from_vec2(out, analysis, nB, nE);
    REF_CATCH
}
}

// ---- neighbourhood_search / calc_gradient -------------------------------------------------------------
extern "C" {
// gridpp::neighbourhood_search, neighbourhood_search.cpp:7-113; apply == NULL -> the default empty ivec2
int ref_neighbourhood_search(const float* array, const float* search, int ny, int nx, int halfwidth, float tmin, float tmax, float delta,
                             const int* apply, float* output) {
    REF_TRY
    gridpp::vec2 a = to_vec2(array, ny, nx), s = to_vec2(search, ny, nx);
    gridpp::ivec2 ap;
    if(apply) {
        ap.assign(ny, gridpp::ivec(nx));
        for(int y = 0; y < ny; y++)
            for(int x = 0; x < nx; x++) ap[y][x] = apply[(size_t) y * nx + x];
    }
    gridpp::vec2 out = gridpp::neighbourhood_search(a, s, halfwidth, tmin, tmax, delta, ap);
    from_vec2(out, output, ny, nx);
    REF_CATCH
}
// gridpp::calc_gradient, calc_gradient.cpp:6-126
int ref_calc_gradient(const float* base, const float* values, int ny, int nx, int gradient_type, int halfwidth, int num_min, float min_range,
                      float default_gradient, float* output) {
    REF_TRY
    gridpp::vec2 b = to_vec2(base, ny, nx), v = to_vec2(values, ny, nx);
    gridpp::vec2 out = gridpp::calc_gradient(b, v, (gridpp::GradientType) gradient_type, halfwidth, num_min, min_range, default_gradient);
    from_vec2(out, output, ny, nx);
    REF_CATCH
}
}
