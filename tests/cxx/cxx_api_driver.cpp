// Test driver for the C++ host layer (include/gridpp.h). Written like a user program of the reference's C++ API.
//   cxx_api_driver checks          host-side behaviour that needs no device (validation, containers, exceptions)
//   cxx_api_driver run <dir>       reads <dir>/in_*.bin + <dir>/meta.txt, runs the whole path on the GPU, writes <dir>/out_*.bin
#include <cstdio>
#include <fstream>
#include <map>

#include "gridpp.h"

using namespace gridpp;

static std::string g_dir;

static vec read_f32(const std::string& name) {
    std::ifstream f(g_dir + "/in_" + name + ".bin", std::ios::binary | std::ios::ate);
    if(!f) throw std::runtime_error("cannot open input " + name);
    const size_t bytes = (size_t) f.tellg();
    vec out(bytes / sizeof(float));
    f.seekg(0);
    f.read((char*) out.data(), bytes);
    return out;
}
template <class T>
static void write_raw(const std::string& name, const std::vector<T>& v) {
    std::ofstream f(g_dir + "/out_" + name + ".bin", std::ios::binary);
    f.write((const char*) v.data(), v.size() * sizeof(T));
}
static void write(const std::string& name, const vec& v) { write_raw(name, v); }
static void write(const std::string& name, const ivec& v) { write_raw(name, v); }
static void write(const std::string& name, const vec2& v) { write_raw(name, b200::flatten(v)); }
static void write(const std::string& name, const vec3& v) { write_raw(name, b200::flatten(v)); }
static void write(const std::string& name, const ivec2& v) {
    ivec flat;
    for(const ivec& r : v) flat.insert(flat.end(), r.begin(), r.end());
    write_raw(name, flat);
}

#define EXPECT_THROW(expr, Exception)                                                        \
    do {                                                                                     \
        bool thrown = false;                                                                 \
        try { (void) (expr); }                                                               \
        catch(const Exception&) { thrown = true; }                                           \
        if(!thrown) { std::printf("FAILED line %d: %s did not throw " #Exception "\n", __LINE__, #expr); return 1; } \
    } while(0)
#define EXPECT(cond)                                                                         \
    do {                                                                                     \
        if(!(cond)) { std::printf("FAILED line %d: %s\n", __LINE__, #cond); return 1; }      \
    } while(0)

static int checks() {
    // containers and accessors (points.cpp:9-31, grid.cpp:12-55)
    Points points({60, 61, 62}, {10, 11, 12}, {100, 200, 300});
    EXPECT(points.size() == 3 && points.get_coordinate_type() == Geodetic);
    EXPECT(points.get_elevs()[1] == 200 && std::isnan(points.get_lafs()[0]));
    EXPECT_THROW(Points({60, 61}, {10}), std::invalid_argument);
    EXPECT_THROW(Points({60, 61}, {10, 11}, {1}), std::invalid_argument);
    EXPECT_THROW(Points({91}, {10}), std::invalid_argument);                 // invalid latitude, util.cpp:596-600
    Grid grid({{0, 0, 0}, {1, 1, 1}}, {{0, 1, 2}, {0, 1, 2}}, vec2(), vec2(), Cartesian);
    EXPECT(grid.size()[0] == 2 && grid.size()[1] == 3 && grid.get_lons()[1][2] == 2 && std::isnan(grid.get_elevs()[0][0]));
    EXPECT(grid.to_points().size() == 6 && grid.to_points().get_coordinate_type() == Cartesian);
    EXPECT_THROW(Grid({{0, 0}}, {{0, 1, 2}}), std::invalid_argument);
    Point p = grid.get_point(1, 2);
    EXPECT(p.x == 2 && p.y == 1 && p.z == 0);                               // the index's coordinates (grid.cpp:232-235)
    Point direct(1, 2, 0, 0, Cartesian);
    EXPECT(direct.x == 1 && direct.y == 2 && direct.z == 0);                // point.cpp:18-22: x = lat, y = lon
    EXPECT(std::isnan(Point(MV, 0, 0, 0, Cartesian).x));                    // ... and no validation for Cartesian points
    EXPECT_THROW(Point(95, 0), std::invalid_argument);
    Point q(0, 90);   // on the equator at 90 E: x ~ 0, y = earth radius
    EXPECT(std::fabs(q.y - 6.378137e6f) < 1 && std::fabs(q.x) < 1 && q.z == 0);
    EXPECT(KDTree::calc_distance(0, 0, 3, 4, Cartesian) == 5);
    EXPECT(std::fabs(KDTree::calc_distance(60, 10, 60, 10.001f) - 55.66f) < 0.1f);
    EXPECT(KDTree::calc_straight_distance(0, 0, 0, 1, 2, 2) == 3);
    // structure functions: constructor validation (structure.cpp:145-152, 287-292, 911-913) and localization distances
    BarnesStructure barnes(10000);
    EXPECT(std::fabs(barnes.localization_distance(p) - 36456.5f) < 1);      // sqrt(-2 log 0.0013) * h
    EXPECT_THROW(BarnesStructure(-1), std::invalid_argument);
    EXPECT_THROW(BarnesStructure(1000, -1), std::invalid_argument);
    EXPECT_THROW(CressmanStructure(MV), std::invalid_argument);
    EXPECT_THROW(CrossValidation(barnes, -1), std::invalid_argument);
    CrossValidation cv(barnes, 500);
    StructureFunctionPtr copy = cv.clone();
    EXPECT(copy->localization_distance(p) == barnes.localization_distance(p));
    EXPECT(BarnesStructure(grid, {{2000}}, {{0}}, {{0}}, 0.5f).localization_distance(p) == std::sqrt(-2 * std::log(0.5f)) * 2000);
    EXPECT_THROW(BarnesStructure(grid, {{1, 2}, {3, 4}}, {{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}), std::invalid_argument);   // not the grid's shape
    // argument checks of the compute entry points come before any device work (oi.cpp:38-63, neighbourhood.cpp:29-32)
    vec2 field = {{1, 2, 3}, {4, 5, 6}};
    Points obs_points({0.5f}, {0.5f}, vec(), vec(), Cartesian);
    EXPECT_THROW(optimal_interpolation(grid, field, obs_points, {1}, {0.5f}, {1}, barnes, -1), std::invalid_argument);
    EXPECT_THROW(optimal_interpolation(grid, field, obs_points, {1, 2}, {0.5f}, {1}, barnes, 10), std::invalid_argument);
    EXPECT_THROW(optimal_interpolation(grid, {{1, 2}}, obs_points, {1}, {0.5f}, {1}, barnes, 10), std::invalid_argument);
    EXPECT_THROW(optimal_interpolation(grid, field, points, {1, 2, 3}, {1, 1, 1}, {1, 2, 3}, barnes, 10), std::invalid_argument);   // coordinate types differ
    EXPECT_THROW(neighbourhood(field, -1, Mean), std::invalid_argument);
    EXPECT_THROW(neighbourhood(field, 1, Quantile), std::invalid_argument);
    EXPECT_THROW(nearest(grid, obs_points, vec2{{1, 2}}), std::invalid_argument);
    EXPECT(neighbourhood(vec2(), 1, Mean).empty());                         // neighbourhood.cpp:33-34
    EXPECT(neighbourhood_brute_force(vec2(), 1, Median).empty() && neighbourhood_quantile(vec2(), 0.5f, 1).empty());
    EXPECT_THROW(neighbourhood_quantile(field, 0.5f, -1), std::invalid_argument);
    EXPECT_THROW(interpolate(0.5f, {0, 1, 2}, {0, 1}), std::invalid_argument);   // util.cpp:380-381
    EXPECT(get_statistic("median") == Median && get_statistic("variance") == Unknown && get_statistic("mean") == Mean);
    EXPECT(is_valid(1.f) && !is_valid(MV));
    EXPECT(get_neighbourhood_thresholds(vec2(), 5).empty());
    // no observations: EnSI returns the background untouched without touching the device (oi_ensi.cpp:49-51)
    vec3 ens = {{{1, 2}, {3, 4}, {5, 6}}, {{7, 8}, {9, 10}, {11, 12}}};
    EXPECT(optimal_interpolation_ensi(grid, ens, Points(vec(), vec(), vec(), vec(), Cartesian), vec(), vec(), vec2(), barnes, 5) == ens);
    EXPECT(optimal_interpolation_ensi_multi_ebesc(grid, vec2(), ens, Points(vec(), vec(), vec(), vec(), Cartesian), vec2(), vec(), vec2(), barnes, 5) == ens);
    EXPECT_THROW(staticcorr_points(grid.to_points(), grid.to_points(), barnes, -1), std::invalid_argument);   // corr_points.cpp:34-35
    EXPECT_THROW(neighbourhood_search(field, field, 1, 2.f, 1.f, 0.f), std::invalid_argument);   // neighbourhood_search.cpp:10-12
    EXPECT_THROW(calc_gradient(field, field, MinMax, 0), std::invalid_argument);                 // calc_gradient.cpp:9-10
    EXPECT(init_vec2(2, 3, 1.5f) == vec2(2, vec(3, 1.5f)) && init_ivec3(1, 2, 2, -1)[0][1][1] == -1 && compatible_size(field, field));
    EXPECT(point_in_rectangle(Point(0, 0), Point(0, 1), Point(1, 1), Point(1, 0), Point(0.5f, 0.5f)));
    EXPECT(!point_in_rectangle(Point(0, 0), Point(0, 1), Point(1, 1), Point(1, 0), Point(1.5f, 0.5f)));
    set_omp_threads(4);
    EXPECT(get_omp_threads() == 4);
    initialize_omp();
    // the compute path has no CPU fallback: with no device it must throw, never return numbers
    int devices = 0;
    gpp_device_count(&devices);
    if(devices == 0) EXPECT_THROW(neighbourhood(field, 1, Mean), std::runtime_error);
    std::printf("checks ok (version %s, %d device(s))\n", version().c_str(), devices);
    return 0;
}

static int run() {
    std::map<std::string, double> meta;
    {
        std::ifstream f(g_dir + "/meta.txt");
        std::string key;
        double value;
        while(f >> key >> value) meta[key] = value;
    }
    const int ny = (int) meta["ny"], nx = (int) meta["nx"], nS = (int) meta["nS"], nE = (int) meta["nE"], hw = (int) meta["halfwidth"];
    const int max_points = (int) meta["max_points"];
    const CoordinateType ctype = (CoordinateType) (int) meta["ctype"];
    const float quantile = (float) meta["quantile"];
    const float radius = (float) meta["radius"];

    Grid grid(b200::unflatten(read_f32("glats"), ny, nx), b200::unflatten(read_f32("glons"), ny, nx), b200::unflatten(read_f32("gelevs"), ny, nx),
              b200::unflatten(read_f32("glafs"), ny, nx), ctype);
    Points points(read_f32("plats"), read_f32("plons"), read_f32("pelevs"), read_f32("plafs"), ctype);
    const vec2 background = b200::unflatten(read_f32("background"), ny, nx);
    const vec2 bvariance = b200::unflatten(read_f32("bvariance"), ny, nx);
    const vec3 ensemble = b200::unflatten(read_f32("ensemble"), ny, nx, nE);
    const vec obs = read_f32("obs"), ratios = read_f32("ratios"), sigmas = read_f32("sigmas"), thresholds = read_f32("thresholds");
    const vec2 quantile_field = b200::unflatten(read_f32("quantile_field"), ny, nx);

    // README.md:42-58 in C++
    BarnesStructure structure((float) meta["h"], (float) meta["v"], (float) meta["w"]);
    const vec pbackground = nearest(grid, points, background);
    write("pbackground", pbackground);
    write("oi", optimal_interpolation(grid, background, points, obs, ratios, pbackground, structure, max_points));
    vec2 variance;
    const vec pbvariance = nearest(grid, points, bvariance);
    write("full", optimal_interpolation_full(grid, background, bvariance, points, obs, ratios, pbackground, pbvariance, structure, max_points, variance));
    write("full_variance", variance);
    // Points overload with a composite structure function, no extrapolation
    BarnesStructure sh((float) meta["h"]);
    LinearStructure sv(0, 0.2f, 0);
    CressmanStructure sw(0, 0, 0.7f);
    MultipleStructure multiple(sh, sv, sw);
    CrossValidation cv(multiple, (float) meta["cv_dist"]);
    write("oi_cv", optimal_interpolation(grid.to_points(), b200::flatten(background), points, obs, ratios, pbackground, cv, max_points, false));
    // ensemble: members at the observation points through nearest(Grid, Points, vec3) with the member index leading
    vec3 members((size_t) nE, vec2((size_t) ny, vec((size_t) nx)));
    for(int e = 0; e < nE; e++)
        for(int y = 0; y < ny; y++)
            for(int x = 0; x < nx; x++) members[e][y][x] = ensemble[y][x][e];
    const vec2 by_member = nearest(grid, points, members);
    vec2 pensemble((size_t) nS, vec((size_t) nE));
    for(int s = 0; s < nS; s++)
        for(int e = 0; e < nE; e++) pensemble[s][e] = by_member[e][s];
    write("pensemble", pensemble);
    write("ensi", optimal_interpolation_ensi(grid, ensemble, points, obs, sigmas, pensemble, structure, max_points));
    {   // the "multi" variants (oi_ensi_multi.cpp) and staticcorr_points (corr_points.cpp); the ensemble doubles as its own *_corr
        vec2 pobs2((size_t) nS, vec((size_t) nE));
        for(int s = 0; s < nS; s++)
            for(int e = 0; e < nE; e++) pobs2[s][e] = obs[s] + 0.125f * e;
        write("ebesc", optimal_interpolation_ensi_multi_ebesc(grid, bvariance, ensemble, points, pobs2, ratios, pensemble, structure, max_points, false));
        write("ebe", optimal_interpolation_ensi_multi_ebe(grid, bvariance, ensemble, ensemble, points, pobs2, ratios, pensemble, pensemble, structure, max_points));
        write("utem", optimal_interpolation_ensi_multi_utem(grid, bvariance, ensemble, ensemble, points, obs, ratios, pensemble, pensemble, structure, max_points, false));
        write("staticcorr", staticcorr_points(Points(read_f32("qlats"), read_f32("qlons"), vec(), vec(), Cartesian), points, structure, 5));
    }

    write("mean", neighbourhood(background, hw, Mean));
    write("min", neighbourhood(background, hw, Min));
    write("max", neighbourhood(background, hw, Max));
    write("ens_mean", neighbourhood(ensemble, hw, Mean));
    write("qf", neighbourhood_quantile_fast(background, quantile, hw, thresholds));
    write("qf_field", neighbourhood_quantile_fast(background, quantile_field, hw, thresholds));
    write("qf_ens", neighbourhood_quantile_fast(ensemble, quantile, hw, thresholds));
    write("thresholds", get_neighbourhood_thresholds(background, (int) thresholds.size()));
    // consumers of the point index (gridding.cpp, count.cpp, distance.cpp, fill.cpp, doping.cpp)
    write("gridding", gridding(grid, points, obs, radius, 2, Mean));
    write("gridding_nearest", gridding_nearest(grid, points, obs, 0, Max));
    write("count", count(points, grid, radius));
    write("distance", distance(grid, points, 3));
    write("fill", fill(grid, background, points, vec((size_t) nS, radius / 3), -1.f, false));
    write("fill_missing", fill_missing(background));
    write("doping_circle", doping_circle(grid, background, points, obs, vec((size_t) nS, radius / 3), 100.f));
    write("doping_square", doping_square(grid, background, points, obs, ivec((size_t) nS, 1)));
    // statistics family (util.cpp:19-215,377-431; neighbourhood.cpp:211-238,528-539)
    write("search", neighbourhood_search(background, bvariance, 2, 1.5f, 2.f, 0.2f));
    write("gradient_minmax", calc_gradient(bvariance, background, MinMax, 2, 3, 0.1f, -1.f));
    write("gradient_regression", calc_gradient(bvariance, background, LinearRegression, 2));
    write("nbh_std", neighbourhood(background, hw, Std));
    write("nbh_median", neighbourhood(background, hw, Median));
    write("nbh_quantile", neighbourhood_quantile(background, quantile, hw));
    write("nbh_quantile_ens", neighbourhood_quantile(ensemble, quantile, hw));
    write("brute_variance", neighbourhood_brute_force(background, hw, Variance));
    {
        vec2 rows((size_t) ny * nx);
        for(int y = 0; y < ny; y++)
            for(int x = 0; x < nx; x++) rows[(size_t) y * nx + x] = ensemble[y][x];
        write("member_median", calc_statistic(rows, Median));
        write("member_q", calc_quantile(rows, quantile));
        write("member_q_field", calc_quantile(ensemble, quantile_field));
        write("interp", interpolate(thresholds, thresholds, calc_statistic(vec2(thresholds.size(), thresholds), Sum)));
        if(calc_statistic(rows[3], Mean) != calc_statistic(vec2(1, rows[3]), Mean)[0]) throw std::runtime_error("calc_statistic(vec) != calc_statistic(vec2)[0]");
    }
    write("nearest_grid", nearest(points, grid, obs));

    // index queries (kdtree.cpp:18-106): integer results, compared bit for bit
    const vec qlats = read_f32("qlats"), qlons = read_f32("qlons");
    ivec nn, counts, neighbours, closest, grid_nn;
    vec distances;
    for(size_t i = 0; i < qlats.size(); i++) {
        nn.push_back(points.get_nearest_neighbour(qlats[i], qlons[i]));
        vec d;
        const ivec found = points.get_neighbours_with_distance(qlats[i], qlons[i], radius, d);
        if(found != points.get_neighbours(qlats[i], qlons[i], radius)) throw std::runtime_error("get_neighbours and get_neighbours_with_distance disagree");
        counts.push_back(points.get_num_neighbours(qlats[i], qlons[i], radius));
        if(counts.back() != (int) found.size()) throw std::runtime_error("get_num_neighbours disagrees with get_neighbours");
        neighbours.insert(neighbours.end(), found.begin(), found.end());
        distances.insert(distances.end(), d.begin(), d.end());
        const ivec near5 = points.get_closest_neighbours(qlats[i], qlons[i], 5);
        closest.insert(closest.end(), near5.begin(), near5.end());
        const ivec yx = grid.get_nearest_neighbour(qlats[i], qlons[i]);
        grid_nn.insert(grid_nn.end(), yx.begin(), yx.end());
    }
    write("grid_neighbours", grid.get_neighbours(qlats[0], qlons[0], radius));
    write("nn", nn);
    write("counts", counts);
    write("neighbours", neighbours);
    write("distances", distances);
    write("closest", closest);
    write("grid_nn", grid_nn);

    // structure function values between point pairs
    std::vector<Point> others;
    for(int s = 1; s < nS; s++) others.push_back(points.get_point(s));
    write("corr", structure.corr(points.get_point(0), others));
    write("corr_cv_background", cv.corr_background(points.get_point(0), others));
    std::printf("run ok\n");
    return 0;
}

int main(int argc, char** argv) {
    try {
        if(argc >= 2 && std::string(argv[1]) == "checks") return checks();
        if(argc >= 3 && std::string(argv[1]) == "run") {
            g_dir = argv[2];
            return run();
        }
        std::printf("usage: cxx_api_driver checks | run <dir>\n");
        return 2;
    }
    catch(const std::exception& e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 1;
    }
}
