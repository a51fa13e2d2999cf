// gridpp.h -- the C++ host layer of the B200-native gridpp hot path.
//
// A header-only `namespace gridpp` with the names, argument order, defaults, value-returning conventions and exception
// types of the reference's public C++ API for the data-parallel path (reference include/gridpp.h: OI :162-294, the
// neighbourhood family :588-679, nearest :879-935, OpenMP control :1378-1386, Point :1713-1743, KDTree :1746-1873,
// Points :1876-1968, Grid :1971-2060, StructureFunction hierarchy :2069-2343). Nothing is computed here: every call
// flattens its std::vector arguments and forwards to the C ABI of libgridpp_b200.so (include/gridpp_b200.h), whose
// kernels are hand-written sm_100a CUDA. There is no CPU fallback: without a usable device the compute calls throw
// std::runtime_error.
//
// Build: g++ -std=c++14 -I include app.cpp -L gridpp_b200 -lgridpp_b200
#ifndef GRIDPP_B200_CXX_H
#define GRIDPP_B200_CXX_H

#include <cmath>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gridpp_b200.h"

namespace gridpp {

typedef std::vector<float> vec;
typedef std::vector<vec> vec2;
typedef std::vector<vec2> vec3;
typedef std::vector<int> ivec;
typedef std::vector<ivec> ivec2;

static const float MV = NAN;   // missing value (reference gridpp.h:49)

enum Statistic { Mean = 0, Min = 10, Median = 20, Max = 30, Quantile = 40, Std = 50, Variance = 60, Sum = 70, Count = 80, RandomChoice = 90, Unknown = -1 };
enum CoordinateType { Geodetic = 0, Cartesian = 1 };
enum GradientType { MinMax = 0, LinearRegression = 10 };   // reference gridpp.h:126-129

class not_implemented_exception : public std::logic_error {
public:
    not_implemented_exception() : std::logic_error("Function not yet implemented") {}
    explicit not_implemented_exception(const std::string& what) : std::logic_error(what) {}
};

// -------------------------------------------------------------------------------------------------------------------
// marshalling between nested std::vectors and the dense row-major buffers of the C ABI
namespace b200 {
// status code -> the exception the reference throws (swig/gridpp.i:21-40 maps them on to ValueError / RuntimeError)
inline void check(int rc) {
    if(rc == GPP_OK) return;
    const std::string msg = gpp_last_error();
    if(rc == GPP_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    if(rc == GPP_ERR_NOT_IMPLEMENTED) throw not_implemented_exception(msg);
    throw std::runtime_error(msg);
}
inline void require(bool ok, const std::string& msg) {
    if(!ok) throw std::invalid_argument(msg);
}
// rows x cols of a vec2; every row must have the length of the first
// How many GPUs optimal_interpolation / optimal_interpolation_full use from this process: 1 (default), n, or 0 = every visible
// device. The rows of the background grid are split into one block per device; results do not depend on the setting.
inline int& devices() {
    static int n = 1;
    return n;
}
inline void use_devices(int n) { devices() = n < 0 ? 1 : n; }

inline void shape_of(const vec2& a, int& rows, int& cols, const char* name) {
    rows = (int) a.size();
    cols = rows ? (int) a[0].size() : 0;
    for(const vec& r : a) require((int) r.size() == cols, std::string(name) + " is not a rectangular array");
}
inline void shape_of(const vec3& a, int& n0, int& n1, int& n2, const char* name) {
    n0 = (int) a.size();
    n1 = n0 ? (int) a[0].size() : 0;
    n2 = (n0 && n1) ? (int) a[0][0].size() : 0;
    for(const vec2& p : a) {
        require((int) p.size() == n1, std::string(name) + " is not a rectangular array");
        for(const vec& r : p) require((int) r.size() == n2, std::string(name) + " is not a rectangular array");
    }
}
inline vec flatten(const vec2& a, const char* name = "input") {
    int rows, cols;
    shape_of(a, rows, cols, name);
    vec flat;
    flat.reserve((size_t) rows * cols);
    for(const vec& r : a) flat.insert(flat.end(), r.begin(), r.end());
    return flat;
}
inline vec flatten(const vec3& a, const char* name = "input") {
    int n0, n1, n2;
    shape_of(a, n0, n1, n2, name);
    vec flat;
    flat.reserve((size_t) n0 * n1 * n2);
    for(const vec2& p : a)
        for(const vec& r : p) flat.insert(flat.end(), r.begin(), r.end());
    return flat;
}
inline vec2 unflatten(const vec& flat, int rows, int cols) {
    vec2 out((size_t) rows);
    for(int r = 0; r < rows; r++) out[r].assign(flat.begin() + (size_t) r * cols, flat.begin() + (size_t) (r + 1) * cols);
    return out;
}
inline vec3 unflatten(const vec& flat, int n0, int n1, int n2) {
    vec3 out((size_t) n0, vec2((size_t) n1));
    size_t at = 0;
    for(int i = 0; i < n0; i++)
        for(int j = 0; j < n1; j++, at += (size_t) n2) out[i][j].assign(flat.begin() + at, flat.begin() + at + (size_t) n2);
    return out;
}
inline const float* ptr_or_null(const vec& v) { return v.empty() ? nullptr : v.data(); }
typedef std::shared_ptr<gpp_points> PointsHandle;
inline PointsHandle make_points(const vec& lats, const vec& lons, const float* elevs, const float* lafs, CoordinateType type) {
    gpp_points* raw = nullptr;
    check(gpp_points_create(lats.data(), lons.data(), elevs, lafs, (int) lats.size(), (int) type, &raw));
    return PointsHandle(raw, gpp_points_destroy);
}
inline int& omp_threads() {
    static int n = 1;
    return n;
}
}  // namespace b200

// -------------------------------------------------------------------------------------------------------------------
inline std::string version() { return gpp_version(); }
inline bool is_valid(float value) { return !std::isnan(value) && !std::isinf(value); }   // util.cpp:16-18

// gridpp.cpp:45-68. The device path has no host thread team; the value is kept so that callers which set and read it
// back behave as before.
inline void set_omp_threads(int num) { b200::omp_threads() = num; }
inline int get_omp_threads() { return b200::omp_threads(); }
inline void initialize_omp() {}

// util.cpp:583-615
inline bool convert_coordinates(const vec& lats, const vec& lons, CoordinateType type, vec& x_coords, vec& y_coords, vec& z_coords) {
    b200::require(lats.size() == lons.size(), "Cannot convert coordinates with unequal lat and lon sizes");
    const int n = (int) lats.size();
    x_coords.resize(n);
    y_coords.resize(n);
    z_coords.resize(n);
    b200::check(gpp_convert_coordinates(lats.data(), lons.data(), n, (int) type, x_coords.data(), y_coords.data(), z_coords.data()));
    return true;
}

class Point {
public:
    // point.cpp:5-23: geodetic points are converted; for Cartesian ones x = lat, y = lon (not the x = lon, y = lat of
    // convert_coordinates) and nothing is validated
    Point(float lat, float lon, float elev = MV, float laf = MV, CoordinateType type = Geodetic) : lat(lat), lon(lon), elev(elev), laf(laf), type(type) {
        if(type == Geodetic) b200::check(gpp_convert_coordinates(&lat, &lon, 1, (int) type, &x, &y, &z));
        else {
            x = lat;
            y = lon;
            z = 0;
        }
    }
    Point(float lat, float lon, float elev, float laf, CoordinateType type, float x, float y, float z) :
            lat(lat), lon(lon), elev(elev), laf(laf), type(type), x(x), y(y), z(z) {}
    float lat, lon, elev, laf;
    CoordinateType type;
    float x, y, z;
};

// -------------------------------------------------------------------------------------------------------------------
// A set of points with its device index. KDTree, Points and (flattened) Grid are all views of one of these.
class KDTree {
public:
    KDTree(vec lats, vec lons, CoordinateType type = Geodetic) : m_lats(std::move(lats)), m_lons(std::move(lons)), m_type(type) {
        b200::require(m_lats.size() == m_lons.size(), "Cannot create KDTree with unequal lat and lon sizes");
        m_handle = b200::make_points(m_lats, m_lons, nullptr, nullptr, type);
    }
    KDTree(CoordinateType type = Geodetic) : m_type(type) { m_handle = b200::make_points(m_lats, m_lons, nullptr, nullptr, type); }

    int get_nearest_neighbour(float lat, float lon, bool include_match = true) const {
        int index = -1;
        b200::check(gpp_points_nearest_host(m_handle.get(), &lat, &lon, 1, include_match, &index));
        return index;
    }
    ivec get_neighbours(float lat, float lon, float radius, bool include_match = true) const {
        return neighbours(lat, lon, radius, include_match, nullptr);
    }
    ivec get_neighbours_with_distance(float lat, float lon, float radius, vec& distances, bool include_match = true) const {
        return neighbours(lat, lon, radius, include_match, &distances);
    }
    int get_num_neighbours(float lat, float lon, float radius, bool include_match = true) const {
        int count = 0;
        b200::check(gpp_points_neighbours_host(m_handle.get(), &lat, &lon, &radius, 1, include_match, 0, nullptr, nullptr, &count));
        return count;
    }
    ivec get_closest_neighbours(float lat, float lon, int num, bool include_match = true) const {
        b200::require(num > 0, "num must be > 0");
        ivec out((size_t) num, -1);
        b200::check(gpp_points_closest_host(m_handle.get(), &lat, &lon, 1, num, include_match, out.data()));
        while(!out.empty() && out.back() < 0) out.pop_back();
        return out;
    }

    // scalar helpers of kdtree.cpp:107-200 (host arithmetic; none of them is on the data-parallel path)
    static float deg2rad(float deg) { return (deg * M_PI / 180); }
    static float rad2deg(float rad) { return (rad * 180 / M_PI); }
    static float calc_straight_distance(float x0, float y0, float z0, float x1, float y1, float z1) {
        return std::sqrt((x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1) + (z0 - z1) * (z0 - z1));
    }
    static float calc_straight_distance(const Point& p1, const Point& p2) { return calc_straight_distance(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z); }
    static float calc_distance(float lat1, float lon1, float lat2, float lon2, CoordinateType type = Geodetic) {
        if(type == Cartesian) {
            const float dx = lon1 - lon2, dy = lat1 - lat2;
            return std::sqrt(dx * dx + dy * dy);
        }
        if(lat1 == lat2 && lon1 == lon2) return 0;
        // great-circle distance from the spherical law of cosines, in double (kdtree.cpp:121-131; like the reference, no
        // clamping of the cosine: a rounding above 1 gives NaN)
        const double a1 = deg2rad(lat1), a2 = deg2rad(lat2), o1 = deg2rad(lon1), o2 = deg2rad(lon2);
        const double c = std::cos(a1) * std::cos(o1) * std::cos(a2) * std::cos(o2) + std::cos(a1) * std::sin(o1) * std::cos(a2) * std::sin(o2) +
                         std::sin(a1) * std::sin(a2);
        return (float) (std::acos(c) * 6.378137e6);
    }
    static float calc_distance(const Point& p1, const Point& p2) {
        if(p1.type != p2.type) throw std::runtime_error("Coordinate types must be the same");
        return calc_distance(p1.lat, p1.lon, p2.lat, p2.lon, p1.type);
    }

    vec get_lats() const { return m_lats; }
    vec get_lons() const { return m_lons; }
    int size() const { return (int) m_lats.size(); }
    CoordinateType get_coordinate_type() const { return m_type; }
    vec get_x() const { return xyz(0); }
    vec get_y() const { return xyz(1); }
    vec get_z() const { return xyz(2); }

    const gpp_points* b200_handle() const { return m_handle.get(); }

protected:
    friend class Points;
    friend class Grid;
    KDTree(vec lats, vec lons, const vec& elevs, const vec& lafs, CoordinateType type) : m_lats(std::move(lats)), m_lons(std::move(lons)), m_type(type) {
        m_handle = b200::make_points(m_lats, m_lons, b200::ptr_or_null(elevs), b200::ptr_or_null(lafs), type);
    }
    ivec neighbours(float lat, float lon, float radius, bool include_match, vec* distances) const {
        int count = 0;
        b200::check(gpp_points_neighbours_host(m_handle.get(), &lat, &lon, &radius, 1, include_match, 0, nullptr, nullptr, &count));
        ivec index((size_t) count);
        if(distances) distances->assign((size_t) count, MV);
        if(count > 0)
            b200::check(gpp_points_neighbours_host(m_handle.get(), &lat, &lon, &radius, 1, include_match, count, index.data(),
                                                   distances ? distances->data() : nullptr, &count));
        return index;
    }
    Point point_at(size_t i, float elev, float laf) const {
        float x, y, z;
        b200::check(gpp_convert_coordinates(&m_lats[i], &m_lons[i], 1, (int) m_type, &x, &y, &z));
        return Point(m_lats[i], m_lons[i], elev, laf, m_type, x, y, z);
    }
    vec xyz(int which) const {
        vec out((size_t) size());
        b200::check(gpp_points_get_xyz(m_handle.get(), which == 0 ? out.data() : nullptr, which == 1 ? out.data() : nullptr,
                                       which == 2 ? out.data() : nullptr));
        return out;
    }
    vec m_lats, m_lons;
    CoordinateType m_type;
    b200::PointsHandle m_handle;
};

class Grid;

class Points {
public:
    Points() : m_tree(Geodetic) {}
    Points(vec lats, vec lons, vec elevs = vec(), vec lafs = vec(), CoordinateType type = Geodetic) : m_tree(type) {
        const size_t n = lats.size();
        b200::require(lons.size() == n, "Cannot create points with unequal lat and lon sizes");
        b200::require(elevs.size() == 0 || elevs.size() == n, "'elevs' must either be size 0 or the same size at lats/lons");
        b200::require(lafs.size() == 0 || lafs.size() == n, "'lafs' must either be size 0 or the same size at lats/lons");
        m_tree = KDTree(std::move(lats), std::move(lons), elevs, lafs, type);
        m_elevs = elevs.size() == n ? std::move(elevs) : vec(n, MV);   // points.cpp:23-30
        m_lafs = lafs.size() == n ? std::move(lafs) : vec(n, MV);
    }
    Points(KDTree tree, vec elevs = vec(), vec lafs = vec()) : Points(tree.get_lats(), tree.get_lons(), std::move(elevs), std::move(lafs), tree.get_coordinate_type()) {}

    int get_nearest_neighbour(float lat, float lon, bool include_match = true) const { return m_tree.get_nearest_neighbour(lat, lon, include_match); }
    ivec get_neighbours(float lat, float lon, float radius, bool include_match = true) const { return m_tree.get_neighbours(lat, lon, radius, include_match); }
    ivec get_neighbours_with_distance(float lat, float lon, float radius, vec& distances, bool include_match = true) const {
        return m_tree.get_neighbours_with_distance(lat, lon, radius, distances, include_match);
    }
    int get_num_neighbours(float lat, float lon, float radius, bool include_match = true) const { return m_tree.get_num_neighbours(lat, lon, radius, include_match); }
    ivec get_closest_neighbours(float lat, float lon, int num, bool include_match = true) const { return m_tree.get_closest_neighbours(lat, lon, num, include_match); }

    vec get_lats() const { return m_tree.get_lats(); }
    vec get_lons() const { return m_tree.get_lons(); }
    vec get_elevs() const { return m_elevs; }
    vec get_lafs() const { return m_lafs; }
    int size() const { return m_tree.size(); }
    CoordinateType get_coordinate_type() const { return m_tree.get_coordinate_type(); }
    Point get_point(int index) const {
        b200::require(index >= 0 && index < size(), "Point index out of range");
        return m_tree.point_at((size_t) index, m_elevs[index], m_lafs[index]);   // points.cpp:128-130: the index's coordinates
    }
    Points subset(const ivec& indices) const {
        vec lats, lons, elevs, lafs;
        for(int i : indices) {
            b200::require(i >= 0 && i < size(), "Index exceeds number of points");
            lats.push_back(m_tree.m_lats[i]);
            lons.push_back(m_tree.m_lons[i]);
            elevs.push_back(m_elevs[i]);
            lafs.push_back(m_lafs[i]);
        }
        return Points(lats, lons, elevs, lafs, get_coordinate_type());
    }
    const gpp_points* b200_handle() const { return m_tree.b200_handle(); }

private:
    friend class Grid;
    KDTree m_tree;
    vec m_elevs, m_lafs;
};

class Grid {
public:
    Grid() : m_tree(Geodetic), m_ny(0), m_nx(0) {}
    Grid(vec2 lats, vec2 lons, vec2 elevs = vec2(), vec2 lafs = vec2(), CoordinateType type = Geodetic) : m_tree(type) {
        int ly, lx;
        b200::shape_of(lats, m_ny, m_nx, "lats");
        b200::shape_of(lons, ly, lx, "lons");
        b200::require(ly == m_ny && lx == m_nx, "Cannot create grid with unequal lat and lon sizes");
        // grid.cpp:41-54: elevations / land fractions that do not have the shape of the grid are replaced by missing values
        int ey, ex, fy, fx;
        b200::shape_of(elevs, ey, ex, "elevs");
        b200::shape_of(lafs, fy, fx, "lafs");
        const size_t n = (size_t) m_ny * m_nx;
        m_elevs = (ey == m_ny && ex == m_nx) ? b200::flatten(elevs) : vec(n, MV);
        m_lafs = (fy == m_ny && fx == m_nx) ? b200::flatten(lafs) : vec(n, MV);
        m_tree = KDTree(b200::flatten(lats), b200::flatten(lons), m_elevs, m_lafs, type);
        if(n > 0) b200::check(gpp_points_set_shape(m_tree.m_handle.get(), m_ny, m_nx));
    }
    ivec get_nearest_neighbour(float lat, float lon, bool include_match = true) const {
        const int i = m_tree.get_nearest_neighbour(lat, lon, include_match);
        return i >= 0 ? yx(i) : ivec();
    }
    ivec2 get_neighbours(float lat, float lon, float radius, bool include_match = true) const { return yx(m_tree.get_neighbours(lat, lon, radius, include_match)); }
    ivec2 get_neighbours_with_distance(float lat, float lon, float radius, vec& distances, bool include_match = true) const {
        return yx(m_tree.get_neighbours_with_distance(lat, lon, radius, distances, include_match));
    }
    int get_num_neighbours(float lat, float lon, float radius, bool include_match = true) const { return m_tree.get_num_neighbours(lat, lon, radius, include_match); }
    ivec2 get_closest_neighbours(float lat, float lon, int num, bool include_match = true) const { return yx(m_tree.get_closest_neighbours(lat, lon, num, include_match)); }

    // grid.cpp:131-145: the same nodes as a flat point set; the device index is shared, not rebuilt
    Points to_points() const {
        Points p;
        p.m_tree = m_tree;
        p.m_elevs = m_elevs;
        p.m_lafs = m_lafs;
        return p;
    }
    vec2 get_lats() const { return b200::unflatten(m_tree.m_lats, m_ny, m_nx); }
    vec2 get_lons() const { return b200::unflatten(m_tree.m_lons, m_ny, m_nx); }
    vec2 get_elevs() const { return b200::unflatten(m_elevs, m_ny, m_nx); }
    vec2 get_lafs() const { return b200::unflatten(m_lafs, m_ny, m_nx); }
    ivec size() const { return ivec{m_ny, m_nx}; }
    CoordinateType get_coordinate_type() const { return m_tree.get_coordinate_type(); }
    Point get_point(int y_index, int x_index) const {
        b200::require(y_index >= 0 && y_index < m_ny && x_index >= 0 && x_index < m_nx, "Grid index out of range");
        const size_t i = (size_t) y_index * m_nx + x_index;
        return m_tree.point_at(i, m_elevs[i], m_lafs[i]);   // grid.cpp:232-235
    }
    const gpp_points* b200_handle() const { return m_tree.b200_handle(); }
    // the flattened coordinates (no copy), for callers that query every node
    const vec& flat_lats() const { return m_tree.m_lats; }
    const vec& flat_lons() const { return m_tree.m_lons; }

private:
    ivec yx(int flat) const { return ivec{flat / m_nx, flat % m_nx}; }   // grid.cpp:108-114
    ivec2 yx(const ivec& flat) const {
        ivec2 out;
        out.reserve(flat.size());
        for(int i : flat) out.push_back(yx(i));
        return out;
    }
    KDTree m_tree;
    int m_ny, m_nx;
    vec m_elevs, m_lafs;
};

// -------------------------------------------------------------------------------------------------------------------
// Structure functions. The reference's virtual hierarchy (gridpp.h:2069-2343) is kept for its names, constructors and
// clone(); what each object carries is the POD descriptor the kernels take by value, plus (for the spatially varying
// constructors) the handle of the device-resident scale fields.
class StructureFunction;
typedef std::shared_ptr<StructureFunction> StructureFunctionPtr;

class StructureFunction {
public:
    virtual ~StructureFunction() {}
    // structure.cpp:13-24 and the overrides: evaluated on the device by the code the OI kernels inline
    virtual float corr(const Point& p1, const Point& p2) const { return evaluate(p1, std::vector<Point>(1, p2), false)[0]; }
    virtual vec corr(const Point& p1, const std::vector<Point>& p2) const { return evaluate(p1, p2, false); }
    virtual float corr_background(const Point& p1, const Point& p2) const { return evaluate(p1, std::vector<Point>(1, p2), true)[0]; }
    virtual vec corr_background(const Point& p1, const std::vector<Point>& p2) const { return evaluate(p1, p2, true); }
    virtual float localization_distance(const Point& p) const {
        if(!m_field) return m_desc.term[0].loc_dist;
        float out = 0;
        b200::check(gpp_structure_field_localization_distance(m_field.get(), m_desc.term[0].type, m_desc.term[0].min_rho, p.lat, p.lon, &out));
        return out;
    }
    virtual StructureFunctionPtr clone() const = 0;
    static constexpr float default_min_rho = 0.0013f;   // structure.cpp:5

    const gpp_structure& b200_descriptor() const { return m_desc; }
    const gpp_structure_field* b200_field() const { return m_field.get(); }

protected:
    StructureFunction() { m_desc = gpp_structure(); }
    vec evaluate(const Point& p1, const std::vector<Point>& p2, bool background) const {
        gpp_structure desc = m_desc;
        if(m_field) {
            // the scales of the node nearest to p1 (structure.cpp:189-199), then the constant-scale evaluation
            float h = 0, v = 0, w = 0;
            b200::check(gpp_structure_field_lookup_host(m_field.get(), &p1.lat, &p1.lon, 1, &h, &v, &w));
            b200::check(gpp_structure_init_min_rho(&desc, m_desc.term[0].type, h, v, w, m_desc.term[0].min_rho));
        }
        const int n = (int) p2.size();
        vec a((size_t) 5 * n), b((size_t) 5 * n), out((size_t) n);
        for(int i = 0; i < n; i++) {
            const float pa[5] = {p1.x, p1.y, p1.z, p1.elev, p1.laf}, pb[5] = {p2[i].x, p2[i].y, p2[i].z, p2[i].elev, p2[i].laf};
            for(int c = 0; c < 5; c++) {
                a[5 * i + c] = pa[c];
                b[5 * i + c] = pb[c];
            }
        }
        b200::check(gpp_structure_corr_host(&desc, a.data(), b.data(), n, background, out.data()));
        return out;
    }
    // <Family>Structure(h, v, w, hmax)
    void init_constant(int family, float h, float v, float w, float hmax) { b200::check(gpp_structure_init(&m_desc, family, h, v, w, hmax)); }
    // <Family>Structure(grid, h, v, w, min_rho): structure.cpp:168-184 (Barnes) and :342, :492, :643, :790
    void init_spatial(int family, const Grid& grid, const vec2& h, const vec2& v, const vec2& w, float min_rho) {
        int hy, hx, vy, vx, wy, wx;
        b200::shape_of(h, hy, hx, "h");
        b200::shape_of(v, vy, vx, "v");
        b200::shape_of(w, wy, wx, "w");
        if(hy == 1 && hx == 1 && vy == 1 && vx == 1 && wy == 1 && wx == 1) {
            b200::check(gpp_structure_init_min_rho(&m_desc, family, h[0][0], v[0][0], w[0][0], min_rho));
            return;
        }
        const ivec shape = grid.size();
        b200::require(hy == shape[0] && hx == shape[1] && vy == hy && vx == hx && wy == hy && wx == hx, "Grid size not the same as scale size");
        m_grid = std::make_shared<Grid>(grid);   // keeps the nodes (and their device index) alive with the field
        gpp_structure_field* raw = nullptr;
        b200::check(gpp_structure_field_create(m_grid->b200_handle(), b200::flatten(h).data(), b200::flatten(v).data(), b200::flatten(w).data(), &raw));
        m_field.reset(raw, gpp_structure_field_destroy);
        m_desc = gpp_structure();
        m_desc.n_terms = 1;
        m_desc.term[0].type = family;
        m_desc.term[0].min_rho = min_rho;
        // placeholder descriptor: the scales live in the field; NaN makes the C entry points that take a plain descriptor refuse it
        m_desc.term[0].h = m_desc.term[0].v = m_desc.term[0].w = m_desc.term[0].loc_dist = MV;
    }
    friend class MultipleStructure;
    friend class CrossValidation;
    gpp_structure m_desc;
    std::shared_ptr<Grid> m_grid;
    std::shared_ptr<gpp_structure_field> m_field;
};

#define GRIDPP_B200_FAMILY(Name, FAMILY)                                                                                        \
    class Name : public StructureFunction {                                                                                     \
    public:                                                                                                                     \
        Name(float h, float v = 0, float w = 0, float hmax = MV) { init_constant(FAMILY, h, v, w, hmax); }                     \
        Name(Grid grid, vec2 h, vec2 v, vec2 w, float min_rho = StructureFunction::default_min_rho) { init_spatial(FAMILY, grid, h, v, w, min_rho); } \
        StructureFunctionPtr clone() const { return std::make_shared<Name>(*this); }                                            \
    };
GRIDPP_B200_FAMILY(BarnesStructure, GPP_STRUCT_BARNES)       // structure.cpp:143-282
GRIDPP_B200_FAMILY(SoarStructure, GPP_STRUCT_SOAR)           // structure.cpp:317-460
GRIDPP_B200_FAMILY(ToarStructure, GPP_STRUCT_TOAR)           // structure.cpp:467-610
GRIDPP_B200_FAMILY(PowerlawStructure, GPP_STRUCT_POWERLAW)   // structure.cpp:618-757
GRIDPP_B200_FAMILY(LinearStructure, GPP_STRUCT_LINEAR)       // structure.cpp:765-904
#undef GRIDPP_B200_FAMILY

class CressmanStructure : public StructureFunction {   // structure.cpp:287-312
public:
    CressmanStructure(float h, float v = 0, float w = 0) { init_constant(GPP_STRUCT_CRESSMAN, h, v, w, MV); }
    StructureFunctionPtr clone() const { return std::make_shared<CressmanStructure>(*this); }
};

class MultipleStructure : public StructureFunction {   // structure.cpp:90-138
public:
    MultipleStructure(const StructureFunction& structure_h, const StructureFunction& structure_v, const StructureFunction& structure_w) {
        if(structure_h.m_field || structure_v.m_field || structure_w.m_field)
            throw not_implemented_exception("spatially varying structure functions cannot be nested in MultipleStructure on the device");
        b200::check(gpp_structure_multiple(&m_desc, &structure_h.m_desc, &structure_v.m_desc, &structure_w.m_desc));
    }
    StructureFunctionPtr clone() const { return std::make_shared<MultipleStructure>(*this); }
};

class CrossValidation : public StructureFunction {   // structure.cpp:909-944
public:
    CrossValidation(StructureFunction& structure, float dist) {
        if(structure.m_field) throw not_implemented_exception("spatially varying structure functions cannot be nested in CrossValidation on the device");
        b200::check(gpp_structure_cross_validation(&m_desc, &structure.m_desc, dist));
    }
    StructureFunctionPtr clone() const { return std::make_shared<CrossValidation>(*this); }
};

// -------------------------------------------------------------------------------------------------------------------
// Optimal interpolation (oi.cpp:26-412). Every overload flattens to b200::run_oi, the one entry to the device.
namespace b200 {
inline void size_error(const char* what, size_t got, const char* of, size_t want) {
    std::stringstream ss;
    ss << what << " (" << got << ") " << of << " (" << want << ")";
    throw std::invalid_argument(ss.str());
}
// bvariance / bvariance_at_points may be NULL (= 1 everywhere, what optimal_interpolation() passes: oi.cpp:125-134)
inline vec run_oi(const Points& bpoints, const vec& background, const float* bvariance, const Points& obs_points, const vec& obs, const vec& obs_variance,
                  const vec& background_at_points, const float* bvariance_at_points, const StructureFunction& structure, int max_points,
                  bool allow_extrapolation, vec* analysis_variance) {
    const size_t nB = (size_t) bpoints.size();
    vec analysis(nB);
    if(analysis_variance) analysis_variance->assign(nB, MV);
    if(nB == 0) return analysis;
    float* variance = analysis_variance ? analysis_variance->data() : nullptr;
    if(const gpp_structure_field* field = structure.b200_field()) {
        const gpp_structure_term& t = structure.b200_descriptor().term[0];
        check(gpp_optimal_interpolation_spatial_host(bpoints.b200_handle(), background.data(), bvariance, obs_points.b200_handle(), obs.data(), obs_variance.data(),
                                                     background_at_points.data(), bvariance_at_points, t.type, field, t.min_rho, max_points, allow_extrapolation,
                                                     analysis.data(), variance));
    }
    else if(devices() != 1) {
        // several GPUs from this one process: the rows of the grid are split over the devices (use_devices)
        check(gpp_optimal_interpolation_multi_gpu_host(devices(), bpoints.b200_handle(), background.data(), bvariance, obs_points.b200_handle(), obs.data(),
                                                       obs_variance.data(), background_at_points.data(), bvariance_at_points, &structure.b200_descriptor(),
                                                       max_points, allow_extrapolation, analysis.data(), variance));
    }
    else {
        check(gpp_optimal_interpolation_host(bpoints.b200_handle(), background.data(), bvariance, obs_points.b200_handle(), obs.data(), obs_variance.data(),
                                             background_at_points.data(), bvariance_at_points, &structure.b200_descriptor(), max_points, allow_extrapolation,
                                             analysis.data(), variance));
    }
    return analysis;
}
inline void check_obs_sizes(const Points& obs_points, const vec& obs, const vec& obs_variance, const char* variance_name, const vec& background_at_points) {
    const size_t nS = (size_t) obs_points.size();
    if(obs.size() != nS) size_error("Observations", obs.size(), "and points size mismatch", nS);
    if(obs_variance.size() != nS) size_error(variance_name, obs_variance.size(), "and points size mismatch", nS);
    if(background_at_points.size() != nS) size_error("Background", background_at_points.size(), "and points size mismatch", nS);
}
}  // namespace b200

// oi.cpp:138-341; argument checks of :151-186
inline vec optimal_interpolation_full(const Points& bpoints, const vec& background, const vec& bvariance, const Points& obs_points, const vec& obs,
                                      const vec& obs_variance, const vec& background_at_points, const vec& bvariance_at_points,
                                      const StructureFunction& structure, int max_points, vec& analysis_variance, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    b200::require(bpoints.get_coordinate_type() == obs_points.get_coordinate_type(),
                  "Both background and observations points must be of same coordinate type (lat/lon or x/y)");
    const size_t nB = (size_t) bpoints.size();
    if(background.size() != nB) b200::size_error("Input field", background.size(), "is not the same size as the grid", nB);
    if(bvariance.size() != nB) b200::size_error("Input bvariance", bvariance.size(), "is not the same size as the grid", nB);
    b200::check_obs_sizes(obs_points, obs, obs_variance, "Observation variance", background_at_points);
    if(bvariance_at_points.size() != (size_t) obs_points.size())
        b200::size_error("Background variance", bvariance_at_points.size(), "and points size mismatch", (size_t) obs_points.size());
    return b200::run_oi(bpoints, background, bvariance.data(), obs_points, obs, obs_variance, background_at_points, bvariance_at_points.data(), structure,
                        max_points, allow_extrapolation, &analysis_variance);
}

// oi.cpp:342-412
inline vec2 optimal_interpolation_full(const Grid& bgrid, const vec2& background, const vec2& bvariance, const Points& obs_points, const vec& obs,
                                       const vec& obs_variance, const vec& background_at_points, const vec& bvariance_at_points,
                                       const StructureFunction& structure, int max_points, vec2& analysis_variance, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    const ivec shape = bgrid.size();
    int by, bx, vy, vx;
    b200::shape_of(background, by, bx, "background");
    b200::shape_of(bvariance, vy, vx, "bvariance");
    b200::require(by == shape[0] && bx == shape[1], "input field is not the same size as the grid");
    b200::require(vy == shape[0] && vx == shape[1], "input bvariance is not the same size as the grid");
    vec variance1;
    vec analysis1 = optimal_interpolation_full(bgrid.to_points(), b200::flatten(background), b200::flatten(bvariance), obs_points, obs, obs_variance,
                                               background_at_points, bvariance_at_points, structure, max_points, variance1, allow_extrapolation);
    analysis_variance = b200::unflatten(variance1, by, bx);
    return b200::unflatten(analysis1, by, bx);
}

// oi.cpp:89-136: unit background variance, the variance ratios as observation variance
inline vec optimal_interpolation(const Points& bpoints, const vec& background, const Points& obs_points, const vec& obs, const vec& variance_ratios,
                                 const vec& background_at_points, const StructureFunction& structure, int max_points, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    b200::require(bpoints.get_coordinate_type() == obs_points.get_coordinate_type(),
                  "Both background and observations points must be of same coordinate type (lat/lon or x/y)");
    if(background.size() != (size_t) bpoints.size()) b200::size_error("Input field", background.size(), "is not the same size as the grid", (size_t) bpoints.size());
    b200::check_obs_sizes(obs_points, obs, variance_ratios, "Ratios", background_at_points);
    return b200::run_oi(bpoints, background, nullptr, obs_points, obs, variance_ratios, background_at_points, nullptr, structure, max_points, allow_extrapolation,
                        nullptr);
}

// oi.cpp:26-87
inline vec2 optimal_interpolation(const Grid& bgrid, const vec2& background, const Points& obs_points, const vec& obs, const vec& variance_ratios,
                                  const vec& background_at_points, const StructureFunction& structure, int max_points, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    b200::require(bgrid.get_coordinate_type() == obs_points.get_coordinate_type(),
                  "Both background grid and observations points must be of same coordinate type (lat/lon or x/y)");
    const ivec shape = bgrid.size();
    int by, bx;
    b200::shape_of(background, by, bx, "background");
    b200::require(by == shape[0] && bx == shape[1], "input field is not the same size as the grid");
    return b200::unflatten(optimal_interpolation(bgrid.to_points(), b200::flatten(background), obs_points, obs, variance_ratios, background_at_points, structure,
                                                 max_points, allow_extrapolation), by, bx);
}

// EnSI (oi_ensi.cpp:33-568). background is nB x nE (member fastest) once flattened.
namespace b200 {
inline vec run_ensi(const Points& bpoints, const vec& background, int nE, const Points& obs_points, const vec& obs, const vec& obs_standard_deviations,
                    const vec2& background_at_points, const StructureFunction& structure, int max_points, bool allow_extrapolation) {
    require(bpoints.get_coordinate_type() == obs_points.get_coordinate_type(),
            "Both background and observations points must be of same coordinate type (lat/lon or x/y)");
    if(structure.b200_field()) throw not_implemented_exception("spatially varying structure functions in optimal_interpolation_ensi");
    const size_t nS = (size_t) obs_points.size();
    int pS, pE;
    shape_of(background_at_points, pS, pE, "background_at_points");
    if(obs.size() != nS) size_error("Observations", obs.size(), "and points size mismatch", nS);
    if(obs_standard_deviations.size() != nS) size_error("Sigmas", obs_standard_deviations.size(), "and points size mismatch", nS);
    if((size_t) pS != nS) size_error("Background", (size_t) pS, "and points size mismatch", nS);
    if(pE != nE) size_error("Ensemble members in gridded background", (size_t) nE, "is not the same as in the point background", (size_t) pE);
    vec analysis(background.size());
    if(analysis.empty()) return analysis;
    int skipped = 0;
    check(gpp_optimal_interpolation_ensi_host(bpoints.b200_handle(), background.data(), nE, obs_points.b200_handle(), obs.data(), obs_standard_deviations.data(),
                                              flatten(background_at_points).data(), &structure.b200_descriptor(), max_points, allow_extrapolation,
                                              analysis.data(), &skipped));
    if(skipped > 0)   // oi_ensi.cpp:557-566
        std::cout << "Warning: Condition number error in " << skipped << " points. Using raw values in those points." << std::endl;
    return analysis;
}
}  // namespace b200

// oi_ensi.cpp:114-568
inline vec2 optimal_interpolation_ensi(const Points& bpoints, const vec2& background, const Points& obs_points, const vec& obs,
                                       const vec& obs_standard_deviations, const vec2& background_at_points, const StructureFunction& structure,
                                       int max_points, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;   // oi_ensi.cpp:137-139
    int nB, nE;
    b200::shape_of(background, nB, nE, "background");
    b200::require(nB == bpoints.size(), "Input field is not the same size as the grid");
    return b200::unflatten(b200::run_ensi(bpoints, b200::flatten(background), nE, obs_points, obs, obs_standard_deviations, background_at_points, structure,
                                          max_points, allow_extrapolation), nB, nE);
}

// oi_ensi.cpp:33-112
inline vec3 optimal_interpolation_ensi(const Grid& bgrid, const vec3& background, const Points& obs_points, const vec& obs,
                                       const vec& obs_standard_deviations, const vec2& background_at_points, const StructureFunction& structure,
                                       int max_points, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;   // oi_ensi.cpp:49-51
    const ivec shape = bgrid.size();
    b200::require(shape[0] != 0 && shape[1] != 0, "Grid size cannot be zero");
    int ny, nx, ne;
    b200::shape_of(background, ny, nx, ne, "background");
    b200::require(ny == shape[0] && nx == shape[1], "Input field is not the same size as the grid");
    return b200::unflatten(b200::run_ensi(bgrid.to_points(), b200::flatten(background), ne, obs_points, obs, obs_standard_deviations, background_at_points,
                                          structure, max_points, allow_extrapolation), ny, nx, ne);
}

// The "multi" EnSI variants (oi_ensi_multi.cpp) and staticcorr_points (corr_points.cpp). kind: 0 ebesc, 1 ebe, 2 utem.
namespace b200 {
inline vec run_ensi_multi(int kind, const Points& bpoints, const vec& bratios, const vec& background, const vec& background_corr, int nE,
                          const Points& obs_points, const vec& pobs, const vec& pratios, const vec2& pbackground, const vec2& pbackground_corr,
                          const StructureFunction& structure, int max_points, bool allow_extrapolation) {
    require(bpoints.get_coordinate_type() == obs_points.get_coordinate_type(),
            "Both background and observations points must be of same coorindate type (lat/lon or x/y)");
    if(structure.b200_field()) throw not_implemented_exception("spatially varying structure functions in optimal_interpolation_ensi_multi");
    const size_t nS = (size_t) obs_points.size(), nB = (size_t) bpoints.size();
    if(background.size() != nB * nE) size_error("Input background field", background.size(), "is not the same size as the grid", nB * nE);
    if(bratios.size() != nB) size_error("Bratios", bratios.size(), "and grid size mismatch", nB);
    if(pobs.size() != (kind == 2 ? nS : nS * nE)) size_error("Observations", pobs.size(), "and points size mismatch", nS);
    if(pratios.size() != nS) size_error("Pratios", pratios.size(), "and points size mismatch", nS);
    int pS, pE;
    shape_of(pbackground, pS, pE, "pbackground");
    if((size_t) pS != nS || pE != nE) size_error("Background", (size_t) pS, "and points size mismatch", nS);
    vec pbc;
    if(kind != 0) {
        if(background_corr.size() != background.size()) size_error("Input background_corr field", background_corr.size(), "is not the same size as the grid", background.size());
        shape_of(pbackground_corr, pS, pE, "pbackground_corr");
        if((size_t) pS != nS || pE != nE) size_error("Background_corr", (size_t) pS, "and points size mismatch", nS);
        pbc = flatten(pbackground_corr);
    }
    vec analysis(background.size());
    if(analysis.empty()) return analysis;
    const vec pb = flatten(pbackground);
    int skipped = 0;
    if(kind == 0)
        check(gpp_optimal_interpolation_ensi_multi_ebesc_host(bpoints.b200_handle(), bratios.data(), background.data(), nE, obs_points.b200_handle(), pobs.data(),
                                                              pratios.data(), pb.data(), &structure.b200_descriptor(), max_points, allow_extrapolation, analysis.data()));
    else if(kind == 1)
        check(gpp_optimal_interpolation_ensi_multi_ebe_host(bpoints.b200_handle(), bratios.data(), background.data(), background_corr.data(), nE,
                                                            obs_points.b200_handle(), pobs.data(), pratios.data(), pb.data(), pbc.data(),
                                                            &structure.b200_descriptor(), max_points, allow_extrapolation, analysis.data()));
    else
        check(gpp_optimal_interpolation_ensi_multi_utem_host(bpoints.b200_handle(), bratios.data(), background.data(), background_corr.data(), nE,
                                                             obs_points.b200_handle(), pobs.data(), pratios.data(), pb.data(), pbc.data(),
                                                             &structure.b200_descriptor(), max_points, allow_extrapolation, analysis.data(), &skipped));
    if(skipped > 0)   // oi_ensi_multi.cpp:1300-1304
        std::cout << "Warning: Condition number error in " << skipped << " points. Using raw values in those points." << std::endl;
    return analysis;
}
inline int members_of(const vec2& a) { return a.empty() ? 0 : (int) a[0].size(); }
inline int members_of(const vec3& a) { return a.empty() || a[0].empty() ? 0 : (int) a[0][0].size(); }
}  // namespace b200

// oi_ensi_multi.cpp:329-627 (Points) and :34-135 (Grid)
inline vec2 optimal_interpolation_ensi_multi_ebe(const Points& bpoints, const vec& bratios, const vec2& background, const vec2& background_corr,
                                                 const Points& obs_points, const vec2& pobs, const vec& pratios, const vec2& pbackground,
                                                 const vec2& pbackground_corr, const StructureFunction& structure, int max_points,
                                                 bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;
    const int nE = b200::members_of(background);
    return b200::unflatten(b200::run_ensi_multi(1, bpoints, bratios, b200::flatten(background), b200::flatten(background_corr), nE, obs_points,
                                                b200::flatten(pobs), pratios, pbackground, pbackground_corr, structure, max_points, allow_extrapolation),
                           (int) background.size(), nE);
}
inline vec3 optimal_interpolation_ensi_multi_ebe(const Grid& bgrid, const vec2& bratios, const vec3& background, const vec3& background_corr,
                                                 const Points& obs_points, const vec2& pobs, const vec& pratios, const vec2& pbackground,
                                                 const vec2& pbackground_corr, const StructureFunction& structure, int max_points,
                                                 bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;
    const ivec shape = bgrid.size();
    b200::require(shape[0] != 0 && shape[1] != 0, "Grid size cannot be zero");
    int ny, nx, ne;
    b200::shape_of(background, ny, nx, ne, "background");
    b200::require(ny == shape[0] && nx == shape[1], "Input background field is not the same size as the grid");
    return b200::unflatten(b200::run_ensi_multi(1, bgrid.to_points(), b200::flatten(bratios), b200::flatten(background), b200::flatten(background_corr), ne,
                                                obs_points, b200::flatten(pobs), pratios, pbackground, pbackground_corr, structure, max_points,
                                                allow_extrapolation), ny, nx, ne);
}
// oi_ensi_multi.cpp:630-859 (Points) and :137-224 (Grid)
inline vec2 optimal_interpolation_ensi_multi_ebesc(const Points& bpoints, const vec& bratios, const vec2& background, const Points& obs_points,
                                                   const vec2& pobs, const vec& pratios, const vec2& pbackground, const StructureFunction& structure,
                                                   int max_points, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;
    const int nE = b200::members_of(background);
    return b200::unflatten(b200::run_ensi_multi(0, bpoints, bratios, b200::flatten(background), vec(), nE, obs_points, b200::flatten(pobs), pratios,
                                                pbackground, vec2(), structure, max_points, allow_extrapolation), (int) background.size(), nE);
}
inline vec3 optimal_interpolation_ensi_multi_ebesc(const Grid& bgrid, const vec2& bratios, const vec3& background, const Points& obs_points,
                                                   const vec2& pobs, const vec& pratios, const vec2& pbackground, const StructureFunction& structure,
                                                   int max_points, bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;
    const ivec shape = bgrid.size();
    b200::require(shape[0] != 0 && shape[1] != 0, "Grid size cannot be zero");
    int ny, nx, ne;
    b200::shape_of(background, ny, nx, ne, "background");
    b200::require(ny == shape[0] && nx == shape[1], "Input background field is not the same size as the grid");
    return b200::unflatten(b200::run_ensi_multi(0, bgrid.to_points(), b200::flatten(bratios), b200::flatten(background), vec(), ne, obs_points,
                                                b200::flatten(pobs), pratios, pbackground, vec2(), structure, max_points, allow_extrapolation), ny, nx, ne);
}
// oi_ensi_multi.cpp:862-1311 (Points) and :226-327 (Grid)
inline vec2 optimal_interpolation_ensi_multi_utem(const Points& bpoints, const vec& bratios, const vec2& background, const vec2& background_corr,
                                                  const Points& obs_points, const vec& pobs, const vec& pratios, const vec2& pbackground,
                                                  const vec2& pbackground_corr, const StructureFunction& structure, int max_points,
                                                  bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;
    const int nE = b200::members_of(background);
    return b200::unflatten(b200::run_ensi_multi(2, bpoints, bratios, b200::flatten(background), b200::flatten(background_corr), nE, obs_points, pobs,
                                                pratios, pbackground, pbackground_corr, structure, max_points, allow_extrapolation),
                           (int) background.size(), nE);
}
inline vec3 optimal_interpolation_ensi_multi_utem(const Grid& bgrid, const vec2& bratios, const vec3& background, const vec3& background_corr,
                                                  const Points& obs_points, const vec& pobs, const vec& pratios, const vec2& pbackground,
                                                  const vec2& pbackground_corr, const StructureFunction& structure, int max_points,
                                                  bool allow_extrapolation = true) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    if(obs_points.size() == 0) return background;
    const ivec shape = bgrid.size();
    b200::require(shape[0] != 0 && shape[1] != 0, "Grid size cannot be zero");
    int ny, nx, ne;
    b200::shape_of(background, ny, nx, ne, "background");
    b200::require(ny == shape[0] && nx == shape[1], "Input background field is not the same size as the grid");
    return b200::unflatten(b200::run_ensi_multi(2, bgrid.to_points(), b200::flatten(bratios), b200::flatten(background), b200::flatten(background_corr), ne,
                                                obs_points, pobs, pratios, pbackground, pbackground_corr, structure, max_points, allow_extrapolation),
                           ny, nx, ne);
}
// corr_points.cpp:26-131
inline vec2 staticcorr_points(const Points& points, const Points& knots, const StructureFunction& structure, int max_points) {
    b200::require(max_points >= 0, "max_points must be >= 0");
    b200::require(points.get_coordinate_type() == knots.get_coordinate_type(),
                  "Both background grid and observations points must be of same coordinate type (lat/lon or x/y)");
    if(structure.b200_field()) throw not_implemented_exception("spatially varying structure functions in staticcorr_points");
    const int nY = points.size(), nS = knots.size();
    vec out((size_t) nY * nS, 0.f);
    if(!out.empty()) b200::check(gpp_staticcorr_points_host(points.b200_handle(), knots.b200_handle(), &structure.b200_descriptor(), max_points, out.data()));
    vec2 result(nY, vec(nS, 0.f));
    for(int y = 0; y < nY; y++) std::copy(out.begin() + (size_t) y * nS, out.begin() + (size_t) (y + 1) * nS, result[y].begin());
    return result;
}

// -------------------------------------------------------------------------------------------------------------------
// Neighbourhood filters (neighbourhood.cpp:12-527)
inline vec2 neighbourhood(const vec2& input, int halfwidth, Statistic statistic) {
    b200::require(halfwidth >= 0, "Half width must be > 0");
    b200::require(statistic != Quantile, "Use neighbourhood_quantile for computing neighbourhood quantiles");
    int ny, nx;
    b200::shape_of(input, ny, nx, "input");
    if(ny == 0 || nx == 0) return vec2();
    vec out((size_t) ny * nx);
    b200::check(gpp_neighbourhood_host(b200::flatten(input).data(), ny, nx, halfwidth, (int) statistic, out.data()));
    return b200::unflatten(out, ny, nx);
}
inline vec2 neighbourhood(const vec3& input, int halfwidth, Statistic statistic) {
    b200::require(halfwidth >= 0, "Half width must be > 0");
    b200::require(statistic != Quantile, "Use neighbourhood_quantile for computing neighbourhood quantiles");
    int ny, nx, ne;
    b200::shape_of(input, ny, nx, ne, "input");
    if(ny == 0 || nx == 0 || ne == 0) return vec2();
    vec out((size_t) ny * nx);
    b200::check(gpp_neighbourhood_ens_host(b200::flatten(input).data(), ny, nx, ne, halfwidth, (int) statistic, out.data()));
    return b200::unflatten(out, ny, nx);
}

namespace b200 {
// quantile: a scalar, or a field of the input's Y x X shape (a (1, 1) field counts as the scalar)
inline vec2 quantile_fast(const vec& flat, int ny, int nx, int ne, float quantile, const vec2* quantile_field, int halfwidth, const vec& thresholds) {
    require(halfwidth >= 0, "Half width must be > 0");
    if(ny == 0 || nx == 0 || (ne == 0)) return vec2();
    vec qflat;
    if(quantile_field) {
        int qy, qx;
        shape_of(*quantile_field, qy, qx, "quantile");
        if(qy == 1 && qx == 1) quantile = (*quantile_field)[0][0];
        else {
            require(qy == ny && qx == nx, "Quantile must have the same Y, X size as input, or have size (1, 1)");
            qflat = flatten(*quantile_field);
        }
    }
    vec out((size_t) ny * nx);
    if(ne < 0)
        check(gpp_neighbourhood_quantile_fast_host(flat.data(), ny, nx, quantile, ptr_or_null(qflat), halfwidth, ptr_or_null(thresholds), (int) thresholds.size(),
                                                   out.data()));
    else
        check(gpp_neighbourhood_quantile_fast_ens_host(flat.data(), ny, nx, ne, quantile, ptr_or_null(qflat), halfwidth, ptr_or_null(thresholds),
                                                       (int) thresholds.size(), out.data()));
    return unflatten(out, ny, nx);
}
}  // namespace b200

inline vec2 neighbourhood_quantile_fast(const vec2& input, float quantile, int halfwidth, const vec& thresholds) {
    int ny, nx;
    b200::shape_of(input, ny, nx, "input");
    return b200::quantile_fast(b200::flatten(input), ny, nx, -1, quantile, nullptr, halfwidth, thresholds);
}
inline vec2 neighbourhood_quantile_fast(const vec2& input, const vec2& quantile, int halfwidth, const vec& thresholds) {
    int ny, nx;
    b200::shape_of(input, ny, nx, "input");
    return b200::quantile_fast(b200::flatten(input), ny, nx, -1, MV, &quantile, halfwidth, thresholds);
}
inline vec2 neighbourhood_quantile_fast(const vec3& input, float quantile, int halfwidth, const vec& thresholds) {
    int ny, nx, ne;
    b200::shape_of(input, ny, nx, ne, "input");
    return b200::quantile_fast(b200::flatten(input), ny, nx, ne, quantile, nullptr, halfwidth, thresholds);
}
inline vec2 neighbourhood_quantile_fast(const vec3& input, const vec2& quantile, int halfwidth, const vec& thresholds) {
    int ny, nx, ne;
    b200::shape_of(input, ny, nx, ne, "input");
    return b200::quantile_fast(b200::flatten(input), ny, nx, ne, MV, &quantile, halfwidth, thresholds);
}

// neighbourhood_brute_force / neighbourhood_quantile (exact), neighbourhood.cpp:528-539, and the deprecated names :540-552
namespace b200 {
inline vec2 brute_force(const vec& flat, int ny, int nx, int ne, int halfwidth, int statistic, float quantile) {
    require(halfwidth >= 0, "Half width must be > 0");
    if(ny == 0 || nx == 0 || ne == 0) return vec2();
    vec out((size_t) ny * nx);
    check(gpp_neighbourhood_brute_force_host(flat.data(), ny, nx, ne, halfwidth, statistic, quantile, out.data()));
    return unflatten(out, ny, nx);
}
}  // namespace b200
// Host-side helpers of the public surface (util.cpp): containers of a given shape, shape checks, point_in_rectangle
inline vec2 init_vec2(int Y, int X, float value = MV) { return vec2((size_t) Y, vec((size_t) X, value)); }
inline ivec2 init_ivec2(int Y, int X, int value) { return ivec2((size_t) Y, ivec((size_t) X, value)); }
inline vec3 init_vec3(int Y, int X, int E, float value = MV) { return vec3((size_t) Y, init_vec2(X, E, value)); }
typedef std::vector<ivec2> ivec3;
inline ivec3 init_ivec3(int Y, int X, int E, int value) { return ivec3((size_t) Y, init_ivec2(X, E, value)); }
inline bool compatible_size(const vec2& a, const vec2& b) {   // util.cpp: same number of rows, every row the same length
    if(a.size() != b.size()) return false;
    for(size_t i = 0; i < a.size(); i++)
        if(a[i].size() != b[i].size()) return false;
    return true;
}
inline bool compatible_size(const vec2& a, const vec3& b) {
    if(a.size() != b.size()) return false;
    for(size_t i = 0; i < a.size(); i++)
        if(a[i].size() != b[i].size()) return false;
    return true;
}
inline bool compatible_size(const vec3& a, const vec3& b) {
    if(a.size() != b.size()) return false;
    for(size_t i = 0; i < a.size(); i++) {
        if(a[i].size() != b[i].size()) return false;
        for(size_t j = 0; j < a[i].size(); j++)
            if(a[i][j].size() != b[i][j].size()) return false;
    }
    return true;
}
inline bool compatible_size(const Grid& grid, const vec2& v) {
    const ivec shape = grid.size();
    return (int) v.size() == shape[0] && (v.empty() || (int) v[0].size() == shape[1]);
}
inline bool compatible_size(const Grid& grid, const vec3& v) {
    const ivec shape = grid.size();
    return (int) v.size() == shape[0] && (v.empty() || (int) v[0].size() == shape[1]);
}
inline bool compatible_size(const Points& points, const vec& v) { return (int) v.size() == points.size(); }
inline bool compatible_size(const Points& points, const vec2& v) { return (int) v.size() == points.size(); }
// util.cpp:562-581: float arithmetic on the corners' lat / lon, both orientations accepted
inline bool point_in_rectangle(const Point& A, const Point& B, const Point& C, const Point& D, const Point& m) {
    auto side = [](const Point& p1, const Point& p2, const Point& q) {
        const float lon = p2.lon - p1.lon, lat = -1 * (p2.lat - p1.lat);
        const float c = -1 * (lat * p1.lon + lon * p1.lat);
        return (lat * q.lon + lon * q.lat) + c;
    };
    const float D1 = side(A, B, m), D2 = side(A, D, m), D3 = side(B, C, m), D4 = side(C, D, m);
    const bool opt1 = 0 >= D1 && 0 >= D4 && 0 <= D2 && 0 >= D3;
    const bool opt2 = 0 <= D1 && 0 <= D4 && 0 >= D2 && 0 <= D3;
    return opt1 || opt2;
}

// neighbourhood_search.cpp:7-113
inline vec2 neighbourhood_search(const vec2& array, const vec2& search_array, int halfwidth, float search_target_min, float search_target_max,
                                 float search_delta, const ivec2& apply_array = ivec2()) {
    b200::require(search_target_min <= search_target_max, "Search_target_min must be smaller than search_target_max");
    b200::require(halfwidth >= 0, "halfwidth must be positive");
    int ny, nx, sy, sx;
    b200::shape_of(array, ny, nx, "array");
    b200::shape_of(search_array, sy, sx, "search_array");
    b200::require(sy == ny && sx == nx, "search_array must either be the same size as array");
    b200::require(apply_array.size() <= 1 || ((int) apply_array.size() == ny && (int) apply_array[0].size() == nx),
                  "apply_array must either be empty or same size as array");
    ivec apply;
    if(!apply_array.empty()) {
        b200::require((int) apply_array.size() == ny, "apply_array must either be empty or same size as array");
        for(const ivec& row : apply_array) {
            b200::require((int) row.size() == nx, "apply_array must either be empty or same size as array");
            apply.insert(apply.end(), row.begin(), row.end());
        }
    }
    vec out((size_t) ny * nx);
    if(!out.empty())
        b200::check(gpp_neighbourhood_search_host(b200::flatten(array).data(), b200::flatten(search_array).data(), ny, nx, halfwidth, search_target_min,
                                                  search_target_max, search_delta, apply.empty() ? nullptr : apply.data(), out.data()));
    return b200::unflatten(out, ny, nx);
}
// calc_gradient.cpp:6-126
inline vec2 calc_gradient(const vec2& base, const vec2& values, GradientType gradient_type, int halfwidth, int min_num = 2, float min_range = MV,
                          float default_gradient = 0) {
    b200::require(halfwidth > 0, "Halwidth cannot be <= 0; must be positive integer");
    b200::require(!(is_valid(min_range) && min_range < 0), "min_range must be >= 0");
    b200::require(min_num >= 0, "num_min must be >= 0");
    b200::require(!base.empty(), "base input has no size");
    int ny, nx, vy, vx;
    b200::shape_of(base, ny, nx, "base");
    b200::shape_of(values, vy, vx, "values");
    b200::require(vy == ny && vx == nx, "base is not the same size as values");
    vec out((size_t) ny * nx);
    if(!out.empty())
        b200::check(gpp_calc_gradient_host(b200::flatten(base).data(), b200::flatten(values).data(), ny, nx, (int) gradient_type, halfwidth, min_num, min_range,
                                           default_gradient, out.data()));
    return b200::unflatten(out, ny, nx);
}
inline vec2 neighbourhood_brute_force(const vec2& input, int halfwidth, Statistic statistic) {
    int ny, nx;
    b200::shape_of(input, ny, nx, "input");
    return b200::brute_force(b200::flatten(input), ny, nx, 1, halfwidth, (int) statistic, 0.f);
}
inline vec2 neighbourhood_brute_force(const vec3& input, int halfwidth, Statistic statistic) {
    int ny, nx, ne;
    b200::shape_of(input, ny, nx, ne, "input");
    return b200::brute_force(b200::flatten(input), ny, nx, ne, halfwidth, (int) statistic, 0.f);
}
inline vec2 neighbourhood_quantile(const vec2& input, float quantile, int halfwidth) {
    int ny, nx;
    b200::shape_of(input, ny, nx, "input");
    return b200::brute_force(b200::flatten(input), ny, nx, 1, halfwidth, (int) Quantile, quantile);
}
inline vec2 neighbourhood_quantile(const vec3& input, float quantile, int halfwidth) {
    int ny, nx, ne;
    b200::shape_of(input, ny, nx, ne, "input");
    return b200::brute_force(b200::flatten(input), ny, nx, ne, halfwidth, (int) Quantile, quantile);
}
inline vec2 neighbourhood_ens(const vec3& input, int halfwidth, Statistic statistic) { return neighbourhood(input, halfwidth, statistic); }
inline vec2 neighbourhood_quantile_ens(const vec3& input, float quantile, int halfwidth) { return neighbourhood_quantile(input, quantile, halfwidth); }
inline vec2 neighbourhood_quantile_ens_fast(const vec3& input, float quantile, int halfwidth, const vec& thresholds) {
    return neighbourhood_quantile_fast(input, quantile, halfwidth, thresholds);
}

// neighbourhood.cpp:243-295
namespace b200 {
inline vec thresholds_of(const vec& flat, int num_thresholds) {
    require(num_thresholds > 0, "num_thresholds must be > 0");
    if(flat.empty()) return vec();
    vec out((size_t) num_thresholds);
    int n = 0;
    check(gpp_get_neighbourhood_thresholds_host(flat.data(), (long long) flat.size(), num_thresholds, out.data(), &n));
    out.resize((size_t) n);
    return out;
}
}  // namespace b200
inline vec get_neighbourhood_thresholds(const vec2& input, int num_thresholds) { return b200::thresholds_of(b200::flatten(input), num_thresholds); }
inline vec get_neighbourhood_thresholds(const vec3& input, int num_thresholds) { return b200::thresholds_of(b200::flatten(input), num_thresholds); }

// -------------------------------------------------------------------------------------------------------------------
// Consumers of the point index: gridding.cpp, count.cpp, distance.cpp, fill.cpp, doping.cpp. A Grid is passed as its flattened
// nodes; results come back in the shape of the output object.
namespace b200 {
inline void check_gridding(const Points& points, const vec& values, float radius, int min_num, bool with_radius) {
    require(points.size() == (int) values.size(), "Points size is not the same as values");
    if(with_radius) require(is_valid(radius) && radius >= 0, "radius must be >= 0");
    require(min_num >= 0, "min_num must be >= 0");
}
}  // namespace b200
inline vec2 gridding(const Grid& grid, const Points& points, const vec& values, float radius, int min_num, Statistic statistic) {
    b200::check_gridding(points, values, radius, min_num, true);
    vec out((size_t) grid.size()[0] * grid.size()[1]);
    b200::check(gpp_gridding_host(grid.b200_handle(), points.b200_handle(), values.data(), radius, min_num, (int) statistic, out.data()));
    return b200::unflatten(out, grid.size()[0], grid.size()[1]);
}
inline vec gridding(const Points& opoints, const Points& ipoints, const vec& values, float radius, int min_num, Statistic statistic) {
    b200::check_gridding(ipoints, values, radius, min_num, true);
    vec out((size_t) opoints.size());
    b200::check(gpp_gridding_host(opoints.b200_handle(), ipoints.b200_handle(), values.data(), radius, min_num, (int) statistic, out.data()));
    return out;
}
inline vec2 gridding_nearest(const Grid& grid, const Points& points, const vec& values, int min_num, Statistic statistic) {
    b200::check_gridding(points, values, 0, min_num, false);
    vec out((size_t) grid.size()[0] * grid.size()[1]);
    b200::check(gpp_gridding_nearest_host(grid.b200_handle(), points.b200_handle(), values.data(), min_num, (int) statistic, out.data()));
    return b200::unflatten(out, grid.size()[0], grid.size()[1]);
}
inline vec gridding_nearest(const Points& opoints, const Points& ipoints, const vec& values, int min_num, Statistic statistic) {
    b200::check_gridding(ipoints, values, 0, min_num, false);
    vec out((size_t) opoints.size());
    b200::check(gpp_gridding_nearest_host(opoints.b200_handle(), ipoints.b200_handle(), values.data(), min_num, (int) statistic, out.data()));
    return out;
}
namespace b200 {
inline vec count_flat(const gpp_points* in, const gpp_points* out_set, size_t n, float radius) {
    vec out(n);
    check(gpp_count_host(in, out_set, radius, out.data()));
    return out;
}
inline vec distance_flat(const gpp_points* in, const gpp_points* out_set, size_t n, int num, bool query_first) {
    vec out(n);
    check(gpp_distance_host(in, out_set, num, query_first ? 1 : 0, out.data()));
    return out;
}
inline size_t nodes(const Grid& g) { return (size_t) g.size()[0] * g.size()[1]; }
}  // namespace b200
inline vec count(const Grid& grid, const Points& points, float radius) { return b200::count_flat(grid.b200_handle(), points.b200_handle(), (size_t) points.size(), radius); }
inline vec2 count(const Grid& igrid, const Grid& ogrid, float radius) {
    return b200::unflatten(b200::count_flat(igrid.b200_handle(), ogrid.b200_handle(), b200::nodes(ogrid), radius), ogrid.size()[0], ogrid.size()[1]);
}
inline vec2 count(const Points& points, const Grid& grid, float radius) {
    return b200::unflatten(b200::count_flat(points.b200_handle(), grid.b200_handle(), b200::nodes(grid), radius), grid.size()[0], grid.size()[1]);
}
inline vec count(const Points& ipoints, const Points& opoints, float radius) { return b200::count_flat(ipoints.b200_handle(), opoints.b200_handle(), (size_t) opoints.size(), radius); }
inline vec distance(const Grid& grid, const Points& points, int num = 1) {
    b200::require(grid.get_coordinate_type() == points.get_coordinate_type(), "Incompatible coordinate types");
    return b200::distance_flat(grid.b200_handle(), points.b200_handle(), (size_t) points.size(), num, true);
}
inline vec2 distance(const Grid& igrid, const Grid& ogrid, int num = 1) {
    b200::require(igrid.get_coordinate_type() == ogrid.get_coordinate_type(), "Incompatible coordinate types");
    return b200::unflatten(b200::distance_flat(igrid.b200_handle(), ogrid.b200_handle(), b200::nodes(ogrid), num, false), ogrid.size()[0], ogrid.size()[1]);
}
inline vec2 distance(const Points& points, const Grid& grid, int num = 1) {
    b200::require(points.get_coordinate_type() == grid.get_coordinate_type(), "Incompatible coordinate types");
    return b200::unflatten(b200::distance_flat(points.b200_handle(), grid.b200_handle(), b200::nodes(grid), num, false), grid.size()[0], grid.size()[1]);
}
inline vec distance(const Points& ipoints, const Points& opoints, int num = 1) {
    b200::require(ipoints.get_coordinate_type() == opoints.get_coordinate_type(), "Incompatible coordinate types");
    return b200::distance_flat(ipoints.b200_handle(), opoints.b200_handle(), (size_t) opoints.size(), num, true);
}
inline vec2 fill(const Grid& igrid, const vec2& input, const Points& points, const vec& radii, float value, bool outside) {
    int ny, nx;
    b200::shape_of(input, ny, nx, "input");
    b200::require(ny == igrid.size()[0] && nx == igrid.size()[1], "Grid size is not the same as values");
    b200::require(points.size() == (int) radii.size(), "Points size is not the same as radii size");
    for(float r : radii) b200::require(!(r < 0), "All radius sizes must be 0 or greater");
    vec out((size_t) ny * nx);
    b200::check(gpp_fill_host(igrid.b200_handle(), b200::flatten(input).data(), points.b200_handle(), radii.data(), value, outside ? 1 : 0, out.data()));
    return b200::unflatten(out, ny, nx);
}
inline vec2 fill_missing(const vec2& values) {
    int ny, nx;
    b200::shape_of(values, ny, nx, "values");
    vec out((size_t) ny * nx);
    if(!out.empty()) b200::check(gpp_fill_missing_host(b200::flatten(values).data(), ny, nx, out.data()));
    return b200::unflatten(out, ny, nx);
}
inline vec2 doping_square(const Grid& igrid, const vec2& background, const Points& points, const vec& observations, const ivec& halfwidth,
                          float max_elev_diff = MV) {
    int ny, nx;
    b200::shape_of(background, ny, nx, "background");
    b200::require(ny == igrid.size()[0] && nx == igrid.size()[1], "Grid size is not the same as observations");
    b200::require(points.size() == (int) observations.size(), "Points size is not the same as observations size");
    b200::require(points.size() == (int) halfwidth.size(), "Points size is not the same as halfwidth size");
    b200::require(!(is_valid(max_elev_diff) && max_elev_diff < 0), "max_elev_diff must be greater than or equal to 0");
    vec out((size_t) ny * nx);
    if(!out.empty())
        b200::check(gpp_doping_square_host(igrid.b200_handle(), b200::flatten(background).data(), points.b200_handle(), observations.data(), halfwidth.data(),
                                           max_elev_diff, out.data()));
    return b200::unflatten(out, ny, nx);
}
inline vec2 doping_circle(const Grid& igrid, const vec2& background, const Points& points, const vec& observations, const vec& radii,
                          float max_elev_diff = MV) {
    int ny, nx;
    b200::shape_of(background, ny, nx, "background");
    b200::require(ny == igrid.size()[0] && nx == igrid.size()[1], "Grid size is not the same as observations");
    b200::require(points.size() == (int) observations.size(), "Points size is not the same as observations size");
    b200::require(points.size() == (int) radii.size(), "Points size is not the same as radii size");
    b200::require(!(is_valid(max_elev_diff) && max_elev_diff < 0), "max_elev_diff must be greater than or equal to 0");
    vec out((size_t) ny * nx);
    if(!out.empty())
        b200::check(gpp_doping_circle_host(igrid.b200_handle(), b200::flatten(background).data(), points.b200_handle(), observations.data(), radii.data(),
                                           max_elev_diff, out.data()));
    return b200::unflatten(out, ny, nx);
}

// -------------------------------------------------------------------------------------------------------------------
// Row statistics (util.cpp:19-215,377-431; gridpp.cpp:11-44): calc_statistic, calc_quantile, interpolate, get_statistic.
inline Statistic get_statistic(std::string name) {
    static const std::pair<const char*, Statistic> names[] = {{"mean", Mean}, {"min", Min}, {"max", Max}, {"median", Median}, {"quantile", Quantile},
                                                              {"std", Std}, {"sum", Sum}, {"count", Count}, {"randomchoice", RandomChoice}};
    for(const auto& n : names)
        if(name == n.first) return n.second;
    return Unknown;
}
namespace b200 {
// rows of different lengths cannot go through the flat entry point at once: one call per row then
inline bool ragged(const vec2& array) {
    for(const vec& row : array)
        if(row.size() != array[0].size()) return true;
    return false;
}
}  // namespace b200
inline float calc_statistic(const vec& array, Statistic statistic) {
    float out = MV;
    b200::check(gpp_calc_statistic_host(array.data(), 1, (int) array.size(), (int) statistic, &out));
    return out;
}
inline vec calc_statistic(const vec2& array, Statistic statistic) {
    vec out(array.size());
    if(array.empty()) return out;
    if(b200::ragged(array)) {
        for(size_t n = 0; n < array.size(); n++) out[n] = calc_statistic(array[n], statistic);
        return out;
    }
    b200::check(gpp_calc_statistic_host(b200::flatten(array).data(), (long long) array.size(), (int) array[0].size(), (int) statistic, out.data()));
    return out;
}
inline float calc_quantile(const vec& array, float quantile) {
    float out = MV;
    b200::check(gpp_calc_quantile_host(array.data(), 1, (int) array.size(), quantile, nullptr, &out));
    return out;
}
inline vec calc_quantile(const vec2& array, float quantile = MV) {
    vec out(array.size());
    if(array.empty()) return out;
    if(b200::ragged(array)) {
        for(size_t n = 0; n < array.size(); n++) out[n] = calc_quantile(array[n], quantile);
        return out;
    }
    b200::check(gpp_calc_quantile_host(b200::flatten(array).data(), (long long) array.size(), (int) array[0].size(), quantile, nullptr, out.data()));
    return out;
}
inline vec2 calc_quantile(const vec3& array, const vec2& quantile) {
    int ny, nx, nt, qy, qx;
    b200::shape_of(array, ny, nx, nt, "array");
    b200::shape_of(quantile, qy, qx, "quantile");
    b200::require(ny == qy && (ny == 0 || nx == qx), "Dimension mismatch between array and quantile");
    if(ny == 0 || nx == 0) return vec2();
    if(nt == 0) return vec2((size_t) ny, vec((size_t) nx, MV));
    vec out((size_t) ny * nx);
    b200::check(gpp_calc_quantile_host(b200::flatten(array).data(), (long long) ny * nx, nt, MV, b200::flatten(quantile).data(), out.data()));
    return b200::unflatten(out, ny, nx);
}
inline vec interpolate(const vec& x, const vec& iX, const vec& iY) {
    b200::require(iX.size() == iY.size(), "Dimension mismatch. Cannot interpolate.");
    vec out(x.size());
    if(x.empty()) return out;
    b200::check(gpp_interpolate_host(x.data(), (long long) x.size(), iX.data(), iY.data(), (int) iX.size(), out.data()));
    return out;
}
inline float interpolate(float x, const vec& iX, const vec& iY) {
    b200::require(iX.size() == iY.size(), "Dimension mismatch. Cannot interpolate.");
    float out = MV;
    b200::check(gpp_interpolate_host(&x, 1, iX.data(), iY.data(), (int) iX.size(), &out));
    return out;
}

// -------------------------------------------------------------------------------------------------------------------
// nearest (nearest.cpp:7-222): all eight overloads gather n_fields flattened input fields at the nearest input node of
// every output location.
namespace b200 {
inline vec gather_nearest(const gpp_points* in, const vec& qlats, const vec& qlons, const vec& fields, int n_fields) {
    vec out((size_t) n_fields * qlats.size());
    if(out.empty()) return out;
    check(gpp_nearest_host(in, qlats.data(), qlons.data(), (int) qlats.size(), fields.data(), n_fields, out.data()));
    return out;
}
}  // namespace b200

inline vec2 nearest(const Grid& igrid, const Grid& ogrid, const vec2& ivalues) {
    int vy, vx;
    b200::shape_of(ivalues, vy, vx, "ivalues");
    b200::require(vy == igrid.size()[0] && vx == igrid.size()[1], "Grid size is not the same as values");
    return b200::unflatten(b200::gather_nearest(igrid.b200_handle(), ogrid.flat_lats(), ogrid.flat_lons(), b200::flatten(ivalues), 1), ogrid.size()[0], ogrid.size()[1]);
}
inline vec3 nearest(const Grid& igrid, const Grid& ogrid, const vec3& ivalues) {
    int nt, vy, vx;
    b200::shape_of(ivalues, nt, vy, vx, "ivalues");
    b200::require(nt == 0 || (vy == igrid.size()[0] && vx == igrid.size()[1]), "Grid size is not the same as values");
    return b200::unflatten(b200::gather_nearest(igrid.b200_handle(), ogrid.flat_lats(), ogrid.flat_lons(), b200::flatten(ivalues), nt), nt, ogrid.size()[0],
                           ogrid.size()[1]);
}
inline vec nearest(const Grid& igrid, const Points& opoints, const vec2& ivalues) {
    int vy, vx;
    b200::shape_of(ivalues, vy, vx, "ivalues");
    b200::require(vy == igrid.size()[0] && vx == igrid.size()[1], "Grid size is not the same as values");
    return b200::gather_nearest(igrid.b200_handle(), opoints.get_lats(), opoints.get_lons(), b200::flatten(ivalues), 1);
}
inline vec2 nearest(const Grid& igrid, const Points& opoints, const vec3& ivalues) {
    int nt, vy, vx;
    b200::shape_of(ivalues, nt, vy, vx, "ivalues");
    b200::require(nt == 0 || (vy == igrid.size()[0] && vx == igrid.size()[1]), "Grid size is not the same as values");
    return b200::unflatten(b200::gather_nearest(igrid.b200_handle(), opoints.get_lats(), opoints.get_lons(), b200::flatten(ivalues), nt), nt, opoints.size());
}
inline vec nearest(const Points& ipoints, const Points& opoints, const vec& ivalues) {
    b200::require((int) ivalues.size() == ipoints.size(), "Points size is not the same as values");
    return b200::gather_nearest(ipoints.b200_handle(), opoints.get_lats(), opoints.get_lons(), ivalues, 1);
}
inline vec2 nearest(const Points& ipoints, const Points& opoints, const vec2& ivalues) {
    int nt, n;
    b200::shape_of(ivalues, nt, n, "ivalues");
    b200::require(nt == 0 || n == ipoints.size(), "Points size is not the same as values");
    return b200::unflatten(b200::gather_nearest(ipoints.b200_handle(), opoints.get_lats(), opoints.get_lons(), b200::flatten(ivalues), nt), nt, opoints.size());
}
inline vec2 nearest(const Points& ipoints, const Grid& ogrid, const vec& ivalues) {
    b200::require((int) ivalues.size() == ipoints.size(), "Points size is not the same as values");
    return b200::unflatten(b200::gather_nearest(ipoints.b200_handle(), ogrid.flat_lats(), ogrid.flat_lons(), ivalues, 1), ogrid.size()[0], ogrid.size()[1]);
}
inline vec3 nearest(const Points& ipoints, const Grid& ogrid, const vec2& ivalues) {
    int nt, n;
    b200::shape_of(ivalues, nt, n, "ivalues");
    b200::require(nt == 0 || n == ipoints.size(), "Points size is not the same as values");
    return b200::unflatten(b200::gather_nearest(ipoints.b200_handle(), ogrid.flat_lats(), ogrid.flat_lons(), b200::flatten(ivalues), nt), nt, ogrid.size()[0],
                           ogrid.size()[1]);
}

}  // namespace gridpp
#endif
