/* gridpp_oracle.c -- plain-C restatement of the metno/gridpp hot path (CPU).
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity oracle for the CUDA path: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may call it; the product (gridpp_b200/) never does.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function below (a) against the known-answer values
 * held by the reference's own tests (tests/test_optimal_interpolation.py, test_barnes_structure.py,
 * test_kdtree.py, test_neighbourhood.py, test_neighbourhood_quantile_fast.py, test_interpolate.py ...),
 * (b) against fixtures in tests/golden/ generated from the reference sources themselves (compiled unmodified
 * into oracle/_ref/libgridpp_ref.so), and (c) directly against that library when it is present.
 *
 * Each function cites the reference file:line it follows (paths relative to the metno/gridpp tree).
 * Floating-point evaluation order and precision (float vs double) mirror the reference statement by
 * statement; compile with -ffp-contract=off and without -ffast-math.
 *
 * Two things are NOT taken from the reference because it leaves them unspecified:
 *   - order of radius-query results (Boost rtree traversal order, kdtree.cpp:53-60): ascending index here;
 *   - ties: nearest neighbour (kdtree.cpp:90-93) -> lowest index; observation selection by rho
 *     (unstable std::sort on rho only, oi.cpp:19-23,266) -> higher rho first, then lower index.
 * Dense linear algebra that the reference delegates to Armadillo/LAPACK (inv oi.cpp:315, oi_ensi.cpp:398;
 * rcond oi_ensi.cpp:386; eig_sym oi_ensi.cpp:401) is restated as Gauss-Jordan with partial pivoting and cyclic
 * Jacobi in fp64 (Armadillo >= 6.5, CMakeLists.txt:21, is not vendored in the reference tree).
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Same PODs as include/gridpp_b200.h (restated here so that this file has no dependency on the product). */
typedef struct { int type; float h, v, w; float min_rho; float loc_dist; } orc_term;
typedef struct { int n_terms; orc_term term[3]; int has_cv; float cv_dist; } orc_structure;
enum { BARNES = 0, CRESSMAN = 1, SOAR = 2, TOAR = 3, POWERLAW = 4, LINEAR = 5 };
enum { GEODETIC = 0, CARTESIAN = 1 };
enum { MEAN = 0, MIN = 10, MEDIAN = 20, MAX = 30, QUANTILE = 40, STD = 50, VARIANCE = 60, SUM = 70, COUNT = 80 };

static const double RADIUS_EARTH = 6.378137e6; /* gridpp.h:56 */
static const float DEFAULT_MIN_RHO = 0.0013f;  /* structure.cpp:5 */

static _Thread_local char g_error[512];
#define FAIL(code, ...) do { snprintf(g_error, sizeof(g_error), __VA_ARGS__); return code; } while(0)

const char* orc_last_error(void) { return g_error; }
const char* orc_version(void) { return "oracle-of-0.8.0.dev1"; }
void orc_set_omp_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}
int orc_get_omp_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 0;
#endif
}
static double now_seconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* util.cpp:16-18 */
static int is_valid(float v) { return !isnan(v) && !isinf(v); }

/* ------------------------------------------------------------------ coordinates ---------------------- */
/* util.cpp:583-615 (+ is_valid_lat/lon :617-624) */
static int convert_one(float lat, float lon, int type, float* x, float* y, float* z) {
    int ok_lat = type == CARTESIAN ? is_valid(lat) : (is_valid(lat) && (lat >= -90.001) && (lat <= 90.001));
    if(!ok_lat || !is_valid(lon)) FAIL(1, "Invalid coords: %g,%g", lat, lon);
    if(type == CARTESIAN) {
        *x = lon;
        *y = lat;
        *z = 0;
    }
    else {
        double lonr = M_PI / 180 * lon;
        double latr = M_PI / 180 * lat;
        *x = cos(latr) * cos(lonr) * RADIUS_EARTH;
        *y = cos(latr) * sin(lonr) * RADIUS_EARTH;
        *z = sin(latr) * RADIUS_EARTH;
    }
    return 0;
}
int orc_convert_coordinates(const float* lats, const float* lons, int n, int type, float* x, float* y, float* z) {
    for(int i = 0; i < n; i++) {
        int rc = convert_one(lats[i], lons[i], type, &x[i], &y[i], &z[i]);
        if(rc) return rc;
    }
    return 0;
}
/* kdtree.cpp:192-194 */
static float straight_distance(float x0, float y0, float z0, float x1, float y1, float z1) {
    return sqrtf((x0 - x1) * (x0 - x1) + (y0 - y1) * (y0 - y1) + (z0 - z1) * (z0 - z1));
}
/* kdtree.cpp:195-197 */
static float deg2rad(float deg) { return (deg * M_PI / 180); }
/* kdtree.cpp:107-135 */
float orc_calc_distance(float lat1, float lon1, float lat2, float lon2, int type) {
    if(type == CARTESIAN) {
        float dx = lon1 - lon2;
        float dy = lat1 - lat2;
        return sqrtf(dx * dx + dy * dy);
    }
    if(lat1 == lat2 && lon1 == lon2) return 0;
    double lat1r = deg2rad(lat1), lat2r = deg2rad(lat2), lon1r = deg2rad(lon1), lon2r = deg2rad(lon2);
    double radiusEarth = 6.378137e6;
    double ratio = cos(lat1r) * cos(lon1r) * cos(lat2r) * cos(lon2r) + cos(lat1r) * sin(lon1r) * cos(lat2r) * sin(lon2r)
                   + sin(lat1r) * sin(lat2r);
    double dist = acos(ratio) * radiusEarth;
    return (float) dist;
}

/* ------------------------------------------------------------------ points --------------------------- */
typedef struct {
    int n, type;
    float *x, *y, *z, *elev, *laf;
} pts_t;

static void pts_free(pts_t* p) {
    free(p->x); free(p->y); free(p->z); free(p->elev); free(p->laf);
    memset(p, 0, sizeof(*p));
}
/* points.cpp:9-31 + kdtree.cpp:6-16 */
static int pts_make(pts_t* p, const float* lats, const float* lons, const float* elevs, const float* lafs, int n, int type) {
    memset(p, 0, sizeof(*p));
    p->n = n;
    p->type = type;
    size_t bytes = sizeof(float) * (size_t) (n > 0 ? n : 1);
    p->x = malloc(bytes); p->y = malloc(bytes); p->z = malloc(bytes); p->elev = malloc(bytes); p->laf = malloc(bytes);
    for(int i = 0; i < n; i++) {
        int rc = convert_one(lats[i], lons[i], type, &p->x[i], &p->y[i], &p->z[i]);
        if(rc) { pts_free(p); return rc; }
        p->elev[i] = elevs ? elevs[i] : NAN; /* points.cpp:23-30 */
        p->laf[i] = lafs ? lafs[i] : NAN;
    }
    return 0;
}

/* A uniform bucket grid over the points, only to make the oracle usable at 1e4 observations; the predicate
 * applied to each candidate is exactly the reference's. */
typedef struct {
    float lo[3];
    double inv[3];
    int n[3];
    int* start;
    int* order;
} cells_t;
static int cell_of(const cells_t* c, int d, float v) {
    double t = floor(((double) v - (double) c->lo[d]) * c->inv[d]);
    if(!(t > 0)) return 0;
    if(t > c->n[d] - 1) return c->n[d] - 1;
    return (int) t;
}
static void cells_free(cells_t* c) { free(c->start); free(c->order); memset(c, 0, sizeof(*c)); }
static void cells_build(cells_t* c, const pts_t* p, double edge) {
    const float* co[3] = {p->x, p->y, p->z};
    float hi[3];
    long long total = 1;
    for(int d = 0; d < 3; d++) {
        c->lo[d] = INFINITY; hi[d] = -INFINITY;
        for(int i = 0; i < p->n; i++) { if(co[d][i] < c->lo[d]) c->lo[d] = co[d][i]; if(co[d][i] > hi[d]) hi[d] = co[d][i]; }
        double ext = p->n > 0 ? (double) hi[d] - (double) c->lo[d] : 0;
        int n = 1;
        if(ext > 0 && edge > 0) { double t = ceil(ext / edge); n = t > 256 ? 256 : (t < 1 ? 1 : (int) t); }
        c->n[d] = n;
        c->inv[d] = ext > 0 ? n / ext : 0;
        total *= n;
    }
    c->start = calloc((size_t) total + 1, sizeof(int));
    c->order = malloc(sizeof(int) * (size_t) (p->n > 0 ? p->n : 1));
    int* id = malloc(sizeof(int) * (size_t) (p->n > 0 ? p->n : 1));
    for(int i = 0; i < p->n; i++) {
        id[i] = (cell_of(c, 2, p->z[i]) * c->n[1] + cell_of(c, 1, p->y[i])) * c->n[0] + cell_of(c, 0, p->x[i]);
        c->start[id[i] + 1]++;
    }
    for(long long i = 0; i < total; i++) c->start[i + 1] += c->start[i];
    int* fill = malloc(sizeof(int) * (size_t) total);
    memcpy(fill, c->start, sizeof(int) * (size_t) total);
    for(int i = 0; i < p->n; i++) c->order[fill[id[i]]++] = i;
    free(fill);
    free(id);
}
static int cmp_int(const void* a, const void* b) { int x = *(const int*) a, y = *(const int*) b; return (x > y) - (x < y); }

/* kdtree.cpp:39-62 with the predicates :247-260: STRICTLY inside the box [q-r, q+r]^3 (Boost within()),
 * straight distance <= r, and distance > 0 unless include_match. Returns the count; indices ascending. */
static int radius_query(const pts_t* p, const cells_t* c, float x, float y, float z, float radius, int include_match,
                        int** out, int* cap) {
    float lo[3] = {x - radius, y - radius, z - radius};
    float hi[3] = {x + radius, y + radius, z + radius};
    int n = 0;
    if(!(lo[0] < hi[0]) || !(lo[1] < hi[1]) || !(lo[2] < hi[2])) return 0;
    int c0[3], c1[3];
    for(int d = 0; d < 3; d++) { c0[d] = cell_of(c, d, lo[d]); c1[d] = cell_of(c, d, hi[d]); }
    for(int cz = c0[2]; cz <= c1[2]; cz++)
        for(int cy = c0[1]; cy <= c1[1]; cy++) {
            int base = (cz * c->n[1] + cy) * c->n[0];
            for(int k = c->start[base + c0[0]]; k < c->start[base + c1[0] + 1]; k++) {
                int i = c->order[k];
                float px = p->x[i], py = p->y[i], pz = p->z[i];
                if(!(px > lo[0] && px < hi[0] && py > lo[1] && py < hi[1] && pz > lo[2] && pz < hi[2])) continue;
                float dist = straight_distance(px, py, pz, x, y, z);
                int ok = include_match ? (dist <= radius) : (dist <= radius && dist > 0);
                if(!ok) continue;
                if(n == *cap) { *cap = *cap ? 2 * *cap : 64; *out = realloc(*out, sizeof(int) * (size_t) *cap); }
                (*out)[n++] = i;
            }
        }
    qsort(*out, (size_t) n, sizeof(int), cmp_int);
    return n;
}

/* kdtree.cpp:82-106 (+ is_not_equal :262-270): k nearest by squared distance in double (Boost computes the
 * comparable distance of float points in double); ties -> lowest index. Brute force. */
static int closest_query(const pts_t* p, float x, float y, float z, int num, int include_match, int* out) {
    int found = 0;
    double* bd = malloc(sizeof(double) * (size_t) (num > 0 ? num : 1));
    for(int i = 0; i < p->n; i++) {
        if(!include_match && !(x != p->x[i] || y != p->y[i] || z != p->z[i])) continue;
        double dx = (double) x - (double) p->x[i], dy = (double) y - (double) p->y[i], dz = (double) z - (double) p->z[i];
        double d2 = dx * dx + dy * dy + dz * dz;
        /* insertion into the sorted (d2, index) list */
        int pos = found;
        while(pos > 0 && d2 < bd[pos - 1]) pos--;
        if(pos >= num) continue;
        int last = found < num ? found : num - 1;
        for(int k = last; k > pos; k--) { bd[k] = bd[k - 1]; out[k] = out[k - 1]; }
        bd[pos] = d2;
        out[pos] = i;
        if(found < num) found++;
    }
    free(bd);
    return found;
}

int orc_points_nearest(const float* lats, const float* lons, int n, int type, const float* qlats, const float* qlons, int nq,
                       int include_match, int* out_index, double* seconds) {
    pts_t p;
    int rc = pts_make(&p, lats, lons, NULL, NULL, n, type);
    if(rc) return rc;
    double t0 = now_seconds();
    int err = 0;
    #pragma omp parallel for
    for(int q = 0; q < nq; q++) {
        float x, y, z;
        if(convert_one(qlats[q], qlons[q], type, &x, &y, &z)) { err = 1; continue; }
        int idx;
        out_index[q] = closest_query(&p, x, y, z, 1, include_match, &idx) ? idx : -1; /* points.cpp:56-62 */
    }
    if(seconds) *seconds = now_seconds() - t0;
    pts_free(&p);
    if(err) FAIL(1, "Invalid coords in query");
    return 0;
}
int orc_points_neighbours(const float* lats, const float* lons, int n, int type, const float* qlats, const float* qlons,
                          const float* radii, int nq, int include_match, int capacity, int* out_index, float* out_dist,
                          int* out_count) {
    pts_t p;
    cells_t c;
    int rc = pts_make(&p, lats, lons, NULL, NULL, n, type);
    if(rc) return rc;
    cells_build(&c, &p, 0);
    int* buf = NULL;
    int cap = 0;
    for(int q = 0; q < nq; q++) {
        float x, y, z;
        rc = convert_one(qlats[q], qlons[q], type, &x, &y, &z);
        if(rc) break;
        int m = radius_query(&p, &c, x, y, z, radii[q], include_match, &buf, &cap);
        out_count[q] = m;
        for(int i = 0; i < m && i < capacity; i++) {
            if(out_index) out_index[(size_t) q * capacity + i] = buf[i];
            /* kdtree.cpp:23-34 */
            if(out_dist) out_dist[(size_t) q * capacity + i] = straight_distance(x, y, z, p.x[buf[i]], p.y[buf[i]], p.z[buf[i]]);
        }
    }
    free(buf);
    cells_free(&c);
    pts_free(&p);
    return rc;
}
int orc_points_neighbours_raw(const float* lats, const float* lons, int n, int type, float qlat, float qlon, float radius,
                              int include_match, int capacity, int* out_index, int* out_count) {
    float r = radius;
    return orc_points_neighbours(lats, lons, n, type, &qlat, &qlon, &r, 1, include_match, capacity, out_index, NULL, out_count);
}
int orc_points_closest(const float* lats, const float* lons, int n, int type, const float* qlats, const float* qlons, int nq,
                       int num, int include_match, int* out_index) {
    pts_t p;
    int rc = pts_make(&p, lats, lons, NULL, NULL, n, type);
    if(rc) return rc;
    for(int q = 0; q < nq && !rc; q++) {
        float x, y, z;
        rc = convert_one(qlats[q], qlons[q], type, &x, &y, &z);
        if(rc) break;
        int* o = out_index + (size_t) q * num;
        int m = closest_query(&p, x, y, z, num, include_match, o);
        for(int i = m; i < num; i++) o[i] = -1;
    }
    pts_free(&p);
    return rc;
}
/* nearest.cpp:124-144,177-197 (and the multi-field forms :145-175,198-222): value of the nearest input point */
int orc_nearest(const float* ilats, const float* ilons, int n_in, int type, const float* qlats, const float* qlons, int nq,
                const float* ivalues, int n_fields, float* out, double* seconds) {
    int* idx = malloc(sizeof(int) * (size_t) (nq > 0 ? nq : 1));
    double t = 0;
    int rc = orc_points_nearest(ilats, ilons, n_in, type, qlats, qlons, nq, 1, idx, &t);
    if(!rc)
        for(int f = 0; f < n_fields; f++)
            for(int q = 0; q < nq; q++) out[(size_t) f * nq + q] = idx[q] >= 0 ? ivalues[(size_t) f * n_in + idx[q]] : NAN;
    if(seconds) *seconds = t;
    free(idx);
    return rc;
}

/* ------------------------------------------------------------------ structure functions -------------- */
/* structure.cpp:26-34 */
static float barnes_rho(float dist, float length) {
    if(!is_valid(length) || length == 0) return 1;
    if(!is_valid(dist)) return 0;
    float v = dist / length;
    return exp(-0.5 * v * v);
}
/* structure.cpp:35-44 */
static float cressman_rho(float dist, float length) {
    if(!is_valid(length) || length == 0) return 1;
    if(!is_valid(dist)) return 0;
    if(dist >= length) return 0;
    return (length * length - dist * dist) / (length * length + dist * dist);
}
/* structure.cpp:46-54; exp(float) resolves to the float overload (see oracle/shims/boost/geometry.hpp) */
static float soar_rho(float dist, float length) {
    if(!is_valid(length) || length == 0) return 1;
    if(!is_valid(dist)) return 0;
    float v = dist / length;
    return (1 + v) * expf(-v);
}
/* structure.cpp:56-64 */
static float toar_rho(float dist, float length) {
    if(!is_valid(length) || length == 0) return 1;
    if(!is_valid(dist)) return 0;
    float v = dist / length;
    return (1 + v + (v * v) / 3) * expf(-v);
}
/* structure.cpp:66-74 */
static float powerlaw_rho(float dist, float length) {
    if(!is_valid(length) || length == 0) return 1;
    if(!is_valid(dist)) return 0;
    float v = dist / length;
    return 1 / (1 + 0.5 * v * v);
}
/* structure.cpp:76-87 */
static float linear_rho(float diff, float min_corr) {
    if(!is_valid(min_corr) || min_corr < 0) return 1;
    if(!is_valid(diff)) return 0;
    float absdiff = fabsf(diff);
    if(absdiff > 1) absdiff = 1;
    return (1 - (1 - min_corr) * absdiff);
}
static float term_rho(int type, float dist, float length) {
    switch(type) {
        case BARNES: return barnes_rho(dist, length);
        case CRESSMAN: return cressman_rho(dist, length);
        case SOAR: return soar_rho(dist, length);
        case TOAR: return toar_rho(dist, length);
        case POWERLAW: return powerlaw_rho(dist, length);
        default: return linear_rho(dist, length);
    }
}
/* localization_distance(h): Barnes structure.cpp:280-282, Soar :454-459, Toar :603-609, Powerlaw :755-757,
 * Linear :902-904, Cressman = base class m_localization_distance = h (structure.cpp:288, :88-89). */
static float term_loc_dist(const orc_term* t) {
    float m = t->min_rho, h = t->h;
    switch(t->type) {
        case BARNES: return sqrtf(-2 * logf(m)) * h;
        case CRESSMAN: return h;
        case SOAR: { float l = logf(m); return (-l + logf(-l)) * h; }
        case TOAR: { float l = logf(m); float ll = logf(-logf(m)); return (-l + ll + 0.5 * ll) * h; }
        case POWERLAW: return sqrtf(2 * (1 - m) / m) * h;
        default: return 0;
    }
}
/* constructors with hmax: structure.cpp:143-167 (Barnes), :317-340 (Soar), :467-490 (Toar), :618-641 (Powerlaw),
 * :765-788 (Linear), :287-297 (Cressman) */
static int term_init(orc_term* t, int type, float h, float v, float w, float hmax) {
    if(type != CRESSMAN && is_valid(hmax) && hmax < 0) FAIL(1, "hmax must be >= 0");
    if(type == CRESSMAN) {
        if(!is_valid(h) || h < 0) FAIL(1, "Structure function initizlied with invalid localization distance");
    }
    else if(!is_valid(h) || h < 0) FAIL(1, "h must be >= 0");
    if(!is_valid(v) || v < 0) FAIL(1, "v must be >= 0");
    if(!is_valid(w) || w < 0) FAIL(1, "w must be >= 0");
    t->type = type; t->h = h; t->v = v; t->w = w;
    t->min_rho = DEFAULT_MIN_RHO;
    if(is_valid(hmax)) {
        switch(type) {
            case BARNES: t->min_rho = exp(pow(hmax / h, 2) / -2); break;
            case SOAR: t->min_rho = (1 + hmax / h) * expf(-hmax / h); break;
            case TOAR: t->min_rho = (1 + hmax / h + pow(hmax / h, 2) / 3) * expf(-hmax / h); break;
            case POWERLAW: t->min_rho = 1 / (1 + 0.5 * pow(hmax / h, 2)); break;
            default: break;
        }
    }
    t->loc_dist = term_loc_dist(t);
    return 0;
}
/* m_min_rho as set by the constructors (not observable through the reference's public API; its effect, the
 * localization distance, is) */
int orc_structure_min_rho(int type, float h, float v, float w, float hmax, float* min_rho) {
    orc_term t;
    int rc = term_init(&t, type, h, v, w, hmax);
    if(rc) return rc;
    *min_rho = t.min_rho;
    return 0;
}
int orc_structure_describe(int type, float h, float v, float w, float hmax, float* loc_dist) {
    orc_term t;
    int rc = term_init(&t, type, h, v, w, hmax);
    if(rc) return rc;
    *loc_dist = t.loc_dist;
    return 0;
}

typedef struct { float x, y, z, elev, laf; } pt_t;

/* <Family>Structure::corr, non-spatial branch: Barnes structure.cpp:214-228, Soar :388-402, Toar :538-552,
 * Powerlaw :689-703, Linear :836-850; Cressman :298-309 (no localization test). */
static float term_corr(const orc_term* t, pt_t p1, pt_t p2) {
    float hdist = straight_distance(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z);
    if(t->type != CRESSMAN && hdist > term_loc_dist(t)) return 0;
    float rho = term_rho(t->type, hdist, t->h);
    if(is_valid(p1.elev) && is_valid(p2.elev)) {
        float vdist = p1.elev - p2.elev;
        rho *= term_rho(t->type, vdist, t->v);
    }
    if(is_valid(p1.laf) && is_valid(p2.laf)) {
        float lafdist = p1.laf - p2.laf;
        rho *= term_rho(t->type, lafdist, t->w);
    }
    return rho;
}
/* StructureFunction::corr for the descriptor: plain term, or MultipleStructure::corr structure.cpp:98-112 */
static float structure_corr(const orc_structure* s, pt_t p1, pt_t p2) {
    if(s->n_terms != 3) return term_corr(&s->term[0], p1, p2);
    pt_t p2_h = {p2.x, p2.y, p2.z, p1.elev, p1.laf};
    pt_t p2_v = {p1.x, p1.y, p1.z, p2.elev, p1.laf};
    pt_t p2_w = {p1.x, p1.y, p1.z, p1.elev, p2.laf};
    float corr_h = term_corr(&s->term[0], p1, p2_h);
    float corr_v = term_corr(&s->term[1], p1, p2_v);
    float corr_w = term_corr(&s->term[2], p1, p2_w);
    return corr_h * corr_v * corr_w;
}
/* corr_background: base class structure.cpp:20-25; CrossValidation structure.cpp:919-935 */
static float structure_corr_background(const orc_structure* s, pt_t p1, pt_t p2) {
    if(s->has_cv) {
        float hdist = straight_distance(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z);
        if(is_valid(s->cv_dist) && hdist <= s->cv_dist) return 0;
    }
    return structure_corr(s, p1, p2);
}
/* localization_distance: term; MultipleStructure structure.cpp:95-97; CrossValidation :942-944 */
static float structure_loc_dist(const orc_structure* s) { return term_loc_dist(&s->term[0]); }

int orc_structure_corr(const orc_structure* s, const float* p1, const float* p2, int n, int background, float* out) {
    for(int i = 0; i < n; i++) {
        pt_t a = {p1[5 * i], p1[5 * i + 1], p1[5 * i + 2], p1[5 * i + 3], p1[5 * i + 4]};
        pt_t b = {p2[5 * i], p2[5 * i + 1], p2[5 * i + 2], p2[5 * i + 3], p2[5 * i + 4]};
        out[i] = background ? structure_corr_background(s, a, b) : structure_corr(s, a, b);
    }
    return 0;
}
int orc_structure_localization_distance(const orc_structure* s, float* out) {
    *out = structure_loc_dist(s);
    return 0;
}

/* ------------------------------------------------------------------ dense fp64 helpers --------------- */
/* column-major n x n. Gauss-Jordan with partial pivoting (stands for arma::inv). Returns 0 when singular. */
static int mat_inv(const double* a, double* out, int n) {
    double* w = malloc(sizeof(double) * (size_t) n * n);
    memcpy(w, a, sizeof(double) * (size_t) n * n);
    for(int i = 0; i < n * n; i++) out[i] = 0;
    for(int i = 0; i < n; i++) out[i + i * n] = 1;
    for(int c = 0; c < n; c++) {
        int p = c;
        double best = fabs(w[c + c * n]);
        for(int i = c + 1; i < n; i++)
            if(fabs(w[i + c * n]) > best) { best = fabs(w[i + c * n]); p = i; }
        if(best == 0 || best != best) { free(w); return 0; }
        if(p != c)
            for(int j = 0; j < n; j++) {
                double t = w[c + j * n]; w[c + j * n] = w[p + j * n]; w[p + j * n] = t;
                t = out[c + j * n]; out[c + j * n] = out[p + j * n]; out[p + j * n] = t;
            }
        double d = 1.0 / w[c + c * n];
        for(int j = 0; j < n; j++) { w[c + j * n] *= d; out[c + j * n] *= d; }
        for(int i = 0; i < n; i++) {
            if(i == c) continue;
            double f = w[i + c * n];
            if(f == 0) continue;
            for(int j = 0; j < n; j++) { w[i + j * n] -= f * w[c + j * n]; out[i + j * n] -= f * out[c + j * n]; }
        }
    }
    free(w);
    return 1;
}
static double mat_norm1(const double* a, int n) {
    double best = 0;
    for(int j = 0; j < n; j++) {
        double s = 0;
        for(int i = 0; i < n; i++) s += fabs(a[i + j * n]);
        if(s != s) return s;
        if(s > best) best = s;
    }
    return best;
}
/* cyclic Jacobi (stands for arma::eig_sym); eigenvalues ascending in val, vectors in the columns of vec */
static int mat_eig_sym(const double* a_in, double* val, double* vec, int n) {
    double* a = malloc(sizeof(double) * (size_t) n * n);
    double* v = malloc(sizeof(double) * (size_t) n * n);
    memcpy(a, a_in, sizeof(double) * (size_t) n * n);
    for(int i = 0; i < n * n; i++) v[i] = 0;
    for(int i = 0; i < n; i++) v[i + i * n] = 1;
    int ok = 1;
    for(int sweep = 0; sweep < 100 && ok; sweep++) {
        double off = 0, diag = 0;
        for(int i = 0; i < n; i++) {
            diag += a[i + i * n] * a[i + i * n];
            for(int j = i + 1; j < n; j++) off += a[i + j * n] * a[i + j * n];
        }
        if(off != off) { ok = 0; break; }
        if(off <= 1e-32 * diag || off == 0) break;
        for(int p = 0; p + 1 < n; p++)
            for(int q = p + 1; q < n; q++) {
                double apq = a[p + q * n];
                if(apq == 0) continue;
                double theta = (a[q + q * n] - a[p + p * n]) / (2 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for(int k = 0; k < n; k++) {
                    double akp = a[k + p * n], akq = a[k + q * n];
                    a[k + p * n] = c * akp - s * akq;
                    a[k + q * n] = s * akp + c * akq;
                }
                for(int k = 0; k < n; k++) {
                    double apk = a[p + k * n], aqk = a[q + k * n];
                    a[p + k * n] = c * apk - s * aqk;
                    a[q + k * n] = s * apk + c * aqk;
                }
                for(int k = 0; k < n; k++) {
                    double vkp = v[k + p * n], vkq = v[k + q * n];
                    v[k + p * n] = c * vkp - s * vkq;
                    v[k + q * n] = s * vkp + c * vkq;
                }
            }
    }
    /* selection sort of the eigenvalues, ascending */
    int* order = malloc(sizeof(int) * (size_t) (n > 0 ? n : 1));
    for(int i = 0; i < n; i++) order[i] = i;
    for(int i = 0; i < n; i++)
        for(int j = i + 1; j < n; j++)
            if(a[order[j] + order[j] * n] < a[order[i] + order[i] * n]) { int t = order[i]; order[i] = order[j]; order[j] = t; }
    for(int j = 0; j < n; j++) {
        val[j] = a[order[j] + order[j] * n];
        for(int i = 0; i < n; i++) vec[i + j * n] = v[i + order[j] * n];
    }
    free(order); free(a); free(v);
    return ok;
}

/* ------------------------------------------------------------------ optimal interpolation ------------ */
typedef struct { float rho; int pos; int index; } cand_t;
/* selection order: higher rho first, then lower observation index (the reference sorts on rho only with an
 * unstable sort, oi.cpp:19-23,266 -- ties there are unspecified) */
static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = a; const cand_t* y = b;
    if(x->rho != y->rho) return x->rho > y->rho ? -1 : 1;
    return (x->index > y->index) - (x->index < y->index);
}

/* oi.cpp:138-341. bvariance / bvariance_at_points may be NULL (= 1, as set by oi.cpp:123-132).
 * Spatially varying structure functions (<Family>Structure(Grid, h, v, w, min_rho), structure.cpp:168-184 and the Soar
 * :342, Toar :492, Powerlaw :643, Linear :790 siblings): bterm / oterm hold, per background point and per observation,
 * the term with the scales of the point's nearest node of the scale grid. corr(p1, p2) then uses the term of p1
 * (structure.cpp:188-212), so the background correlations use the grid point's scales and row i of P uses those of
 * observation i; the localization radius is that of the grid point (oi.cpp:229 -> structure.cpp:271-279). */
static int oi_core(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB,
                   const float* background, const float* bvariance, const float* plats, const float* plons,
                   const float* pelevs, const float* plafs, int nS, int type, const float* pobs,
                   const float* obs_variance, const float* pbackground, const float* bvariance_at_points,
                   const orc_structure* s, const orc_term* bterm, const orc_term* oterm, int max_points, int allow_extrapolation,
                   float* analysis, float* analysis_variance, double* seconds) {
    if(max_points < 0) FAIL(1, "max_points must be >= 0"); /* oi.cpp:152 */
    pts_t bp, op;
    int rc = pts_make(&bp, blats, blons, belevs, blafs, nB, type);
    if(rc) return rc;
    rc = pts_make(&op, plats, plons, pelevs, plafs, nS, type);
    if(rc) { pts_free(&bp); return rc; }
    /* oi.cpp:201-203: output := background, analysis_variance := bvariance */
    for(int y = 0; y < nB; y++) {
        analysis[y] = background[y];
        if(analysis_variance) analysis_variance[y] = bvariance ? bvariance[y] : 1.0f;
    }
    if(nS == 0) { pts_free(&bp); pts_free(&op); if(seconds) *seconds = 0; return 0; } /* oi.cpp:189-190 */

    float* pratios = malloc(sizeof(float) * (size_t) nS);
    for(int i = 0; i < nS; i++) pratios[i] = obs_variance[i] / (bvariance_at_points ? bvariance_at_points[i] : 1.0f); /* :192-195 */
    cells_t cells;
    float R = bterm ? 0.f : structure_loc_dist(s);
    if(bterm)
        for(int y = 0; y < nB; y++)
            if(bterm[y].loc_dist > R) R = bterm[y].loc_dist;
    cells_build(&cells, &op, R > 0 ? 0.5 * R : 0);

    double t0 = now_seconds();
    #pragma omp parallel
    {
        int* nb = NULL;
        int nb_cap = 0;
        cand_t* cand = NULL;
        int cand_cap = 0;
        #pragma omp for schedule(dynamic, 64)
        for(int y = 0; y < nB; y++) {
            if(!is_valid(background[y])) continue; /* oi.cpp:223 */
            pt_t p1 = {bp.x[y], bp.y[y], bp.z[y], bp.elev[y], bp.laf[y]};
            float localizationRadius = bterm ? bterm[y].loc_dist : structure_loc_dist(s); /* oi.cpp:229 */
            int n0 = radius_query(&op, &cells, p1.x, p1.y, p1.z, localizationRadius, 1, &nb, &nb_cap); /* :233 */
            if(n0 == 0) continue;
            if(n0 > cand_cap) { cand_cap = 2 * n0; cand = realloc(cand, sizeof(cand_t) * (size_t) cand_cap); }
            int nc = 0;
            for(int i = 0; i < n0; i++) { /* oi.cpp:244-258 */
                int index = nb[i];
                pt_t p2 = {op.x[index], op.y[index], op.z[index], op.elev[index], op.laf[index]};
                float rho = bterm ? term_corr(&bterm[y], p1, p2) : structure_corr_background(s, p1, p2);
                if(is_valid(pobs[index]) && is_valid(pbackground[index]) && rho > 0) {
                    cand[nc].rho = rho; cand[nc].pos = i; cand[nc].index = index; nc++;
                }
            }
            int lS = nc;
            if(max_points > 0 && nc > max_points) { /* oi.cpp:262-273 */
                qsort(cand, (size_t) nc, sizeof(cand_t), cmp_cand);
                lS = max_points;
            }
            if(lS == 0) continue; /* oi.cpp:284-287 */
            double* lP = malloc(sizeof(double) * (size_t) lS * lS);
            double* lInv = malloc(sizeof(double) * (size_t) lS * lS);
            double* lG = malloc(sizeof(double) * (size_t) lS);
            double* d = malloc(sizeof(double) * (size_t) lS);
            double* lGSR = malloc(sizeof(double) * (size_t) lS);
            for(int i = 0; i < lS; i++) { /* oi.cpp:298-314 */
                int index = cand[i].index;
                d[i] = (double) pobs[index] - (double) pbackground[index];
                lG[i] = cand[i].rho;
                pt_t pi = {op.x[index], op.y[index], op.z[index], op.elev[index], op.laf[index]};
                for(int j = 0; j < lS; j++) {
                    int index_j = cand[j].index;
                    pt_t pj = {op.x[index_j], op.y[index_j], op.z[index_j], op.elev[index_j], op.laf[index_j]};
                    lP[i + j * lS] = oterm ? term_corr(&oterm[index], pi, pj) : structure_corr(s, pi, pj);
                }
                lP[i + i * lS] += (double) pratios[index];
            }
            if(mat_inv(lP, lInv, lS)) {
                /* oi.cpp:315-317: lGSR = lG * inv(lP + lR); dx = lGSR * (lObs - lY) */
                double dx = 0, a = 0;
                for(int j = 0; j < lS; j++) {
                    double acc = 0;
                    for(int k = 0; k < lS; k++) acc += lG[k] * lInv[k + j * lS];
                    lGSR[j] = acc;
                }
                for(int j = 0; j < lS; j++) dx += lGSR[j] * d[j];
                for(int j = 0; j < lS; j++) a += lGSR[j] * lG[j];
                float increment = dx;
                if(!allow_extrapolation) { /* oi.cpp:318-334 */
                    double mx = d[0], mn = d[0];
                    for(int j = 1; j < lS; j++) { if(d[j] > mx) mx = d[j]; if(d[j] < mn) mn = d[j]; }
                    float maxInc = mx, minInc = mn;
                    if(maxInc > 0 && increment > maxInc) increment = maxInc;
                    else if(maxInc < 0 && increment > 0) increment = maxInc;
                    else if(minInc < 0 && increment < minInc) increment = minInc;
                    else if(minInc > 0 && increment < 0) increment = minInc;
                }
                analysis[y] = background[y] + increment; /* oi.cpp:335 */
                if(analysis_variance) analysis_variance[y] = (bvariance ? bvariance[y] : 1.0f) * (1 - a); /* :336-337 */
            }
            free(lP); free(lInv); free(lG); free(d); free(lGSR);
        }
        free(nb);
        free(cand);
    }
    if(seconds) *seconds = now_seconds() - t0;
    cells_free(&cells);
    free(pratios);
    pts_free(&bp);
    pts_free(&op);
    return 0;
}

int orc_optimal_interpolation(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB,
                              const float* background, const float* bvariance, const float* plats, const float* plons,
                              const float* pelevs, const float* plafs, int nS, int type, const float* pobs,
                              const float* obs_variance, const float* pbackground, const float* bvariance_at_points,
                              const orc_structure* s, int max_points, int allow_extrapolation, float* analysis,
                              float* analysis_variance, double* seconds) {
    return oi_core(blats, blons, belevs, blafs, nB, background, bvariance, plats, plons, pelevs, plafs, nS, type, pobs, obs_variance,
                   pbackground, bvariance_at_points, s, NULL, NULL, max_points, allow_extrapolation, analysis, analysis_variance, seconds);
}

/* The term a spatially varying structure function uses at (lat, lon): the scales of the nearest node of its grid
 * (structure.cpp:189-199; Grid::get_nearest_neighbour, grid.cpp:76-82; ties resolve to the lowest node index here),
 * with localization_distance(h) of that node (structure.cpp:280-282 and siblings). */
static int spatial_terms(orc_term* out, const float* lats, const float* lons, int n, int type, const pts_t* nodes, int stype,
                         const float* h, const float* v, const float* w, float min_rho) {
    int err = 0;
    #pragma omp parallel for
    for(int i = 0; i < n; i++) {
        float x, y, z;
        int node = 0;
        if(convert_one(lats[i], lons[i], type, &x, &y, &z) || !closest_query(nodes, x, y, z, 1, 1, &node)) { err = 1; continue; }
        out[i].type = stype;
        out[i].h = h[node]; out[i].v = v[node]; out[i].w = w[node];
        out[i].min_rho = min_rho;
        out[i].loc_dist = term_loc_dist(&out[i]);
    }
    if(err) FAIL(1, "Invalid coords or empty scale grid");
    return 0;
}

/* optimal_interpolation_full(Points...) (oi.cpp:138-341) with unit background variances and a spatially varying
 * structure function: h, v, w on the gny x gnx nodes (glats, glons) of its grid. Same flat signature as
 * ref_optimal_interpolation_spatial in ref_capi.cpp. */
int orc_optimal_interpolation_spatial(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB,
                                      const float* background, const float* plats, const float* plons, const float* pelevs,
                                      const float* plafs, int nS, int type, const float* pobs, const float* obs_variance,
                                      const float* pbackground, int stype, const float* glats, const float* glons, int gny, int gnx,
                                      const float* h, const float* v, const float* w, float min_rho, int max_points,
                                      int allow_extrapolation, float* analysis, float* analysis_variance) {
    if(stype == CRESSMAN || stype < BARNES || stype > LINEAR) FAIL(1, "structure function type has no spatially varying form");
    pts_t nodes;
    int rc = pts_make(&nodes, glats, glons, NULL, NULL, gny * gnx, type);
    if(rc) return rc;
    orc_term* bterm = malloc(sizeof(orc_term) * (size_t) (nB > 0 ? nB : 1));
    orc_term* oterm = malloc(sizeof(orc_term) * (size_t) (nS > 0 ? nS : 1));
    rc = spatial_terms(bterm, blats, blons, nB, type, &nodes, stype, h, v, w, min_rho);
    if(!rc) rc = spatial_terms(oterm, plats, plons, nS, type, &nodes, stype, h, v, w, min_rho);
    if(!rc)
        rc = oi_core(blats, blons, belevs, blafs, nB, background, NULL, plats, plons, pelevs, plafs, nS, type, pobs, obs_variance, pbackground,
                     NULL, NULL, bterm, oterm, max_points, allow_extrapolation, analysis, analysis_variance, NULL);
    free(bterm);
    free(oterm);
    pts_free(&nodes);
    return rc;
}

/* util.cpp:19-43 (Mean branch of calc_statistic): float accumulation */
static float mean_valid(const float* a, int n) {
    float total = 0;
    int count = 0;
    for(int i = 0; i < n; i++)
        if(is_valid(a[i])) { total += a[i]; count++; }
    return count > 0 ? total / count : NAN;
}

/* oi_ensi.cpp:114-568. background nB x nE, pbackground nS x nE (member fastest). *num_skipped (may be NULL)
 * counts the grid points left at their raw values because rcond(Pinv) <= 0 (oi_ensi.cpp:386-390). */
int orc_optimal_interpolation_ensi(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB,
                                   const float* background, int nEns, const float* plats, const float* plons,
                                   const float* pelevs, const float* plafs, int nS, int type, const float* pobs,
                                   const float* psigmas, const float* pbackground, const orc_structure* s, int max_points,
                                   int allow_extrapolation, float* analysis, double* seconds) {
    if(max_points < 0) FAIL(1, "max_points must be >= 0");
    pts_t bp, op;
    int rc = pts_make(&bp, blats, blons, belevs, blafs, nB, type);
    if(rc) return rc;
    rc = pts_make(&op, plats, plons, pelevs, plafs, nS, type);
    if(rc) { pts_free(&bp); return rc; }
    memcpy(analysis, background, sizeof(float) * (size_t) nB * nEns); /* oi_ensi.cpp:148 */
    if(nS == 0 || nB == 0 || nEns == 0) { pts_free(&bp); pts_free(&op); if(seconds) *seconds = 0; return 0; }

    /* oi_ensi.cpp:163-178: remove the ensemble mean at the observation points */
    float* gY = malloc(sizeof(float) * (size_t) nS * nEns);
    float* gYhat = malloc(sizeof(float) * (size_t) nS);
    memcpy(gY, pbackground, sizeof(float) * (size_t) nS * nEns);
    for(int i = 0; i < nS; i++) {
        float mean = mean_valid(gY + (size_t) i * nEns, nEns);
        for(int e = 0; e < nEns; e++) {
            float value = gY[(size_t) i * nEns + e];
            if(is_valid(value) && is_valid(mean)) gY[(size_t) i * nEns + e] -= mean;
        }
        gYhat[i] = mean;
    }
    /* oi_ensi.cpp:187-201: members without any invalid value anywhere */
    int* validEns = malloc(sizeof(int) * (size_t) nEns);
    int nValidEns = 0;
    for(int e = 0; e < nEns; e++) {
        int numInvalid = 0;
        for(int y = 0; y < nB; y++)
            if(!is_valid(background[(size_t) y * nEns + e])) numInvalid++;
        if(numInvalid == 0) validEns[nValidEns++] = e;
    }
    cells_t cells;
    float R0 = structure_loc_dist(s);
    cells_build(&cells, &op, R0 > 0 ? 0.5 * R0 : 0);
    const int E = nValidEns;
    double t0 = now_seconds();
    /* the reference loop is serial (oi_ensi.cpp:203-207); iterations are independent, so run them in parallel */
    #pragma omp parallel
    {
        int* nb = NULL;
        int nb_cap = 0;
        cand_t* cand = NULL;
        int cand_cap = 0;
        #pragma omp for schedule(dynamic, 16)
        for(int y = 0; y < nB; y++) {
            pt_t p1 = {bp.x[y], bp.y[y], bp.z[y], bp.elev[y], bp.laf[y]};
            float localizationRadius = structure_loc_dist(s);
            int n0 = radius_query(&op, &cells, p1.x, p1.y, p1.z, localizationRadius, 1, &nb, &nb_cap);
            if(n0 == 0) continue;
            if(n0 > cand_cap) { cand_cap = 2 * n0; cand = realloc(cand, sizeof(cand_t) * (size_t) cand_cap); }
            int nc = 0;
            for(int i = 0; i < n0; i++) { /* oi_ensi.cpp:226-240: only pobs validity is tested here */
                int index = nb[i];
                pt_t p2 = {op.x[index], op.y[index], op.z[index], op.elev[index], op.laf[index]};
                float rho = structure_corr_background(s, p1, p2);
                if(is_valid(pobs[index]) && rho > 0) { cand[nc].rho = rho; cand[nc].pos = i; cand[nc].index = index; nc++; }
            }
            int lS = nc;
            if(max_points > 0 && nc > max_points) { qsort(cand, (size_t) nc, sizeof(cand_t), cmp_cand); lS = max_points; }
            if(lS == 0 || E == 0) continue;

            /* lY (lS x E), Rinv diag, C = lY' * Rinv (E x lS), Pinv = C*lY + diag*I : oi_ensi.cpp:282-385 */
            double* lY = malloc(sizeof(double) * (size_t) lS * E);
            double* Cm = malloc(sizeof(double) * (size_t) E * lS);
            double* Pinv = malloc(sizeof(double) * (size_t) E * E);
            double* P = malloc(sizeof(double) * (size_t) E * E);
            double* dd = malloc(sizeof(double) * (size_t) lS);
            for(int i = 0; i < lS; i++) {
                int index = cand[i].index;
                for(int e = 0; e < E; e++) lY[i + (size_t) e * lS] = gY[(size_t) index * nEns + validEns[e]];
                double rinv = (double) cand[i].rho / (psigmas[index] * psigmas[index]); /* oi_ensi.cpp:300 */
                for(int e = 0; e < E; e++) Cm[e + (size_t) i * E] = lY[i + (size_t) e * lS] * rinv;
                dd[i] = (double) pobs[index] - (double) gYhat[index];
            }
            float diag = 1 / 1.0f * (E - 1); /* oi_ensi.cpp:383, delta = 1 */
            for(int a = 0; a < E; a++)
                for(int b = 0; b < E; b++) {
                    double acc = 0;
                    for(int i = 0; i < lS; i++) acc += Cm[a + (size_t) i * E] * lY[i + (size_t) b * lS];
                    Pinv[a + (size_t) b * E] = acc + (a == b ? (double) diag : 0.0);
                }
            int ok = mat_inv(Pinv, P, E);
            float cond = 0;
            if(ok) { double nn = mat_norm1(Pinv, E) * mat_norm1(P, E); cond = (nn != nn || nn == 0) ? 0 : 1.0 / nn; }
            if(ok && cond > 0) { /* oi_ensi.cpp:386-390 */
                double* S = malloc(sizeof(double) * (size_t) E * E);
                double* val = malloc(sizeof(double) * (size_t) E);
                double* vec = malloc(sizeof(double) * (size_t) E * E);
                double* W = malloc(sizeof(double) * (size_t) E * E);
                double* w = malloc(sizeof(double) * (size_t) E);
                double* X = malloc(sizeof(double) * (size_t) E);
                for(int i = 0; i < E * E; i++) S[i] = (double) (E - 1) * P[i]; /* oi_ensi.cpp:401 */
                mat_eig_sym(S, val, vec, E);
                for(int a = 0; a < E; a++) /* W = V sqrt(L) V' : oi_ensi.cpp:419-421 */
                    for(int b = 0; b < E; b++) {
                        double acc = 0;
                        for(int k = 0; k < E; k++) acc += vec[a + (size_t) k * E] * sqrt(val[k]) * vec[b + (size_t) k * E];
                        W[a + (size_t) b * E] = acc;
                    }
                for(int a = 0; a < E; a++) { /* w = P*C*(lObs - lYhat) : oi_ensi.cpp:428-437 */
                    double acc = 0;
                    for(int i = 0; i < lS; i++) {
                        double pc = 0;
                        for(int b = 0; b < E; b++) pc += P[a + (size_t) b * E] * Cm[b + (size_t) i * E];
                        acc += pc * dd[i];
                    }
                    w[a] = acc;
                }
                for(int a = 0; a < E; a++)
                    for(int b = 0; b < E; b++) W[a + (size_t) b * E] += w[a]; /* oi_ensi.cpp:440-444 */
                /* oi_ensi.cpp:447-462 */
                float total = 0;
                int count = 0;
                for(int e = 0; e < E; e++) {
                    float value = background[(size_t) y * nEns + validEns[e]];
                    X[e] = 0;
                    if(is_valid(value)) { X[e] = value; total += value; count++; }
                }
                float ensMean = total / count;
                for(int e = 0; e < E; e++) X[e] -= ensMean;
                for(int e = 0; e < E; e++) { /* oi_ensi.cpp:506-554 */
                    float tot = 0;
                    for(int k = 0; k < E; k++) tot += X[k] * W[k + (size_t) e * E];
                    float currIncrement = tot;
                    if(!allow_extrapolation) {
                        /* lY[e] is a LINEAR (column-major) index into the lS x E matrix: oi_ensi.cpp:523-524 */
                        double lYe = lY[e];
                        double mx = -INFINITY, mn = INFINITY;
                        for(int i = 0; i < lS; i++) {
                            double v = (double) pobs[cand[i].index] - (lYe + (double) gYhat[cand[i].index]);
                            if(v > mx) mx = v;
                            if(v < mn) mn = v;
                        }
                        float maxInc = mx, minInc = mn;
                        float memberIncrement = currIncrement - X[e];
                        if(maxInc > 0 && memberIncrement > maxInc) currIncrement = maxInc + X[e];
                        else if(maxInc < 0 && memberIncrement > 0) currIncrement = 0 + X[e];
                        else if(minInc < 0 && memberIncrement < minInc) currIncrement = minInc + X[e];
                        else if(minInc > 0 && memberIncrement < 0) currIncrement = 0 + X[e];
                    }
                    analysis[(size_t) y * nEns + validEns[e]] = ensMean + currIncrement;
                }
                free(S); free(val); free(vec); free(W); free(w); free(X);
            }
            free(lY); free(Cm); free(Pinv); free(P); free(dd);
        }
        free(nb);
        free(cand);
    }
    if(seconds) *seconds = now_seconds() - t0;
    cells_free(&cells);
    free(gY); free(gYhat); free(validEns);
    pts_free(&bp);
    pts_free(&op);
    return 0;
}

/* ------------------------------------------------------------------ statistics ----------------------- */
int orc_calc_statistic(const float* array, int n, int statistic, float* out);
static int cmp_float(const void* a, const void* b) { float x = *(const float*) a, y = *(const float*) b; return (x > y) - (x < y); }
/* util.cpp:111-178 */
int orc_calc_quantile(const float* array, int T, float quantile, float* out) {
    if(quantile < 0 || quantile > 1) FAIL(1, "calc_quantile: Quantile must be between 0 and 1 inclusive");
    *out = NAN;
    if(!is_valid(quantile) || T == 0) return 0;
    if(quantile == 0 || quantile == 1) {
        float best = NAN;
        for(int i = 0; i < T; i++) {
            float val = array[i];
            if(!is_valid(val)) continue;
            else if(!is_valid(best)) best = val;
            else if(quantile == 0 ? val < best : val > best) best = val;
        }
        *out = best;
        return 0;
    }
    float* clean = malloc(sizeof(float) * (size_t) T);
    int N = 0;
    for(int i = 0; i < T; i++)
        if(is_valid(array[i])) clean[N++] = array[i];
    if(N > 0) {
        qsort(clean, (size_t) N, sizeof(float), cmp_float);
        int lowerIndex = floor(quantile * (N - 1));
        int upperIndex = ceil(quantile * (N - 1));
        float lowerQuantile = (float) lowerIndex / (N - 1);
        float upperQuantile = (float) upperIndex / (N - 1);
        float lowerValue = clean[lowerIndex];
        float upperValue = clean[upperIndex];
        if(lowerIndex == upperIndex) *out = lowerValue;
        else {
            float f = (quantile - lowerQuantile) / (upperQuantile - lowerQuantile);
            *out = lowerValue + (upperValue - lowerValue) * f;
        }
    }
    free(clean);
    return 0;
}
/* util.cpp:19-110 for Mean/Sum/Count/Min/Median/Max/Std/Variance */
int orc_calc_statistic(const float* array, int n, int statistic, float* out) {
    float value = NAN;
    if(statistic == MEAN || statistic == SUM || statistic == COUNT) {
        float total = 0;
        int count = 0;
        for(int i = 0; i < n; i++)
            if(is_valid(array[i])) { total += array[i]; count++; }
        if(statistic == COUNT) value = count;
        else if(count > 0) value = statistic == MEAN ? total / count : total;
    }
    else if(statistic == STD || statistic == VARIANCE) {
        float total = 0, total2 = 0, K = NAN;
        int count = 0;
        for(int i = 0; i < n; i++)
            if(is_valid(array[i])) {
                if(!is_valid(K)) K = array[i];
                total += array[i] - K;
                total2 += (array[i] - K) * (array[i] - K);
                count++;
            }
        if(count > 0) {
            float mean = total / count, mean2 = total2 / count;
            float var = mean2 - mean * mean;
            if(var < 0) var = 0;
            value = statistic == STD ? sqrtf(var) : var;
        }
    }
    else {
        float q = statistic == MIN ? 0 : (statistic == MEDIAN ? 0.5f : (statistic == MAX ? 1 : NAN));
        if(isnan(q)) FAIL(2, "Internal error. Cannot compute statistic");
        return orc_calc_quantile(array, n, q, out);
    }
    *out = value;
    return 0;
}

/* util.cpp:339-376. `int index = gridpp::MV` in the reference is a NaN->int conversion; -1 stands for it here
 * (unreachable from interpolate() when iValues holds no NaN). */
static int get_lower_index(float x, const float* v, int n) {
    int index = -1;
    for(int i = 0; i < n; i++) {
        float c = v[i];
        if(is_valid(c)) {
            if(c < x) index = i;
            else if(c == x) { index = i; break; }
            else if(c > x) break;
        }
    }
    return index;
}
static int get_upper_index(float x, const float* v, int n) {
    int index = -1;
    for(int i = n - 1; i >= 0; i--) {
        float c = v[i];
        if(is_valid(c)) {
            if(c > x) index = i;
            else if(c == x) { index = i; break; }
            else if(c < x) break;
        }
    }
    return index;
}
/* util.cpp:377-414 */
static float interpolate(float x, const float* iX, const float* iY, int n) {
    if(!is_valid(x)) return NAN;
    float y = NAN;
    if(n == 0) return NAN;
    if(x > iX[n - 1]) return iY[n - 1];
    if(x < iX[0]) return iY[0];
    int i0 = get_lower_index(x, iX, n);
    int i1 = get_upper_index(x, iX, n);
    if(i0 < 0 || i1 < 0) return NAN;
    float x0 = iX[i0], x1 = iX[i1], y0 = iY[i0], y1 = iY[i1];
    if(x0 == x1) {
        if(i0 == 0 && i1 == n - 1) y = (y0 + y1) / 2;
        else if(i0 == 0) y = y1;
        else if(i1 == n - 1) y = y0;
        else y = (y0 + y1) / 2;
    }
    else
        y = y0 + (y1 - y0) * (x - x0) / (x1 - x0);
    return y;
}
int orc_interpolate(float x, const float* ix, const float* iy, int n, float* out) {
    *out = interpolate(x, ix, iy, n);
    return 0;
}

/* ------------------------------------------------------------------ neighbourhood -------------------- */
/* neighbourhood.cpp:45-145: double summed-area table + int count table, 4-corner extraction, window clipped */
static void neighbourhood_sat(const float* in, int nY, int nX, int halfwidth, int statistic, float* out) {
    double* values = calloc((size_t) nY * nX, sizeof(double));
    int* counts = calloc((size_t) nY * nX, sizeof(int));
#define V(i, j) values[(size_t) (i) * nX + (j)]
#define K(i, j) counts[(size_t) (i) * nX + (j)]
    for(int i = 0; i < nY; i++)
        for(int j = 0; j < nX; j++) {
            float value = in[(size_t) i * nX + j];
            int ok = is_valid(value);
            if(j == 0 && i == 0) { if(ok) { V(i, j) = value; K(i, j) = 1; } }
            else if(j == 0) { V(i, j) = ok ? V(i - 1, j) + value : V(i - 1, j); K(i, j) = K(i - 1, j) + ok; }
            else if(i == 0) { V(i, j) = ok ? V(i, j - 1) + value : V(i, j - 1); K(i, j) = K(i, j - 1) + ok; }
            else {
                V(i, j) = ok ? V(i, j - 1) + V(i - 1, j) - V(i - 1, j - 1) + value : V(i, j - 1) + V(i - 1, j) - V(i - 1, j - 1);
                K(i, j) = K(i, j - 1) + K(i - 1, j) - K(i - 1, j - 1) + ok;
            }
        }
    #pragma omp parallel for
    for(int i = 0; i < nY; i++)
        for(int j = 0; j < nX; j++) {
            int i1 = i + halfwidth < nY - 1 ? i + halfwidth : nY - 1;
            int j1 = j + halfwidth < nX - 1 ? j + halfwidth : nX - 1;
            int i0 = i - halfwidth - 1, j0 = j - halfwidth - 1;
            double value11 = V(i1, j1), value00 = 0, value10 = 0, value01 = 0;
            int count11 = K(i1, j1), count00 = 0, count10 = 0, count01 = 0;
            if(i0 >= 0 && j0 >= 0) {
                value00 = V(i0, j0); value10 = V(i1, j0); value01 = V(i0, j1);
                count00 = K(i0, j0); count10 = K(i1, j0); count01 = K(i0, j1);
            }
            else if(j0 >= 0) { value10 = V(i1, j0); count10 = K(i1, j0); }
            else if(i0 >= 0) { value01 = V(i0, j1); count01 = K(i0, j1); }
            double value = value11 + value00 - value10 - value01;
            int count = count11 + count00 - count10 - count01;
            float* o = &out[(size_t) i * nX + j];
            *o = NAN;
            if(statistic == COUNT) *o = count;
            else if(count > 0) {
                if(statistic == MEAN) value /= count;
                *o = value;
            }
        }
#undef V
#undef K
    free(values);
    free(counts);
}
/* neighbourhood.cpp:146-210 (Min/Max; the sliver scheme and the border brute force both reduce to the extreme
 * of the valid values in the clipped window) and :557-605 (brute force, any statistic) */
static int neighbourhood_window(const float* in, int nY, int nX, int halfwidth, int statistic, float* out) {
    int rc = 0;
    #pragma omp parallel
    {
        float* hood = malloc(sizeof(float) * (size_t) (2 * halfwidth + 1) * (2 * halfwidth + 1));
        #pragma omp for
        for(int i = 0; i < nY; i++)
            for(int j = 0; j < nX; j++) {
                int n = 0;
                int i0 = i - halfwidth > 0 ? i - halfwidth : 0, i1 = i + halfwidth < nY - 1 ? i + halfwidth : nY - 1;
                int j0 = j - halfwidth > 0 ? j - halfwidth : 0, j1 = j + halfwidth < nX - 1 ? j + halfwidth : nX - 1;
                for(int ii = i0; ii <= i1; ii++)
                    for(int jj = j0; jj <= j1; jj++) hood[n++] = in[(size_t) ii * nX + jj];
                if(orc_calc_statistic(hood, n, statistic, &out[(size_t) i * nX + j])) rc = 2;
            }
        free(hood);
    }
    return rc;
}
/* neighbourhood.cpp:28-242 */
int orc_neighbourhood(const float* input, int ny, int nx, int halfwidth, int statistic, float* output, double* seconds) {
    if(halfwidth < 0) FAIL(1, "Half width must be > 0");
    if(statistic == QUANTILE) FAIL(1, "Use neighbourhood_quantile for computing neighbourhood quantiles");
    if(ny == 0 || nx == 0) return 0;
    double t0 = now_seconds();
    int rc = 0;
    if(statistic == MEAN || statistic == SUM || statistic == COUNT) neighbourhood_sat(input, ny, nx, halfwidth, statistic, output);
    else if(statistic == STD || statistic == VARIANCE) { /* neighbourhood.cpp:211-235: two Mean filters, float arithmetic */
        size_t N = (size_t) ny * nx;
        float *mean = malloc(sizeof(float) * N), *input2 = malloc(sizeof(float) * N), *mean2 = malloc(sizeof(float) * N);
        neighbourhood_sat(input, ny, nx, halfwidth, MEAN, mean);
        for(size_t i = 0; i < N; i++) input2[i] = input[i] * input[i];
        neighbourhood_sat(input2, ny, nx, halfwidth, MEAN, mean2);
        for(size_t i = 0; i < N; i++) {
            float var = mean2[i] - mean[i] * mean[i];
            output[i] = statistic == STD ? sqrtf(var) : var;
        }
        free(mean); free(input2); free(mean2);
    }
    else rc = neighbourhood_window(input, ny, nx, halfwidth, statistic, output);   /* :146-210 (Min/Max) and :237-238 (the rest) */
    if(seconds) *seconds = now_seconds() - t0;
    return rc;
}
int orc_neighbourhood_brute_force(const float* input, int ny, int nx, int halfwidth, int statistic, float* output) {
    if(halfwidth < 0) FAIL(1, "Half width must be > 0");
    if(ny == 0 || nx == 0) return 0;
    return neighbourhood_window(input, ny, nx, halfwidth, statistic, output);
}
/* neighbourhood.cpp:547-630: the brute-force helpers behind neighbourhood_brute_force(vec2 | vec3) (:528-533) and
 * neighbourhood_quantile(vec2 | vec3) (:534-539, statistic == QUANTILE). input is ny x nx x ne, member fastest. */
int orc_neighbourhood_window_ens(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float quantile, float* output) {
    if(halfwidth < 0) FAIL(1, "Half width must be > 0");
    if(ny == 0 || nx == 0 || ne == 0) return 0;
    int rc = 0;
    #pragma omp parallel
    {
        float* hood = malloc(sizeof(float) * (size_t) (2 * halfwidth + 1) * (2 * halfwidth + 1) * ne);
        #pragma omp for
        for(int i = 0; i < ny; i++)
            for(int j = 0; j < nx; j++) {
                int n = 0;
                int i0 = i - halfwidth > 0 ? i - halfwidth : 0, i1 = i + halfwidth < ny - 1 ? i + halfwidth : ny - 1;
                int j0 = j - halfwidth > 0 ? j - halfwidth : 0, j1 = j + halfwidth < nx - 1 ? j + halfwidth : nx - 1;
                for(int ii = i0; ii <= i1; ii++)
                    for(int jj = j0; jj <= j1; jj++)
                        for(int e = 0; e < ne; e++) hood[n++] = input[((size_t) ii * nx + jj) * ne + e];
                float* o = &output[(size_t) i * nx + j];
                int r = statistic == QUANTILE ? orc_calc_quantile(hood, n, quantile, o) : orc_calc_statistic(hood, n, statistic, o);
                if(r) rc = r;
            }
        free(hood);
    }
    return rc;
}

/* neighbourhood.cpp:302-409 (scalar quantile: :296-301) */
int orc_neighbourhood_quantile_fast(const float* input, int nY, int nX, float quantile, const float* quantile_field,
                                    int halfwidth, const float* thresholds, int T, float* output, double* seconds) {
    if(halfwidth < 0) FAIL(1, "Half width must be > 0");
    if(nY == 0 || nX == 0) return 0;
    size_t N = (size_t) nY * nX;
    if(quantile_field) {
        for(size_t i = 0; i < N; i++)
            if(is_valid(quantile_field[i]) && (quantile_field[i] < 0 || quantile_field[i] > 1))
                FAIL(1, "All quantiles must be >= 0 and <= 1");
    }
    else if(is_valid(quantile) && (quantile < 0 || quantile > 1)) FAIL(1, "All quantiles must be >= 0 and <= 1");
    double t0 = now_seconds();
    for(size_t i = 0; i < N; i++) output[i] = NAN;
    if(T == 0) { if(seconds) *seconds = 0; return 0; }
    float* stats = malloc(sizeof(float) * N * (size_t) T);
    float* temp = malloc(sizeof(float) * N);
    for(int t = 0; t < T; t++) { /* neighbourhood.cpp:339-358 */
        for(size_t i = 0; i < N; i++) {
            int sum = 0, count = 0;
            temp[i] = NAN;
            if(is_valid(input[i])) { if(input[i] <= thresholds[t]) sum++; count++; }
            if(count > 0) temp[i] = (float) sum / count;
        }
        neighbourhood_sat(temp, nY, nX, halfwidth, MEAN, stats + (size_t) t * N);
    }
    free(temp);
    #pragma omp parallel
    {
        float* yarray = malloc(sizeof(float) * (size_t) T);
        #pragma omp for
        for(long long i = 0; i < (long long) N; i++) { /* neighbourhood.cpp:367-405 */
            float curr_quantile = quantile_field ? quantile_field[i] : quantile;
            int is_missing = 0;
            for(int t = 0; t < T; t++) {
                float sum = 0;
                int count = 0;
                float st = stats[(size_t) t * N + i];
                if(is_valid(st)) { sum = st; count++; }
                if(count > 0) {
                    yarray[t] = sum / count;
                    if(yarray[t] > 1) yarray[t] = 1;
                    else if(yarray[t] < 0) yarray[t] = 0;
                }
                else is_missing = 1;
            }
            if(!is_missing) {
                if(curr_quantile == 1 && yarray[0] == 1) output[i] = thresholds[0];
                else if(curr_quantile == 0 && yarray[T - 1] == 0) output[i] = thresholds[T - 1];
                else output[i] = interpolate(curr_quantile, yarray, thresholds, T);
            }
        }
        free(yarray);
    }
    free(stats);
    if(seconds) *seconds = now_seconds() - t0;
    return 0;
}

/* gridpp::neighbourhood(vec3, halfwidth, statistic), neighbourhood.cpp:12-27. input is ny x nx x ne, member fastest. */
int orc_neighbourhood_ens(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float* output) {
    if(ny == 0 || nx == 0 || ne == 0) return 0;
    size_t N = (size_t) ny * nx;
    float* flat = malloc(sizeof(float) * N);
    int rc = 0;
    for(size_t i = 0; i < N && rc == 0; i++) rc = orc_calc_statistic(input + i * ne, ne, statistic, &flat[i]); /* :21-23 */
    if(rc == 0) rc = orc_neighbourhood(flat, ny, nx, halfwidth, statistic, output, NULL);                       /* :25 */
    free(flat);
    return rc;
}

/* gridpp::neighbourhood_quantile_fast(vec3, quantile | vec2, halfwidth, thresholds), neighbourhood.cpp:411-527 */
int orc_neighbourhood_quantile_fast_ens(const float* input, int nY, int nX, int nE, float quantile, const float* quantile_field,
                                        int halfwidth, const float* thresholds, int T, float* output) {
    if(halfwidth < 0) FAIL(1, "Half width must be > 0");
    if(nY == 0 || nX == 0 || nE == 0) return 0;
    size_t N = (size_t) nY * nX;
    if(quantile_field) {
        for(size_t i = 0; i < N; i++)
            if(is_valid(quantile_field[i]) && (quantile_field[i] < 0 || quantile_field[i] > 1))
                FAIL(1, "All quantiles must be >= 0 and <= 1");
    }
    else if(is_valid(quantile) && (quantile < 0 || quantile > 1)) FAIL(1, "All quantiles must be >= 0 and <= 1");
    for(size_t i = 0; i < N; i++) output[i] = NAN;
    if(T == 0) return 0;
    float* stats = malloc(sizeof(float) * N * (size_t) T);
    float* temp = malloc(sizeof(float) * N);
    for(int t = 0; t < T; t++) { /* :456-472 */
        for(size_t i = 0; i < N; i++) {
            int sum = 0, count = 0;
            temp[i] = NAN;
            for(int e = 0; e < nE; e++) {
                float v = input[i * nE + e];
                if(is_valid(v)) { if(v <= thresholds[t]) sum++; count++; }
            }
            if(count > 0) temp[i] = (float) sum / count;
        }
        neighbourhood_sat(temp, nY, nX, halfwidth, MEAN, stats + (size_t) t * N);
    }
    free(temp);
    float* yarray = malloc(sizeof(float) * (size_t) T);
    for(size_t i = 0; i < N; i++) { /* :483-521 */
        float curr_quantile = quantile_field ? quantile_field[i] : quantile;
        int is_missing = 0;
        for(int t = 0; t < T; t++) {
            float sum = 0;
            int count = 0;
            float st = stats[(size_t) t * N + i];
            for(int e = 0; e < nE; e++)
                if(is_valid(st)) { sum += st; count++; }
            if(count > 0) {
                yarray[t] = sum / count;
                if(yarray[t] > 1) yarray[t] = 1;
                else if(yarray[t] < 0) yarray[t] = 0;
            }
            else is_missing = 1;
        }
        if(!is_missing) {
            if(curr_quantile == 1 && yarray[0] == 1) output[i] = thresholds[0];
            else if(curr_quantile == 0 && yarray[T - 1] == 0) output[i] = thresholds[T - 1];
            else output[i] = interpolate(curr_quantile, yarray, thresholds, T);
        }
    }
    free(yarray);
    free(stats);
    return 0;
}

/* util.cpp:261-338 */
static int calc_even_quantiles(const float* sorted, int size, int num, float* q) {
    int nq = 0;
    if(num == 0 || size == 0) return 0;
    if(num >= size) {
        q[nq++] = sorted[0];
        for(int i = 1; i < size; i++)
            if(sorted[i] != sorted[i - 1]) q[nq++] = sorted[i];
        return nq;
    }
    float lowest = sorted[0], highest = sorted[size - 1];
    int count_lower = 0;
    for(int i = 0; i < size; i++) { if(sorted[i] != lowest) break; count_lower++; }
    q[nq++] = lowest;
    if(num == 2) { if(lowest != highest) q[nq++] = highest; return nq; }
    int repeated_at_beginning = count_lower < size && count_lower > size / num;
    if(repeated_at_beginning) q[nq++] = sorted[count_lower];
    float last_added = q[nq - 1];
    float* uniq = malloc(sizeof(float) * (size_t) size);
    int nu = 0;
    for(int i = 0; i < size; i++)
        if(sorted[i] > last_added && (nu == 0 || sorted[i] != uniq[nu - 1])) uniq[nu++] = sorted[i];
    if(nu > 0) {
        int num_left = num - nq;
        for(int i = 1; i <= num_left; i++) {
            float f = (float) i / (num_left);
            int index = nu * f - 1;
            if(index >= 0) q[nq++] = uniq[index];
            else { free(uniq); return -1; }
        }
    }
    free(uniq);
    return nq;
}
/* neighbourhood.cpp:243-266 */
int orc_get_neighbourhood_thresholds(const float* input, int ny, int nx, int num, float* out, int* out_n) {
    if(num <= 0) FAIL(1, "num_thresholds must be > 0");
    *out_n = 0;
    if(ny == 0 || nx == 0) return 0;
    size_t N = (size_t) ny * nx;
    float* all = malloc(sizeof(float) * N);
    int n = 0;
    for(size_t i = 0; i < N; i++)
        if(is_valid(input[i])) all[n++] = input[i];
    qsort(all, (size_t) n, sizeof(float), cmp_float);
    float* q = malloc(sizeof(float) * (size_t) (num + n + 2));
    int nq = calc_even_quantiles(all, n, num, q);
    free(all);
    if(nq < 0) { free(q); FAIL(2, "Internal error in calc_even_quantiles."); }
    *out_n = nq;
    for(int i = 0; i < nq && i < num; i++) out[i] = q[i];
    free(q);
    return 0;
}


/* ------------------------------------------------------------------ consumers of the point index ----- */
/* A Grid is its flattened nodes: every function takes flat lat / lon arrays for both sides. */
/* gridding.cpp:6-61 (both overloads) */
int orc_gridding(const float* olats, const float* olons, int nO, const float* ilats, const float* ilons, int nI, int type, const float* values,
                 float radius, int min_num, int statistic, float* output) {
    if(!is_valid(radius) || radius < 0) FAIL(1, "radius must be >= 0");
    if(min_num < 0) FAIL(1, "min_num must be >= 0");
    pts_t ip;
    int rc = pts_make(&ip, ilats, ilons, NULL, NULL, nI, type);
    if(rc) return rc;
    cells_t cells;
    cells_build(&cells, &ip, radius > 0 ? (double) radius : 1.0);
    int err = 0;
    #pragma omp parallel
    {
        int* I = NULL;
        int cap = 0;
        float* curr = NULL;
        int ccap = 0;
        #pragma omp for
        for(int o = 0; o < nO; o++) {
            float x, y, z;
            output[o] = NAN;
            if(convert_one(olats[o], olons[o], type, &x, &y, &z)) { err = 1; continue; }
            int n = radius_query(&ip, &cells, x, y, z, radius, 1, &I, &cap);
            if(min_num <= 0 || n >= min_num) {
                if(n > ccap) { ccap = n; curr = realloc(curr, sizeof(float) * (size_t) ccap); }
                for(int i = 0; i < n; i++) curr[i] = values[I[i]];
                if(orc_calc_statistic(curr, n, statistic, &output[o])) err = 2;
            }
        }
        free(I);
        free(curr);
    }
    cells_free(&cells);
    pts_free(&ip);
    if(err == 1) FAIL(1, "Invalid coords in query");
    if(err == 2) FAIL(2, "Internal error. Cannot compute statistic");
    return 0;
}
/* gridding.cpp:63-131 (both overloads) */
int orc_gridding_nearest(const float* olats, const float* olons, int nO, const float* ilats, const float* ilons, int nI, int type,
                         const float* values, int min_num, int statistic, float* output) {
    if(min_num < 0) FAIL(1, "min_num must be >= 0");
    pts_t op;
    int rc = pts_make(&op, olats, olons, NULL, NULL, nO, type);
    if(rc) return rc;
    int* node = malloc(sizeof(int) * (size_t) (nI > 0 ? nI : 1));
    int* count = calloc((size_t) (nO > 0 ? nO : 1), sizeof(int));
    for(int s = 0; s < nI; s++) {
        float x, y, z;
        if(convert_one(ilats[s], ilons[s], type, &x, &y, &z)) { free(node); free(count); pts_free(&op); FAIL(1, "Invalid coords"); }
        int idx = -1;
        node[s] = closest_query(&op, x, y, z, 1, 1, &idx) ? idx : -1;
        if(node[s] >= 0) count[node[s]]++;
    }
    float* curr = malloc(sizeof(float) * (size_t) (nI > 0 ? nI : 1));
    int err = 0;
    for(int o = 0; o < nO; o++) {
        output[o] = NAN;
        if(count[o] == 0) continue;
        if(min_num > 0 && count[o] < min_num) continue;
        int n = 0;
        for(int s = 0; s < nI; s++)
            if(node[s] == o) curr[n++] = values[s];
        if(orc_calc_statistic(curr, n, statistic, &output[o])) err = 2;
    }
    free(curr); free(node); free(count);
    pts_free(&op);
    if(err) FAIL(2, "Internal error. Cannot compute statistic");
    return 0;
}
/* count.cpp:6-66 (all four overloads) */
int orc_count(const float* ilats, const float* ilons, int nI, const float* olats, const float* olons, int nO, int type, float radius,
              float* output) {
    pts_t ip;
    int rc = pts_make(&ip, ilats, ilons, NULL, NULL, nI, type);
    if(rc) return rc;
    cells_t cells;
    cells_build(&cells, &ip, radius > 0 && is_valid(radius) ? (double) radius : 1.0);
    int err = 0;
    #pragma omp parallel
    {
        int* I = NULL;
        int cap = 0;
        #pragma omp for
        for(int o = 0; o < nO; o++) {
            float x, y, z;
            if(convert_one(olats[o], olons[o], type, &x, &y, &z)) { err = 1; continue; }
            output[o] = (float) radius_query(&ip, &cells, x, y, z, radius, 1, &I, &cap);
        }
        free(I);
    }
    cells_free(&cells);
    pts_free(&ip);
    if(err) FAIL(1, "Invalid coords in query");
    return 0;
}
/* distance.cpp:6-120; query_first: calc_distance(output location, input point) (:21, :111) or the reverse (:52, :83) */
int orc_distance(const float* ilats, const float* ilons, int nI, const float* olats, const float* olons, int nO, int type, int num,
                 int query_first, float* output) {
    pts_t ip;
    int rc = pts_make(&ip, ilats, ilons, NULL, NULL, nI, type);
    if(rc) return rc;
    int err = 0;
    #pragma omp parallel
    {
        int* idx = malloc(sizeof(int) * (size_t) (num > 0 ? num : 1));
        #pragma omp for
        for(int o = 0; o < nO; o++) {
            float x, y, z;
            if(convert_one(olats[o], olons[o], type, &x, &y, &z)) { err = 1; continue; }
            int found = closest_query(&ip, x, y, z, num, 1, idx);
            float max_dist = 0;
            for(int k = 0; k < found; k++) {
                float dist = query_first ? orc_calc_distance(olats[o], olons[o], ilats[idx[k]], ilons[idx[k]], type)
                                         : orc_calc_distance(ilats[idx[k]], ilons[idx[k]], olats[o], olons[o], type);
                if(dist > max_dist) max_dist = dist;
            }
            output[o] = max_dist;
        }
        free(idx);
    }
    pts_free(&ip);
    if(err) FAIL(1, "Invalid coords in query");
    return 0;
}
/* fill.cpp:6-43 (mode 1: inside takes `value`; mode 2: outside takes `value`) and doping_circle, doping.cpp:52-93 (mode 0) */
static int circles(const float* glats, const float* glons, const float* gelevs, int nG, const float* plats, const float* plons, const float* pelevs,
                   int nP, int type, const float* radii, int mode, const float* input, const float* obs, float value, float max_elev_diff,
                   float* output) {
    pts_t gp;
    int rc = pts_make(&gp, glats, glons, gelevs, NULL, nG, type);
    if(rc) return rc;
    float rmax = 0;
    for(int i = 0; i < nP; i++) if(radii[i] > rmax) rmax = radii[i];
    cells_t cells;
    cells_build(&cells, &gp, rmax > 0 ? (double) rmax : 1.0);
    for(int c = 0; c < nG; c++) output[c] = mode == 2 ? value : input[c];
    int* I = NULL;
    int cap = 0, check_elev = is_valid(max_elev_diff) && mode == 0;
    for(int i = 0; i < nP; i++) {
        float x, y, z;
        if(convert_one(plats[i], plons[i], type, &x, &y, &z)) { cells_free(&cells); pts_free(&gp); free(I); FAIL(1, "Invalid coords"); }
        int n = radius_query(&gp, &cells, x, y, z, radii[i], 1, &I, &cap);
        for(int j = 0; j < n; j++) {
            if(check_elev) {
                float diff = fabsf(pelevs[i] - gp.elev[I[j]]);
                if(diff > max_elev_diff) continue;
            }
            output[I[j]] = mode == 0 ? obs[i] : (mode == 1 ? value : input[I[j]]);
        }
    }
    free(I);
    cells_free(&cells);
    pts_free(&gp);
    return 0;
}
int orc_fill(const float* glats, const float* glons, int nG, const float* input, const float* plats, const float* plons, int nP, int type,
             const float* radii, float value, int outside, float* output) {
    for(int i = 0; i < nP; i++) if(radii[i] < 0) FAIL(1, "All radius sizes must be 0 or greater");
    return circles(glats, glons, NULL, nG, plats, plons, NULL, nP, type, radii, outside ? 2 : 1, input, NULL, value, NAN, output);
}
int orc_doping_circle(const float* glats, const float* glons, const float* gelevs, int nG, const float* background, const float* plats,
                      const float* plons, const float* pelevs, int nP, int type, const float* obs, const float* radii, float max_elev_diff,
                      float* output) {
    if(is_valid(max_elev_diff) && max_elev_diff < 0) FAIL(1, "max_elev_diff must be greater than or equal to 0");
    for(int i = 0; i < nP; i++) if(radii[i] < 0) FAIL(1, "radii must be greater than or equal to 0");
    return circles(glats, glons, gelevs, nG, plats, plons, pelevs, nP, type, radii, 0, background, obs, 0, max_elev_diff, output);
}
/* doping.cpp:5-51 */
int orc_doping_square(const float* glats, const float* glons, const float* gelevs, int ny, int nx, const float* background, const float* plats,
                      const float* plons, const float* pelevs, int nP, int type, const float* obs, const int* halfwidth, float max_elev_diff,
                      float* output) {
    if(is_valid(max_elev_diff) && max_elev_diff < 0) FAIL(1, "max_elev_diff must be greater than or equal to 0");
    for(int i = 0; i < nP; i++) if(halfwidth[i] < 0) FAIL(1, "All halfwidth must be greater than or equal to 0");
    pts_t gp;
    int rc = pts_make(&gp, glats, glons, gelevs, NULL, ny * nx, type);
    if(rc) return rc;
    for(int c = 0; c < ny * nx; c++) output[c] = background[c];
    int check_elev = is_valid(max_elev_diff);
    for(int i = 0; i < nP; i++) {
        float x, y, z;
        if(convert_one(plats[i], plons[i], type, &x, &y, &z)) { pts_free(&gp); FAIL(1, "Invalid coords"); }
        int node;
        if(!closest_query(&gp, x, y, z, 1, 1, &node)) continue;
        int cy = node / nx, cx = node % nx;
        for(int yy = cy - halfwidth[i] > 0 ? cy - halfwidth[i] : 0; yy <= (cy + halfwidth[i] < ny - 1 ? cy + halfwidth[i] : ny - 1); yy++)
            for(int xx = cx - halfwidth[i] > 0 ? cx - halfwidth[i] : 0; xx <= (cx + halfwidth[i] < nx - 1 ? cx + halfwidth[i] : nx - 1); xx++) {
                if(check_elev) {
                    float diff = fabsf(pelevs[i] - gp.elev[yy * nx + xx]);
                    if(diff > max_elev_diff) continue;
                }
                output[yy * nx + xx] = obs[i];
            }
    }
    pts_free(&gp);
    return 0;
}
/* fill.cpp:44-134 */
int orc_fill_missing(const float* values, int Y, int X, float* output) {
    size_t N = (size_t) Y * X;
    float *ry = malloc(sizeof(float) * (N ? N : 1)), *rx = malloc(sizeof(float) * (N ? N : 1));
    for(size_t i = 0; i < N; i++) { ry[i] = NAN; rx[i] = NAN; }
    for(int y = 0; y < Y; y++) {
        int last = 0, next = -1;
        for(int x = 0; x < X; x++) {
            float curr = values[(size_t) y * X + x];
            if(!is_valid(curr)) {
                if(next < x)
                    for(next = x; next < X; next++)
                        if(is_valid(values[(size_t) y * X + next])) break;
                if(next >= X) continue;
                float value_last = values[(size_t) y * X + last], value_next = values[(size_t) y * X + next];
                ry[(size_t) y * X + x] = (value_last) + (value_next - value_last) * (x - last) / (next - last);
            }
            else { last = x; ry[(size_t) y * X + x] = curr; }
        }
    }
    for(int x = 0; x < X; x++) {
        int last = 0, next = -1;
        for(int y = 0; y < Y; y++) {
            float curr = values[(size_t) y * X + x];
            if(!is_valid(curr)) {
                if(next < y)
                    for(next = y; next < Y; next++)
                        if(is_valid(values[(size_t) next * X + x])) break;
                if(next >= Y) continue;
                float value_last = values[(size_t) last * X + x], value_next = values[(size_t) next * X + x];
                rx[(size_t) y * X + x] = (value_last) + (value_next - value_last) * (y - last) / (next - last);
            }
            else { last = y; rx[(size_t) y * X + x] = curr; }
        }
    }
    for(size_t i = 0; i < N; i++) {
        int count = 0;
        float total = 0;
        if(is_valid(ry[i])) { total += ry[i]; count++; }
        if(is_valid(rx[i])) { total += rx[i]; count++; }
        output[i] = count > 0 ? total / count : NAN;
    }
    free(ry); free(rx);
    return 0;
}

/* ------------------------------------------------------------------ ensi_multi / staticcorr_points ---- */
/* The observation selection shared by oi_ensi_multi.cpp:436-486,716-767,1008-1056 and corr_points.cpp:61-110: neighbours within
 * the localization radius in the order of the radius query, those with a valid value (ok[index]) and rho > 0, cut to the
 * max_points best (best first) when there are more. Returns lS; cand[0..lS) holds the selection. */
static int select_observations(const pts_t* op, const cells_t* cells, const orc_structure* s, pt_t p1, const char* ok, int max_points,
                               int** nb, int* nb_cap, cand_t** cand, int* cand_cap) {
    float localizationRadius = structure_loc_dist(s);
    int n0 = radius_query(op, cells, p1.x, p1.y, p1.z, localizationRadius, 1, nb, nb_cap);
    if(n0 == 0) return 0;
    if(n0 > *cand_cap) { *cand_cap = 2 * n0; *cand = realloc(*cand, sizeof(cand_t) * (size_t) *cand_cap); }
    int nc = 0;
    for(int i = 0; i < n0; i++) {
        int index = (*nb)[i];
        pt_t p2 = {op->x[index], op->y[index], op->z[index], op->elev[index], op->laf[index]};
        float rho = structure_corr_background(s, p1, p2);
        if((!ok || ok[index]) && rho > 0) { (*cand)[nc].rho = rho; (*cand)[nc].pos = i; (*cand)[nc].index = index; nc++; }
    }
    if(max_points > 0 && nc > max_points) { qsort(*cand, (size_t) nc, sizeof(cand_t), cmp_cand); return max_points; }
    return nc;
}

/* corr_points.cpp:26-131: out is nY x nS, row y holds corr_background(point y, knot) for the selected knots, 0 elsewhere */
int orc_staticcorr_points(const float* lats, const float* lons, const float* elevs, const float* lafs, int nY, const float* klats,
                          const float* klons, const float* kelevs, const float* klafs, int nS, int type, const orc_structure* s,
                          int max_points, float* output) {
    if(max_points < 0) FAIL(1, "max_points must be >= 0");
    pts_t bp, op;
    int rc = pts_make(&bp, lats, lons, elevs, lafs, nY, type);
    if(rc) return rc;
    rc = pts_make(&op, klats, klons, kelevs, klafs, nS, type);
    if(rc) { pts_free(&bp); return rc; }
    memset(output, 0, sizeof(float) * (size_t) nY * nS);
    cells_t cells;
    float R0 = structure_loc_dist(s);
    cells_build(&cells, &op, R0 > 0 ? 0.5 * R0 : 0);
    #pragma omp parallel
    {
        int* nb = NULL;
        int nb_cap = 0, cand_cap = 0;
        cand_t* cand = NULL;
        #pragma omp for
        for(int y = 0; y < nY; y++) {
            pt_t p1 = {bp.x[y], bp.y[y], bp.z[y], bp.elev[y], bp.laf[y]};
            int lS = select_observations(&op, &cells, s, p1, NULL, max_points, &nb, &nb_cap, &cand, &cand_cap);
            for(int i = 0; i < lS; i++) output[(size_t) y * nS + cand[i].index] = (float) (double) cand[i].rho; /* lRhos is an arma::vec */
        }
        free(nb);
        free(cand);
    }
    cells_free(&cells);
    pts_free(&bp);
    pts_free(&op);
    return 0;
}

/* members without an invalid value anywhere in the given fields (oi_ensi_multi.cpp:399-421,693-712,933-953); returns their number */
static int valid_members(const float* a, const float* a2, int nA, const float* b, const float* b2, int nB, int nEns, int* validEns) {
    int n = 0;
    for(int e = 0; e < nEns; e++) {
        int bad = 0;
        for(int y = 0; y < nA && !bad; y++) bad = !is_valid(a[(size_t) y * nEns + e]) || (a2 && !is_valid(a2[(size_t) y * nEns + e]));
        for(int i = 0; i < nB && !bad; i++) bad = !is_valid(b[(size_t) i * nEns + e]) || (b2 && !is_valid(b2[(size_t) i * nEns + e]));
        if(!bad) validEns[n++] = e;
    }
    return n;
}
/* 1 / sqrt(n - 1) * (v - mean) / std with mean / std from calc_statistic over the valid members (oi_ensi_multi.cpp:427-448,
 * 501-510): out[e] as a double; zeros when the statistics are invalid or std <= 0.0013 */
static void normalised_perturbations(const float* row, const int* validEns, int E, double* out) {
    float* v = malloc(sizeof(float) * (size_t) (E > 0 ? E : 1));
    for(int e = 0; e < E; e++) v[e] = row[validEns[e]];
    float mean, std;
    orc_calc_statistic(v, E, MEAN, &mean);
    orc_calc_statistic(v, E, STD, &std);
    float default_min_std = 0.0013;
    for(int e = 0; e < E; e++) out[e] = 0;
    if(is_valid(mean) && is_valid(std) && std > default_min_std)
        for(int e = 0; e < E; e++) out[e] = 1 / sqrt((double) (E - 1)) * (v[e] - mean) / std;
    free(v);
}
/* oi_ensi_multi.cpp:582-609,825-852: the anti-extrapolation filter of ebe / ebesc on one member's increment */
static double clamp_increment(double dx, double maxIncD, double minIncD) {
    float increment = dx, maxInc = maxIncD, minInc = minIncD;
    if(maxInc > 0 && increment > maxInc) increment = maxInc;
    else if(maxInc < 0 && increment > 0) increment = 0;
    else if(minInc < 0 && increment < minInc) increment = minInc;
    else if(minInc > 0 && increment < 0) increment = 0;
    return increment;
}

/* oi_ensi_multi.cpp:329-628 (ebe: ensemble-based correlations, with_ens != 0; background_corr / pbackground_corr given) and
 * :630-860 (ebesc: static correlations, with_ens == 0). background nB x nEns, pobs / pbackground nS x nEns.
 * The reference addresses the innovation matrix (lS x nValidEns) by the ORIGINAL member index (:564,:808): with an invalid member
 * that is not the last one it writes past the matrix (Armadillo throws); here that case returns an error as well. */
int orc_ensi_multi_ebe(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB, const float* bratios,
                       const float* background, const float* background_corr, int nEns, const float* plats, const float* plons,
                       const float* pelevs, const float* plafs, int nS, int type, const float* pobs, const float* pratios,
                       const float* pbackground, const float* pbackground_corr, const orc_structure* s, int max_points,
                       int allow_extrapolation, int with_ens, float* analysis) {
    if(max_points < 0) FAIL(1, "max_points must be >= 0");
    pts_t bp, op;
    int rc = pts_make(&bp, blats, blons, belevs, blafs, nB, type);
    if(rc) return rc;
    rc = pts_make(&op, plats, plons, pelevs, plafs, nS, type);
    if(rc) { pts_free(&bp); return rc; }
    memcpy(analysis, background, sizeof(float) * (size_t) nB * nEns);
    int* validEns = malloc(sizeof(int) * (size_t) (nEns > 0 ? nEns : 1));
    int E = (nS == 0 || nB == 0) ? 0 : valid_members(background, with_ens ? background_corr : NULL, nB, pbackground, with_ens ? pbackground_corr : NULL, nS, nEns, validEns);
    if(E == 0) { free(validEns); pts_free(&bp); pts_free(&op); return 0; }
    for(int e = 0; e < E; e++)
        if(validEns[e] >= E) { free(validEns); pts_free(&bp); pts_free(&op); FAIL(2, "Mat::operator(): index out of bounds"); }
    char* ok = malloc((size_t) nS);
    for(int i = 0; i < nS; i++) ok[i] = is_valid(pobs[(size_t) i * nEns]) ? 1 : 0;   /* pobs[index][0], :463,:742 */
    double* gZ = NULL;   /* :424-449, stored as float in the reference (vec2 gZ_R) */
    if(with_ens) {
        gZ = malloc(sizeof(double) * (size_t) nS * E);
        for(int i = 0; i < nS; i++) {
            normalised_perturbations(pbackground_corr + (size_t) i * nEns, validEns, E, gZ + (size_t) i * E);
            for(int e = 0; e < E; e++) gZ[(size_t) i * E + e] = (float) gZ[(size_t) i * E + e];
        }
    }
    cells_t cells;
    float R0 = structure_loc_dist(s);
    cells_build(&cells, &op, R0 > 0 ? 0.5 * R0 : 0);
    int err = 0;
    #pragma omp parallel
    {
        int* nb = NULL;
        int nb_cap = 0, cand_cap = 0;
        cand_t* cand = NULL;
        double* XL = malloc(sizeof(double) * (size_t) E);
        #pragma omp for schedule(dynamic, 16)
        for(int y = 0; y < nB; y++) {
            pt_t p1 = {bp.x[y], bp.y[y], bp.z[y], bp.elev[y], bp.laf[y]};
            int lS = select_observations(&op, &cells, s, p1, ok, max_points, &nb, &nb_cap, &cand, &cand_cap);
            if(lS == 0) continue;
            if(with_ens) normalised_perturbations(background_corr + (size_t) y * nEns, validEns, E, XL);   /* lX_L, :495-510 */
            double* A = malloc(sizeof(double) * (size_t) lS * lS);      /* lR_rr + lR_dd (ebe) / lCorr2D + lR_dd (ebesc), column-major */
            double* Ainv = malloc(sizeof(double) * (size_t) lS * lS);
            double* r = malloc(sizeof(double) * (size_t) lS);           /* lr_lr (ebe) / lCorr1D (ebesc) */
            double* K = malloc(sizeof(double) * (size_t) lS);
            for(int i = 0; i < lS; i++) {
                int index = cand[i].index;
                pt_t pi = {op.x[index], op.y[index], op.z[index], op.elev[index], op.laf[index]};
                double xz = 1;
                if(with_ens) {
                    xz = 0;
                    for(int e = 0; e < E; e++) xz += XL[e] * gZ[(size_t) index * E + e];
                }
                r[i] = (double) cand[i].rho * xz;
                for(int j = 0; j < lS; j++) {
                    int index_j = cand[j].index;
                    pt_t pj = {op.x[index_j], op.y[index_j], op.z[index_j], op.elev[index_j], op.laf[index_j]};
                    double zz = 1;
                    if(with_ens) {
                        zz = 0;
                        for(int e = 0; e < E; e++) zz += gZ[(size_t) index * E + e] * gZ[(size_t) index_j * E + e];
                    }
                    A[i + (size_t) j * lS] = (double) structure_corr(s, pi, pj) * zz + (i == j ? (double) pratios[index] : 0.0);
                }
            }
            if(!mat_inv(A, Ainv, lS)) { err = 2; free(A); free(Ainv); free(r); free(K); continue; }
            for(int j = 0; j < lS; j++) {   /* lK = r * inv(A), :577,:820 */
                double acc = 0;
                for(int i = 0; i < lS; i++) acc += r[i] * Ainv[i + (size_t) j * lS];
                K[j] = acc;
            }
            float std_ratios_lr = bratios[y];
            for(int e = 0; e < E; e++) {
                int ei = validEns[e];   /* == e, see above */
                double acc = 0, mx = -INFINITY, mn = INFINITY;
                for(int i = 0; i < lS; i++) {
                    double innov = (float) (pobs[(size_t) cand[i].index * nEns + ei] - pbackground[(size_t) cand[i].index * nEns + ei]);
                    acc += K[i] * innov;
                    if(innov > mx) mx = innov;
                    if(innov < mn) mn = innov;
                }
                double dx = std_ratios_lr * acc;
                if(!allow_extrapolation) dx = clamp_increment(dx, mx, mn);
                analysis[(size_t) y * nEns + ei] = background[(size_t) y * nEns + ei] + dx;
            }
            free(A); free(Ainv); free(r); free(K);
        }
        free(nb);
        free(cand);
        free(XL);
    }
    cells_free(&cells);
    free(gZ); free(ok); free(validEns);
    pts_free(&bp);
    pts_free(&op);
    if(err) FAIL(2, "inv(): matrix is singular");
    return 0;
}

/* oi_ensi_multi.cpp:862-1311 (utem: the transform is computed from the standardised *_corr ensembles and applied to the
 * ensemble mean / spread of background). pobs nS; background* nB x nEns; pbackground* nS x nEns. *num_skipped (may be NULL)
 * counts the points left at their raw values because rcond(Pinv) <= 0 (:1106-1110). */
int orc_ensi_multi_utem(const float* blats, const float* blons, const float* belevs, const float* blafs, int nB, const float* bratios,
                        const float* background, const float* background_corr, int nEns, const float* plats, const float* plons,
                        const float* pelevs, const float* plafs, int nS, int type, const float* pobs, const float* pratios,
                        const float* pbackground, const float* pbackground_corr, const orc_structure* s, int max_points,
                        int allow_extrapolation, float* analysis) {
    if(max_points < 0) FAIL(1, "max_points must be >= 0");
    pts_t bp, op;
    int rc = pts_make(&bp, blats, blons, belevs, blafs, nB, type);
    if(rc) return rc;
    rc = pts_make(&op, plats, plons, pelevs, plafs, nS, type);
    if(rc) { pts_free(&bp); return rc; }
    memcpy(analysis, background, sizeof(float) * (size_t) nB * nEns);
    int* validEns = malloc(sizeof(int) * (size_t) (nEns > 0 ? nEns : 1));
    const int E = (nS == 0 || nB == 0) ? 0 : valid_members(background, background_corr, nB, pbackground, pbackground_corr, nS, nEns, validEns);
    if(E == 0) { free(validEns); pts_free(&bp); pts_free(&op); return 0; }
    const float default_min_std = 0.0013;
    const float const_fact = 1 / sqrt((double) (E - 1));   /* :959, a float here (ebe keeps the double) */
    float* gY = malloc(sizeof(float) * (size_t) nS * E);
    float* gY_corr = malloc(sizeof(float) * (size_t) nS * E);
    float* gYhat = malloc(sizeof(float) * (size_t) nS);
    char* ok = malloc((size_t) nS);
    float* tmp = malloc(sizeof(float) * (size_t) E);
    for(int i = 0; i < nS; i++) {   /* :960-994 */
        ok[i] = is_valid(pobs[i]) ? 1 : 0;
        for(int e = 0; e < E; e++) tmp[e] = pbackground[(size_t) i * nEns + validEns[e]];
        float mean;
        orc_calc_statistic(tmp, E, MEAN, &mean);
        for(int e = 0; e < E; e++) gY[(size_t) i * E + e] = is_valid(mean) ? tmp[e] - mean : 0;
        gYhat[i] = mean;
        for(int e = 0; e < E; e++) tmp[e] = pbackground_corr[(size_t) i * nEns + validEns[e]];
        float mean_corr, std_corr;
        orc_calc_statistic(tmp, E, MEAN, &mean_corr);
        orc_calc_statistic(tmp, E, STD, &std_corr);
        int use = is_valid(mean_corr) && is_valid(std_corr) && std_corr > default_min_std;
        for(int e = 0; e < E; e++) gY_corr[(size_t) i * E + e] = use ? const_fact * (tmp[e] - mean_corr) / std_corr : 0;
    }
    free(tmp);
    cells_t cells;
    float R0 = structure_loc_dist(s);
    cells_build(&cells, &op, R0 > 0 ? 0.5 * R0 : 0);
    #pragma omp parallel
    {
        int* nb = NULL;
        int nb_cap = 0, cand_cap = 0;
        cand_t* cand = NULL;
        double* Pinv = malloc(sizeof(double) * (size_t) E * E);
        double* P = malloc(sizeof(double) * (size_t) E * E);
        double* S = malloc(sizeof(double) * (size_t) E * E);
        double* val = malloc(sizeof(double) * (size_t) E);
        double* vec = malloc(sizeof(double) * (size_t) E * E);
        double* W = malloc(sizeof(double) * (size_t) E * E);
        double* w = malloc(sizeof(double) * (size_t) E);
        double* X = malloc(sizeof(double) * (size_t) E);
        double* X_corr = malloc(sizeof(double) * (size_t) E);
        float* X1 = malloc(sizeof(float) * (size_t) E);
        float* X_corr1 = malloc(sizeof(float) * (size_t) E);
        #pragma omp for schedule(dynamic, 16)
        for(int y = 0; y < nB; y++) {
            pt_t p1 = {bp.x[y], bp.y[y], bp.z[y], bp.elev[y], bp.laf[y]};
            int lS = select_observations(&op, &cells, s, p1, ok, max_points, &nb, &nb_cap, &cand, &cand_cap);
            if(lS == 0) continue;
            float std_ratios_lr = bratios[y];
            double* lY = malloc(sizeof(double) * (size_t) lS * E);        /* column-major, like the Armadillo matrix */
            double* lYc = malloc(sizeof(double) * (size_t) lS * E);
            double* Cm = malloc(sizeof(double) * (size_t) E * lS);
            double* dd = malloc(sizeof(double) * (size_t) lS);
            for(int i = 0; i < lS; i++) {
                int index = cand[i].index;
                double rinv = (double) cand[i].rho / pratios[index];        /* :1094 */
                for(int e = 0; e < E; e++) {
                    lY[i + (size_t) e * lS] = gY[(size_t) index * E + e];
                    lYc[i + (size_t) e * lS] = gY_corr[(size_t) index * E + e];
                    Cm[e + (size_t) i * E] = lYc[i + (size_t) e * lS] * rinv;
                }
                dd[i] = (double) pobs[index] - (double) gYhat[index];
            }
            for(int a = 0; a < E; a++)
                for(int b = 0; b < E; b++) {
                    double acc = 0;
                    for(int i = 0; i < lS; i++) acc += Cm[a + (size_t) i * E] * lYc[i + (size_t) b * lS];
                    Pinv[a + (size_t) b * E] = acc + (a == b ? 1.0 : 0.0);   /* :1105 */
                }
            int inv_ok = mat_inv(Pinv, P, E);
            float cond = 0;
            if(inv_ok) { double nn = mat_norm1(Pinv, E) * mat_norm1(P, E); cond = (nn != nn || nn == 0) ? 0 : 1.0 / nn; }
            if(inv_ok && cond > 0) {
                for(int i = 0; i < E * E; i++) S[i] = (double) (E - 1) * P[i];
                mat_eig_sym(S, val, vec, E);
                for(int a = 0; a < E; a++)
                    for(int b = 0; b < E; b++) {
                        double acc = 0;
                        for(int k = 0; k < E; k++) acc += vec[a + (size_t) k * E] * sqrt(val[k]) * vec[b + (size_t) k * E];
                        W[a + (size_t) b * E] = acc;
                    }
                for(int a = 0; a < E; a++) {
                    double acc = 0;
                    for(int i = 0; i < lS; i++) {
                        double pc = 0;
                        for(int b = 0; b < E; b++) pc += P[a + (size_t) b * E] * Cm[b + (size_t) i * E];
                        acc += pc * dd[i];
                    }
                    w[a] = acc;
                }
                /* :1160-1198; every member is valid here (the members were screened above) */
                float total = 0, total_corr = 0;
                for(int e = 0; e < E; e++) {
                    float value = background[(size_t) y * nEns + validEns[e]], value_corr = background_corr[(size_t) y * nEns + validEns[e]];
                    X[e] = X1[e] = value;
                    X_corr[e] = X_corr1[e] = value_corr;
                    total += value;
                    total_corr += value_corr;
                }
                float ensMean = total / E, ensMean_corr = total_corr / E, ensStd, ensStd_corr;
                orc_calc_statistic(X1, E, STD, &ensStd);
                orc_calc_statistic(X_corr1, E, STD, &ensStd_corr);
                for(int e = 0; e < E; e++) {
                    X[e] -= ensMean;
                    float value_corr = X_corr[e];
                    X_corr[e] = ensStd_corr <= default_min_std ? 0 : const_fact * (value_corr - ensMean_corr) / ensStd_corr;
                }
                for(int a = 0; a < E; a++)
                    for(int b = 0; b < E; b++) W[a + (size_t) b * E] = ensStd * W[a + (size_t) b * E] + std_ratios_lr * w[a];   /* :1201-1205 */
                for(int e = 0; e < E; e++) {
                    float tot = 0;
                    for(int k = 0; k < E; k++) tot += X_corr[k] * W[k + (size_t) e * E];
                    float currIncrement = tot;
                    if(!allow_extrapolation) {
                        double lYe = lY[e];   /* a LINEAR index into the lS x E matrix, :1266-1267 */
                        double mx = -INFINITY, mn = INFINITY;
                        for(int i = 0; i < lS; i++) {
                            double v = (double) pobs[cand[i].index] - (lYe + (double) gYhat[cand[i].index]);
                            if(v > mx) mx = v;
                            if(v < mn) mn = v;
                        }
                        float maxInc = mx, minInc = mn;
                        float memberIncrement = currIncrement - X[e];
                        if(maxInc > 0 && memberIncrement > maxInc) currIncrement = maxInc + X[e];
                        else if(maxInc < 0 && memberIncrement > 0) currIncrement = 0 + X[e];
                        else if(minInc < 0 && memberIncrement < minInc) currIncrement = minInc + X[e];
                        else if(minInc > 0 && memberIncrement < 0) currIncrement = 0 + X[e];
                    }
                    analysis[(size_t) y * nEns + validEns[e]] = ensMean + currIncrement;
                }
            }
            free(lY); free(lYc); free(Cm); free(dd);
        }
        free(nb); free(cand);
        free(Pinv); free(P); free(S); free(val); free(vec); free(W); free(w); free(X); free(X_corr); free(X1); free(X_corr1);
    }
    cells_free(&cells);
    free(gY); free(gY_corr); free(gYhat); free(ok); free(validEns);
    pts_free(&bp);
    pts_free(&op);
    return 0;
}

/* ------------------------------------------------------------------ neighbourhood_search / calc_gradient ---- */
/* neighbourhood_search.cpp:7-113. apply may be NULL (the reference's empty ivec2: every point is treated) */
int orc_neighbourhood_search(const float* array, const float* search, int ny, int nx, int halfwidth, float tmin, float tmax, float delta,
                             const int* apply, float* output) {
    if(tmin > tmax) FAIL(1, "Search_target_min must be smaller than search_target_max");
    if(halfwidth < 0) FAIL(1, "halfwidth must be positive");
    for(int y = 0; y < ny; y++)
        for(int x = 0; x < nx; x++) {
            size_t o = (size_t) y * nx + x;
            float nearest_target = NAN, accum_temp = 0;
            int iy = 0, ix = 0, counter = 0;
            if(!is_valid(search[o])) { output[o] = array[o]; continue; }                 /* :40-44 */
            if(apply && apply[o] == 0) { output[o] = array[o]; continue; }               /* :46-50 (an int is always "valid") */
            for(int yy = y - halfwidth < 0 ? 0 : y - halfwidth; yy <= (y + halfwidth > ny - 1 ? ny - 1 : y + halfwidth); yy++)
                for(int xx = x - halfwidth < 0 ? 0 : x - halfwidth; xx <= (x + halfwidth > nx - 1 ? nx - 1 : x + halfwidth); xx++) {
                    size_t i = (size_t) yy * nx + xx;
                    if(!is_valid(search[i]) || !is_valid(array[i])) continue;
                    if(!apply || apply[o] == 1) {
                        if(search[i] >= tmin && search[i] <= tmax) { counter++; accum_temp = accum_temp + array[i]; }
                        else if(counter > 0) continue;
                        else if(fabsf(search[i] - search[o]) >= delta) {
                            if(!is_valid(nearest_target)) { nearest_target = search[i]; iy = yy; ix = xx; }
                            else {
                                float curr = fminf(fabsf(search[i] - tmin), fabsf(search[i] - tmax));
                                float best = fminf(fabsf(nearest_target - tmin), fabsf(nearest_target - tmax));
                                if(curr < best) { nearest_target = search[i]; iy = yy; ix = xx; }
                            }
                        }
                    }
                }
            if(counter > 0) output[o] = accum_temp / counter;
            else if(is_valid(nearest_target)) output[o] = array[(size_t) iy * nx + ix];
            else output[o] = array[o];
        }
    return 0;
}

/* calc_gradient.cpp:6-126. gradient_type 0 = MinMax, 10 = LinearRegression (gridpp.h:126-129) */
int orc_calc_gradient(const float* base, const float* values, int ny, int nx, int gradient_type, int halfwidth, int num_min, float min_range,
                      float default_gradient, float* output) {
    if(halfwidth <= 0) FAIL(1, "Halwidth cannot be <= 0; must be positive integer");
    if(is_valid(min_range) && min_range < 0) FAIL(1, "min_range must be >= 0");
    if(num_min < 0) FAIL(1, "num_min must be >= 0");
    if(ny == 0) FAIL(1, "base input has no size");
    size_t n = (size_t) ny * nx;
    for(size_t i = 0; i < n; i++) output[i] = default_gradient;
    if(gradient_type == 0) {
        for(int y = 0; y < ny; y++)
            for(int x = 0; x < nx; x++) {
                float current_max = NAN, current_min = NAN;
                size_t imax = 0, imin = 0;
                int count = 0;
                for(int yy = y - halfwidth < 0 ? 0 : y - halfwidth; yy <= (y + halfwidth > ny - 1 ? ny - 1 : y + halfwidth); yy++)
                    for(int xx = x - halfwidth < 0 ? 0 : x - halfwidth; xx <= (x + halfwidth > nx - 1 ? nx - 1 : x + halfwidth); xx++) {
                        size_t i = (size_t) yy * nx + xx;
                        float current_base = base[i];
                        if(!is_valid(current_base) || !is_valid(values[i])) continue;
                        if(!is_valid(current_max) || current_base > current_max) { current_max = current_base; imax = i; }
                        if(!is_valid(current_min) || current_base < current_min) { current_min = current_base; imin = i; }
                        count++;
                    }
                size_t o = (size_t) y * nx + x;
                if(count < num_min) output[o] = default_gradient;
                else if(!is_valid(current_max) || !is_valid(current_min)) output[o] = default_gradient;
                else if(fabsf(current_max - current_min) <= min_range) output[o] = default_gradient;   /* abs(float): the float overload */
                else {
                    float diffBase = current_max - current_min;
                    float diffValues = values[imax] - values[imin];
                    output[o] = diffValues / diffBase;
                }
            }
    }
    else if(gradient_type == 10) {
        float* f[5];
        float* m[5];
        for(int k = 0; k < 5; k++) { f[k] = malloc(sizeof(float) * n); m[k] = malloc(sizeof(float) * n); }
        for(size_t i = 0; i < n; i++) {
            int ok = is_valid(base[i]) && is_valid(values[i]);
            f[0][i] = ok ? base[i] : NAN;
            f[1][i] = ok ? values[i] : NAN;
            f[2][i] = ok ? (float) pow(base[i], 2) : NAN;
            f[3][i] = ok ? base[i] * values[i] : NAN;
            f[4][i] = ok ? 1 : 0;
        }
        int rc = 0;
        for(int k = 0; k < 5 && rc == 0; k++) rc = orc_neighbourhood(f[k], ny, nx, halfwidth, k < 4 ? MEAN : SUM, m[k], NULL);
        for(size_t i = 0; i < n && rc == 0; i++) {
            float meanX = m[0][i], meanY = m[1][i], meanXX = m[2][i], meanXY = m[3][i], count = m[4][i];
            output[i] = default_gradient;
            if(count >= num_min && is_valid(meanXX) && is_valid(meanXY) && is_valid(meanX) && meanXX - meanX * meanX != 0) {
                int valid_range = 1;
                if(is_valid(min_range)) {
                    float range = sqrtf(meanXX - meanX * meanX);
                    if(!is_valid(range)) valid_range = 0;
                    else if(range < min_range) valid_range = 0;
                }
                if(valid_range) output[i] = (meanXY - meanX * meanY) / (meanXX - meanX * meanX);
            }
        }
        for(int k = 0; k < 5; k++) { free(f[k]); free(m[k]); }
        return rc;
    }
    return 0;
}
