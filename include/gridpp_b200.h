/* gridpp_b200.h -- C ABI of the B200-native gridpp hot path (libgridpp_b200.so).
 *
 * This is the drop-in boundary. Every entry point states which reference interface (file:line under the
 * metno/gridpp tree) it replaces. Signatures are plain C: pointers, sizes, PODs, integer status. No torch,
 * no C++ types. The C++ `gridpp::` host layer (include/gridpp.h in this repo) and the Python mirror
 * (gridpp_b200/) are thin callers of these functions.
 *
 * Conventions
 *   - Return value: GPP_OK or an error code; gpp_last_error() returns the message of the last failure on the
 *     calling thread. GPP_ERR_INVALID_ARGUMENT corresponds to the reference's std::invalid_argument,
 *     everything else to std::runtime_error (swig/gridpp.i:21-40 maps them to ValueError / RuntimeError).
 *   - `*_host` entry points take HOST buffers, do the host<->device copies themselves and synchronise
 *     before returning (this is what the reference-facing API calls). `*_device` entry points take DEVICE
 *     buffers of the current device, enqueue on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream) and do not synchronise.
 *   - All fields are dense row-major float32; missing value = NaN (gridpp.h:49); a value is "valid" when it
 *     is neither NaN nor +-inf (util.cpp:16-18).
 *   - There is NO CPU fallback: every compute entry point fails with GPP_ERR_CUDA when no device is usable.
 */
#ifndef GRIDPP_B200_H
#define GRIDPP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPP_OK 0
#define GPP_ERR_INVALID_ARGUMENT 1
#define GPP_ERR_RUNTIME 2
#define GPP_ERR_NOT_IMPLEMENTED 3
#define GPP_ERR_CUDA 4

/* gridpp::CoordinateType, gridpp.h:120-123 (numeric values kept) */
#define GPP_GEODETIC 0
#define GPP_CARTESIAN 1

/* gridpp::Statistic, gridpp.h:88-100 (numeric values kept) */
#define GPP_MEAN 0
#define GPP_MIN 10
#define GPP_MEDIAN 20
#define GPP_MAX 30
#define GPP_QUANTILE 40
#define GPP_STD 50
#define GPP_VARIANCE 60
#define GPP_SUM 70
#define GPP_COUNT 80
#define GPP_RANDOMCHOICE 90

/* Structure-function families, src/api/structure.cpp */
#define GPP_STRUCT_BARNES 0   /* structure.cpp:143-282 */
#define GPP_STRUCT_CRESSMAN 1 /* structure.cpp:287-312 */
#define GPP_STRUCT_SOAR 2     /* structure.cpp:317-460 */
#define GPP_STRUCT_TOAR 3     /* structure.cpp:467-610 */
#define GPP_STRUCT_POWERLAW 4 /* structure.cpp:618-757 */
#define GPP_STRUCT_LINEAR 5   /* structure.cpp:765-904 */

/* One non-spatial structure function: rho = rho_h(dist/h) * rho_v(delev/v) * rho_w(dlaf/w), zero beyond
 * loc_dist. POD replacement for the virtual class gridpp::StructureFunction (gridpp.h:2069-2343). */
typedef struct gpp_structure_term {
    int type;       /* GPP_STRUCT_* */
    float h, v, w;  /* length scales (Linear: minimum correlations) */
    float min_rho;  /* truncation correlation (structure.cpp:5 default 0.0013) */
    float loc_dist; /* localization_distance(), computed on the host exactly as the reference does */
} gpp_structure_term;

/* n_terms == 1: a plain structure function. n_terms == 3: gridpp::MultipleStructure(h, v, w)
 * (structure.cpp:90-138): term[0] sees only the horizontal separation, term[1] only the elevation
 * difference, term[2] only the land-area-fraction difference; localization comes from term[0].
 * has_cv != 0: wrapped in gridpp::CrossValidation(structure, cv_dist) (structure.cpp:909-944): background
 * correlations are zeroed for observations closer than cv_dist. */
typedef struct gpp_structure {
    int n_terms;
    gpp_structure_term term[3];
    int has_cv;
    float cv_dist;
} gpp_structure;

/* ---------------------------------------------------------------- library ---------------------------- */
const char* gpp_version(void);                 /* gridpp::version(), gridpp.cpp:8-10 */
const char* gpp_last_error(void);              /* message of the last failing call on this thread */
int gpp_device_count(int* count);              /* number of usable CUDA devices (0 is not an error) */
int gpp_set_device(int device);                /* device used by subsequent calls of this thread */
int gpp_device_synchronize(void);

/* ---------------------------------------------------------------- structure functions ---------------- */
/* BarnesStructure(h,v,w,hmax) structure.cpp:143-167, CressmanStructure :287-297, SoarStructure :317-340,
 * ToarStructure :467-490, PowerlawStructure :618-641, LinearStructure :765-788. hmax = NaN -> default
 * min_rho. Validation errors match the constructors (GPP_ERR_INVALID_ARGUMENT). */
int gpp_structure_init(gpp_structure* out, int type, float h, float v, float w, float hmax);
/* The constant-scale case of <Family>Structure(Grid, vec2 h, vec2 v, vec2 w, min_rho) with (1, 1) fields
 * (structure.cpp:168-176 and siblings): scales h, v, w with an explicit truncation correlation. */
int gpp_structure_init_min_rho(gpp_structure* out, int type, float h, float v, float w, float min_rho);
/* MultipleStructure(structure_h, structure_v, structure_w), structure.cpp:90-94. Each input must be a
 * single-term structure. */
int gpp_structure_multiple(gpp_structure* out, const gpp_structure* sh, const gpp_structure* sv, const gpp_structure* sw);
/* CrossValidation(structure, dist), structure.cpp:909-915 */
int gpp_structure_cross_validation(gpp_structure* out, const gpp_structure* in, float dist);
/* StructureFunction::corr / corr_background between two points given as (x,y,z,elev,laf) and
 * localization_distance(): evaluated ON THE DEVICE with the same code the OI kernels inline.
 * p1 and p2 hold n points each as 5 floats (x, y, z, elev, laf); background != 0 -> corr_background. */
int gpp_structure_corr_host(const gpp_structure* s, const float* p1, const float* p2, int n, int background, float* out);

/* ---------------------------------------------------------------- points / grid index ---------------- */
/* Replaces gridpp::Points (points.cpp:9-31) + gridpp::KDTree (kdtree.cpp:6-16); a gridpp::Grid is the same
 * object over its Y*X flattened nodes (grid.cpp:12-55). Host arrays are copied; elevs/lafs may be NULL
 * (filled with NaN, points.cpp:23-30). Coordinates are converted on the host with the reference's formula
 * (util.cpp:583-615) so x/y/z match the reference bit for bit; invalid coordinates -> INVALID_ARGUMENT. */
typedef struct gpp_points gpp_points;
int gpp_points_create(const float* lats, const float* lons, const float* elevs, const float* lafs, int n,
                      int coordinate_type, gpp_points** out);
void gpp_points_destroy(gpp_points* p);
int gpp_points_size(const gpp_points* p);
int gpp_points_coordinate_type(const gpp_points* p);
/* Declares the points to be the row-major flattening of an ny x nx grid (what gridpp::Grid is, grid.cpp:12-55).
 * Purely a traversal hint: the OI kernels then walk 4 x 4 tiles instead of 16-point row segments, which raises
 * the reuse of solved systems. ny * nx must equal the number of points. Results do not depend on it. */
int gpp_points_set_shape(gpp_points* p, int ny, int nx);
/* KDTree::get_x/get_y/get_z (kdtree.cpp:213-221): copies n floats each; any pointer may be NULL */
int gpp_points_get_xyz(const gpp_points* p, float* x, float* y, float* z);
/* gridpp::convert_coordinates(lats, lons, type, x, y, z), util.cpp:583-615: host arithmetic, the same the point
 * sets use (Geodetic -> earth-centred metres, Cartesian -> x = lon, y = lat, z = 0). Invalid coordinates ->
 * INVALID_ARGUMENT (util.cpp:596-600). Needs no device. */
int gpp_convert_coordinates(const float* lats, const float* lons, int n, int coordinate_type, float* x, float* y, float* z);

/* KDTree::get_closest_neighbours(lat,lon,1) for nq query points at once (kdtree.cpp:82-106; the per-point
 * loop of nearest.cpp:7-222). out_index[q] = index of the nearest point, -1 if none. Ties resolve to the
 * lowest index (Boost leaves them unspecified). */
int gpp_points_nearest_host(const gpp_points* p, const float* qlats, const float* qlons, int nq,
                            int include_match, int* out_index);
/* KDTree::get_neighbours / get_neighbours_with_distance / get_num_neighbours (kdtree.cpp:18-80) for nq
 * query points: a point is returned when it lies STRICTLY inside the box [q-r, q+r]^3 and its straight-line
 * distance is <= r (and > 0 unless include_match). Results for query q are written at
 * out_index[q*capacity ...] in ascending point index, out_count[q] holds the TOTAL number of neighbours
 * (which may exceed capacity; only the first `capacity` are stored). out_index/out_dist may be NULL with
 * capacity 0 to only count. */
int gpp_points_neighbours_host(const gpp_points* p, const float* qlats, const float* qlons, const float* radii,
                               int nq, int include_match, int capacity, int* out_index, float* out_dist,
                               int* out_count);
/* KDTree::get_closest_neighbours(lat, lon, num) (kdtree.cpp:82-103) for nq query points; out_index is
 * nq*num, padded with -1; sorted by (distance, index). */
int gpp_points_closest_host(const gpp_points* p, const float* qlats, const float* qlons, int nq, int num,
                            int include_match, int* out_index);

/* gridpp::nearest(...) (nearest.cpp:7-222): for nq output points take the value of the nearest input point.
 * ivalues is n_fields x n_in (row-major), out is n_fields x nq. An empty input set yields NaN. */
int gpp_nearest_host(const gpp_points* ipoints, const float* qlats, const float* qlons, int nq,
                     const float* ivalues, int n_fields, float* out);

/* ---------------------------------------------------------------- optimal interpolation -------------- */
/* Opaque, reusable observation-side state for OI: the cell index over the valid observations with their
 * innovations and variance ratios resident on the device. */
typedef struct gpp_oi_obs gpp_oi_obs;

/* gridpp::optimal_interpolation_full(Points...) oi.cpp:138-341 (and, with bvariance = NULL,
 * bvariance_at_points = NULL and obs_variance = the variance ratios, gridpp::optimal_interpolation(Points...)
 * oi.cpp:89-136; the Grid overloads oi.cpp:26-87,342-412 flatten to this). All pointers are HOST memory.
 *   bpoints            background points (a flattened grid), size nB
 *   background[nB]     background field
 *   bvariance[nB]      background variance or NULL (= 1)
 *   opoints            observation points, size nS
 *   pobs, obs_variance, pbackground [nS]; bvariance_at_points [nS] or NULL (= 1)
 *   max_points         0 = unlimited
 *   analysis[nB]       output; analysis_variance[nB] output or NULL
 * Argument checks and their order follow oi.cpp:151-186. */
int gpp_optimal_interpolation_host(const gpp_points* bpoints, const float* background, const float* bvariance,
                                   const gpp_points* opoints, const float* pobs, const float* obs_variance,
                                   const float* pbackground, const float* bvariance_at_points,
                                   const gpp_structure* structure, int max_points, int allow_extrapolation,
                                   float* analysis, float* analysis_variance);

/* Device-resident form of the same call, split so that the observation side is prepared once and the
 * (row-sharded) background side streams through: build the observation state from HOST observation arrays,
 * then analyse background points [first, first+count) of `bpoints` reading d_background[first...] and writing
 * d_analysis[first...] (device pointers indexed like the full field). d_bvariance / d_analysis_variance may
 * be NULL. */
int gpp_oi_obs_create(const gpp_points* opoints, const float* pobs, const float* obs_variance,
                      const float* pbackground, const float* bvariance_at_points, const gpp_structure* structure,
                      gpp_oi_obs** out);
void gpp_oi_obs_destroy(gpp_oi_obs* obs);
int gpp_optimal_interpolation_device(const gpp_points* bpoints, int first, int count, const float* d_background,
                                     const float* d_bvariance, const gpp_oi_obs* obs,
                                     const gpp_structure* structure, int max_points, int allow_extrapolation,
                                     float* d_analysis, float* d_analysis_variance, void* stream);

/* The same call with an explicit launch workspace instead of the slots the observation state owns (4 of them, used
 * round-robin: enough for one stream or the library's own two-stream pipeline, not for more than 4 concurrent analyses
 * sharing one state). d_workspace: gpp_oi_workspace_bytes() bytes of device memory, zero-filled ONCE before its first
 * use (the kernels leave it ready for the next launch); one workspace per launch in flight. With it the register path
 * (symmetric structure function, max_points 1..30) performs no allocation and no memset: a step is one kernel launch. */
size_t gpp_oi_workspace_bytes(void);
int gpp_optimal_interpolation_device_ws(const gpp_points* bpoints, int first, int count, const float* d_background,
                                        const float* d_bvariance, const gpp_oi_obs* obs,
                                        const gpp_structure* structure, int max_points, int allow_extrapolation,
                                        float* d_analysis, float* d_analysis_variance, void* d_workspace,
                                        size_t workspace_bytes, void* stream);

/* The same analysis spread over several devices by ONE process: the rows of the background grid (single points of a point
 * set) are split into contiguous blocks, one per device (n_devices <= 0: every visible device); each device runs
 * gpp_optimal_interpolation_host on its block from its own host thread against its own copy of the observation table.
 * Grid points are independent (oi.cpp:221-338): no exchange, results identical to the single-device call bit for bit.
 * (One process per GPU -- MPI, torch.distributed -- shards the same way with gpp_optimal_interpolation_device.) */
int gpp_optimal_interpolation_multi_gpu_host(int n_devices, const gpp_points* bpoints, const float* background, const float* bvariance,
                                             const gpp_points* opoints, const float* pobs, const float* obs_variance,
                                             const float* pbackground, const float* bvariance_at_points,
                                             const gpp_structure* structure, int max_points, int allow_extrapolation,
                                             float* analysis, float* analysis_variance);

/* gridpp::optimal_interpolation_ensi(Points...) oi_ensi.cpp:114-568 (the Grid overload :33-112 flattens to
 * this). background and analysis are nB x nE (member fastest), pbackground is nS x nE. HOST memory. */
int gpp_optimal_interpolation_ensi_host(const gpp_points* bpoints, const float* background, int nE,
                                        const gpp_points* opoints, const float* pobs, const float* psigmas,
                                        const float* pbackground, const gpp_structure* structure, int max_points,
                                        int allow_extrapolation, float* analysis, int* num_skipped);

/* Device-resident form of optimal_interpolation_ensi, split like the deterministic one. The observation side is built
 * once from HOST arrays (oi_ensi.cpp:163-178: perturbations of pbackground about its ensemble mean; :232: only
 * observations with a valid value enter). member_valid: nE flags (0 = the member has an invalid value somewhere in the
 * background and is left untouched, oi_ensi.cpp:187-201), or NULL = every member is valid;
 * gpp_ensi_valid_members_device computes the flags of a device-resident nB x nE background (it synchronises `stream`).
 * gpp_optimal_interpolation_ensi_device analyses points [first, first+count) of `bpoints`: d_background and d_analysis
 * are nB x nE device arrays indexed like the full field (they may be the same array); one kernel launch, no
 * synchronisation, no allocation (max_points = 0 with more than 64 observations in reach adds a counting pass that
 * synchronises). d_num_skipped: device int incremented once per point skipped for a numerically bad Pinv
 * (oi_ensi.cpp:386-390), or NULL. At most 16 analyses sharing one observation state may be in flight at once. */
typedef struct gpp_ensi_obs gpp_ensi_obs;
int gpp_ensi_obs_create(const gpp_points* opoints, const float* pobs, const float* psigmas, const float* pbackground,
                        int nE, const int* member_valid, const gpp_structure* structure, gpp_ensi_obs** out);
void gpp_ensi_obs_destroy(gpp_ensi_obs* obs);
int gpp_ensi_valid_members_device(const float* d_background, long long n_points, int nE, int* member_valid, void* stream);
int gpp_optimal_interpolation_ensi_device(const gpp_points* bpoints, int first, int count, const float* d_background,
                                          int nE, const gpp_ensi_obs* obs, const gpp_structure* structure,
                                          int max_points, int allow_extrapolation, float* d_analysis,
                                          int* d_num_skipped, void* stream);

/* Spatially varying structure functions: gridpp::BarnesStructure(Grid, vec2 h, vec2 v, vec2 w, min_rho)
 * structure.cpp:168-184 and the Soar (:342), Toar (:492), Powerlaw (:643), Linear (:790) siblings. h, v, w hold one
 * value per node of `grid` (row-major); a point uses the scales of its nearest node (structure.cpp:189-199). The grid
 * must outlive the field. */
typedef struct gpp_structure_field gpp_structure_field;
int gpp_structure_field_create(const gpp_points* grid, const float* h, const float* v, const float* w, gpp_structure_field** out);
void gpp_structure_field_destroy(gpp_structure_field* field);
int gpp_structure_field_lookup_host(const gpp_structure_field* field, const float* lats, const float* lons, int n,
                                    float* h, float* v, float* w);
/* <Family>Structure::localization_distance(Point) for the spatial form (structure.cpp:271-282) */
int gpp_structure_field_localization_distance(const gpp_structure_field* field, int type, float min_rho, float lat, float lon,
                                              float* out);
/* gridpp::optimal_interpolation / optimal_interpolation_full (oi.cpp:26-412) with such a structure function
 * (type: GPP_STRUCT_BARNES, _SOAR, _TOAR, _POWERLAW or _LINEAR). Other arguments as gpp_optimal_interpolation_host. */
int gpp_optimal_interpolation_spatial_host(const gpp_points* bpoints, const float* background, const float* bvariance,
                                           const gpp_points* opoints, const float* pobs, const float* obs_variance,
                                           const float* pbackground, const float* bvariance_at_points, int structure_type,
                                           const gpp_structure_field* field, float min_rho, int max_points,
                                           int allow_extrapolation, float* analysis, float* analysis_variance);

/* ---------------------------------------------------------------- neighbourhood filters -------------- */
/* gridpp::neighbourhood(vec2, halfwidth, statistic) neighbourhood.cpp:28-242: NaN-aware statistic over the
 * (2*halfwidth+1)^2 window CLIPPED to the domain (neighbourhood.cpp:104-107,160-167); a window without valid values gives
 * NaN (Count gives 0). Mean / Sum / Count / Min / Max run the TMA-staged stencil kernels; Std / Variance are two Mean
 * filters (:211-235); Median / RandomChoice gather the window (gpp_neighbourhood_brute_force_*). */
int gpp_neighbourhood_host(const float* input, int ny, int nx, int halfwidth, int statistic, float* output);
/* Device form with explicit row window, for row-tiled multi-GPU use: d_input holds n_rows_in rows (a tile
 * plus whatever halo rows exist; the domain is taken to be exactly these rows), output rows
 * [row0, row0 + n_rows_out) are written to d_output[0 .. n_rows_out*nx). */
int gpp_neighbourhood_device(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out,
                             int halfwidth, int statistic, float* d_output, void* stream);

/* Halo exchange of a row-tiled field as one kernel over peer memory. d_buf holds this rank's tile with room for the halos:
 * [halfwidth rows | rows rows | halfwidth rows] x nx. d_from_above points at the LAST `halfwidth` tile rows of the upper
 * neighbour, d_from_below at the FIRST `halfwidth` tile rows of the lower neighbour -- device pointers of another GPU mapped
 * into this process (symmetric memory / CUDA IPC between ranks, peer access inside one process); NULL at a domain edge.
 * The caller orders the call after the neighbours' tiles are complete (a barrier on the exchange's signal pad) and before
 * gpp_neighbourhood_device / gpp_neighbourhood_quantile_fast_device on the tile with its halo. */
int gpp_halo_pull_device(float* d_buf, int rows, int nx, int halfwidth, const float* d_from_above, const float* d_from_below,
                         void* stream);

/* gridpp::neighbourhood_quantile_fast(vec2, quantile | vec2 quantile, halfwidth, thresholds)
 * neighbourhood.cpp:296-409. quantile_field (ny x nx) may be NULL, then `quantile` applies everywhere.
 * thresholds is a HOST array in both forms. */
int gpp_neighbourhood_quantile_fast_host(const float* input, int ny, int nx, float quantile,
                                         const float* quantile_field, int halfwidth, const float* thresholds,
                                         int num_thresholds, float* output);
int gpp_neighbourhood_quantile_fast_device(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out,
                                           float quantile, const float* d_quantile_field, int halfwidth,
                                           const float* thresholds, int num_thresholds, float* d_output,
                                           void* stream);

/* The ensemble (vec3) forms. input is ny x nx x ne, member fastest.
 * gridpp::neighbourhood(vec3, halfwidth, statistic) neighbourhood.cpp:12-27: calc_statistic over the members of every
 * cell (util.cpp:19-110), then the 2-D filter. */
int gpp_neighbourhood_ens_host(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float* output);
int gpp_neighbourhood_ens_device(const float* d_input, int ny, int nx, int ne, int halfwidth, int statistic,
                                 float* d_output, void* stream);
/* gridpp::neighbourhood_quantile_fast(vec3, quantile | vec2 quantile, halfwidth, thresholds) neighbourhood.cpp:411-527. */
int gpp_neighbourhood_quantile_fast_ens_host(const float* input, int ny, int nx, int ne, float quantile,
                                             const float* quantile_field, int halfwidth, const float* thresholds,
                                             int num_thresholds, float* output);
int gpp_neighbourhood_quantile_fast_ens_device(const float* d_input, int ny, int nx, int ne, float quantile,
                                               const float* d_quantile_field, int halfwidth, const float* thresholds,
                                               int num_thresholds, float* d_output, void* stream);

/* gridpp::get_neighbourhood_thresholds(vec2 | vec3, num_thresholds) neighbourhood.cpp:243-295 (both forms pool every
 * valid value, so the field is passed flattened): up to num_thresholds values are written, *num_out tells how many. */
int gpp_get_neighbourhood_thresholds_host(const float* input, long long n_values, int num_thresholds, float* thresholds,
                                          int* num_out);

/* ---------------------------------------------------------------- consumers of the point index -------- */
/* All take point sets (a Grid is its flattened nodes; results are flat in the same order) and HOST arrays.
 * gridpp::gridding (gridding.cpp:6-61): out[o] = statistic of values[i] over the input points i within `radius` of output
 * location o (KDTree::get_neighbours: strictly inside the box, straight distance <= radius; values taken in ascending
 * point index); fewer than min_num (> 0) neighbours -> missing. */
int gpp_gridding_host(const gpp_points* opoints, const gpp_points* ipoints, const float* values, float radius, int min_num,
                      int statistic, float* output);
/* gridpp::gridding_nearest (gridding.cpp:63-131): every input point is assigned to its nearest output location; out[o] =
 * statistic of the values assigned to o (in ascending point index), missing when none (or fewer than min_num > 0). */
int gpp_gridding_nearest_host(const gpp_points* opoints, const gpp_points* ipoints, const float* values, int min_num,
                              int statistic, float* output);
/* gridpp::count (count.cpp:6-66, all four overloads): out[o] = number of input points within `radius` of output location o */
int gpp_count_host(const gpp_points* ipoints, const gpp_points* opoints, float radius, float* output);
/* gridpp::distance (distance.cpp:6-120): out[o] = largest KDTree::calc_distance from output location o to its `num` closest
 * input points (num <= 16). query_first: the argument order of calc_distance in the overload (distance.cpp:21,111 pass the
 * output location first; :52,83 the input point) -- it only matters for the rounding of the great-circle formula. */
int gpp_distance_host(const gpp_points* ipoints, const gpp_points* opoints, int num, int query_first, float* output);
/* gridpp::fill (fill.cpp:6-43): outside == 0: cells of `igrid` within radii[i] of point i take `value`, the others keep
 * `input`; outside != 0: the reverse. */
int gpp_fill_host(const gpp_points* igrid, const float* input, const gpp_points* points, const float* radii, float value,
                  int outside, float* output);
/* gridpp::fill_missing (fill.cpp:44-134): missing values replaced by the mean of the linear interpolations along x and y */
int gpp_fill_missing_host(const float* values, int ny, int nx, float* output);
/* gridpp::doping_square / doping_circle (doping.cpp:5-93): observations[i] is written into the square of half-width
 * halfwidth[i] cells around the node nearest to point i / into the nodes within radii[i] of it, unless the elevation of the
 * node differs from the point's by more than max_elev_diff (NaN: no check); where several points claim a cell the one with
 * the highest index wins (the reference writes them in index order). doping_square needs the grid shape
 * (gpp_points_set_shape). */
int gpp_doping_square_host(const gpp_points* igrid, const float* background, const gpp_points* points, const float* observations,
                           const int* halfwidth, float max_elev_diff, float* output);
int gpp_doping_circle_host(const gpp_points* igrid, const float* background, const gpp_points* points, const float* observations,
                           const float* radii, float max_elev_diff, float* output);

/* ---------------------------------------------------------------- row statistics ---------------------- */
/* gridpp::calc_statistic(vec2, statistic) util.cpp:208-215 (and, with n_rows = 1, the vec form :19-110): `array` holds
 * n_rows rows of row_length values; out[r] = statistic of the valid values of row r (Mean, Min, Median, Max, Std,
 * Variance, Sum, Count; float accumulation in element order as the reference; RandomChoice picks a valid value of the row
 * with a hash of the call count and the row -- the reference draws from the C library's rand(), whose sequence is not
 * reproduced). The device form is what the ensemble filters chain (member fastest). */
int gpp_calc_statistic_host(const float* array, long long n_rows, int row_length, int statistic, float* out);
int gpp_calc_statistic_device(const float* d_array, long long n_rows, int row_length, int statistic, float* d_out, void* stream);
/* gridpp::neighbourhood_brute_force(vec2 | vec3, halfwidth, statistic) and gridpp::neighbourhood_quantile(vec2 | vec3,
 * quantile, halfwidth) (statistic = GPP_QUANTILE), neighbourhood.cpp:528-630: the clipped window of every pixel (all members
 * for a vec3; ne = 1 for a vec2) is reduced with calc_statistic / calc_quantile. The device form takes a row window like
 * gpp_neighbourhood_device. Also what gpp_neighbourhood_* run for Median and RandomChoice (neighbourhood.cpp:237-238). */
int gpp_neighbourhood_brute_force_host(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float quantile, float* output);
int gpp_neighbourhood_brute_force_device(const float* d_input, int n_rows_in, int nx, int ne, int row0, int n_rows_out, int halfwidth,
                                         int statistic, float quantile, float* d_output, void* stream);
/* gridpp::calc_quantile(vec, q) util.cpp:111-178, (vec2, q) :179-186 and (vec3, vec2 q) :187-207: quantile_rows (one level
 * per row) may be NULL, then `quantile` applies to every row. A level outside [0, 1] -> INVALID_ARGUMENT. */
int gpp_calc_quantile_host(const float* array, long long n_rows, int row_length, float quantile, const float* quantile_rows, float* out);
/* gridpp::interpolate(vec x, iX, iY) util.cpp:377-431 (n = 1: the scalar form): piecewise-linear with the reference's
 * plateau rules (get_lower_index / get_upper_index, util.cpp:339-376). */
int gpp_interpolate_host(const float* x, long long n, const float* iX, const float* iY, int m, float* out);

/* ---------------------------------------------------------------- instrumentation -------------------- */
/* Number of kernels this library has launched on the calling process so far (bench.py's gpu_launches). */
unsigned long long gpp_kernel_launch_count(void);
/* Measured fp64 FMA throughput of the current device in TFLOP/s (2 flops per FMA): the roofline denominator of
 * the OI kernels, whose elimination runs on the fp64 CUDA cores. */
int gpp_measure_fp64_fma_peak(double* tflops);

/* ---- the "multi" EnSI variants and the static correlations between two point sets. Points overloads
 * (the Grid overloads of the reference, oi_ensi_multi.cpp:33-327, only flatten the grid: pass a flattened-grid gpp_points).
 * Fields are member-fastest: background* [L][nE], pobs (ebe / ebesc) and pbackground* [S][nE]. analysis is [L][nE]. */

/* gridpp::optimal_interpolation_ensi_multi_ebe(Points, ...), src/api/oi_ensi_multi.cpp:329-627 */
int gpp_optimal_interpolation_ensi_multi_ebe_host(const gpp_points* bpoints, const float* bratios, const float* background,
                                                  const float* background_corr, int nE, const gpp_points* opoints, const float* pobs,
                                                  const float* pratios, const float* pbackground, const float* pbackground_corr,
                                                  const gpp_structure* structure, int max_points, int allow_extrapolation, float* analysis);
/* gridpp::optimal_interpolation_ensi_multi_ebesc(Points, ...), src/api/oi_ensi_multi.cpp:630-859 */
int gpp_optimal_interpolation_ensi_multi_ebesc_host(const gpp_points* bpoints, const float* bratios, const float* background, int nE,
                                                    const gpp_points* opoints, const float* pobs, const float* pratios,
                                                    const float* pbackground, const gpp_structure* structure, int max_points,
                                                    int allow_extrapolation, float* analysis);
/* gridpp::optimal_interpolation_ensi_multi_utem(Points, ...), src/api/oi_ensi_multi.cpp:862-1311. pobs is [S].
 * *num_skipped (may be NULL) receives the number of points left at their background because rcond(Pinv) <= 0 (:1106-1110). */
int gpp_optimal_interpolation_ensi_multi_utem_host(const gpp_points* bpoints, const float* bratios, const float* background,
                                                   const float* background_corr, int nE, const gpp_points* opoints, const float* pobs,
                                                   const float* pratios, const float* pbackground, const float* pbackground_corr,
                                                   const gpp_structure* structure, int max_points, int allow_extrapolation, float* analysis,
                                                   int* num_skipped);
/* gridpp::staticcorr_points(points, knots, structure, max_points), src/api/corr_points.cpp:26-131. output is [L][K]. */
int gpp_staticcorr_points_host(const gpp_points* points, const gpp_points* knots, const gpp_structure* structure, int max_points,
                               float* output);

/* ---- the last two window filters of SURVEY.md 8(f)#4 */
#define GPP_GRADIENT_MINMAX 0                /* gridpp::MinMax, gridpp.h:126-129 */
#define GPP_GRADIENT_LINEAR_REGRESSION 10    /* gridpp::LinearRegression */
/* gridpp::neighbourhood_search(array, search_array, halfwidth, min, max, delta, apply_array), src/api/neighbourhood_search.cpp:7-113.
 * apply_array: ny x nx ints (1 = correct this point), or NULL for the reference's empty ivec2 (every point). */
int gpp_neighbourhood_search_host(const float* array, const float* search_array, int ny, int nx, int halfwidth, float search_target_min,
                                  float search_target_max, float search_delta, const int* apply_array, float* output);
/* gridpp::calc_gradient(base, values, gradient_type, halfwidth, num_min, min_range, default_gradient), src/api/calc_gradient.cpp:6-126 */
int gpp_calc_gradient_host(const float* base, const float* values, int ny, int nx, int gradient_type, int halfwidth, int num_min, float min_range,
                           float default_gradient, float* output);

#ifdef __cplusplus
}
#endif
#endif
