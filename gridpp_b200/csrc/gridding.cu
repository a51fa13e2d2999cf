// The radius-query consumers of the point index: gridpp::gridding / gridding_nearest
// (src/api/gridding.cpp:6-131), gridpp::count (count.cpp:6-66), gridpp::distance (distance.cpp:6-120), gridpp::fill /
// fill_missing (fill.cpp:6-134) and gridpp::doping_square / doping_circle (doping.cpp:5-93).
//
// All of them are "for every output location, look up neighbours in a point set and reduce": the lookups run on the
// device bucket grid of gpp_points (points.cu) with the reference's exact predicate (strictly inside the box, straight
// distance <= radius, kdtree.cpp:39-62,247-260); the reductions are the row statistics of rowstats.cuh.
#include "points.cuh"
#include "rowstats.cuh"

#include <thrust/binary_search.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/scan.h>
#include <thrust/sort.h>

#include <cstring>

using namespace gpp;
using namespace gpp::rowstats;

namespace {

struct IndexView {
    CellGeom geom;
    const int* cell_start;
    const int* order;
    const float *sx, *sy, *sz;
};
IndexView view_of(const gpp_points* p) {
    return IndexView{p->index.geom, p->index.cell_start.ptr, p->index.order.ptr, p->index.sx.ptr, p->index.sy.ptr, p->index.sz.ptr};
}

// KDTree::get_neighbours (kdtree.cpp:39-62, within_radius :247-260) for one query: f(point index) for every point STRICTLY
// inside the box [q - r, q + r]^3 whose straight distance is <= r; returns how many
template <class F>
__device__ __forceinline__ int for_each_neighbour(const IndexView& ix, float x, float y, float z, float radius, F f) {
    const float lo[3] = {__fsub_rn(x, radius), __fsub_rn(y, radius), __fsub_rn(z, radius)};
    const float hi[3] = {__fadd_rn(x, radius), __fadd_rn(y, radius), __fadd_rn(z, radius)};
    int n = 0;
    if(!(lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2])) return 0;
    int c0[3], c1[3];
    #pragma unroll
    for(int d = 0; d < 3; d++) { c0[d] = cell_coord(ix.geom, d, lo[d]); c1[d] = cell_coord(ix.geom, d, hi[d]); }
    for(int cz = c0[2]; cz <= c1[2]; cz++)
        for(int cy = c0[1]; cy <= c1[1]; cy++) {
            const int base = (cz * ix.geom.n[1] + cy) * ix.geom.n[0];
            for(int s = ix.cell_start[base + c0[0]]; s < ix.cell_start[base + c1[0] + 1]; s++) {
                const float px = ix.sx[s], py = ix.sy[s], pz = ix.sz[s];
                if(!(px > lo[0] && px < hi[0] && py > lo[1] && py < hi[1] && pz > lo[2] && pz < hi[2])) continue;
                if(straight_distance(px, py, pz, x, y, z) <= radius) {
                    f(ix.order[s]);
                    n++;
                }
            }
        }
    return n;
}

__global__ void count_neighbours_kernel(IndexView ix, const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
                                        long long q0, int nq, float radius, long long* __restrict__ counts, float* __restrict__ out_float) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= nq) return;
    const long long q = q0 + t;
    const int n = for_each_neighbour(ix, qx[q], qy[q], qz[q], radius, [](int) {});
    if(counts) counts[t] = n;
    if(out_float) out_float[q] = (float) n;
}
// offsets[t] = first slot of query t's neighbour list; keys = (t << 32) | point index
__global__ void list_neighbours_kernel(IndexView ix, const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
                                       long long q0, int nq, float radius, const long long* __restrict__ offsets, unsigned long long* __restrict__ keys) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= nq) return;
    const long long q = q0 + t;
    unsigned long long* mine = keys + offsets[t];
    int k = 0;
    for_each_neighbour(ix, qx[q], qy[q], qz[q], radius, [&](int i) { mine[k++] = ((unsigned long long) t << 32) | (unsigned) i; });
}
// gridding.cpp:24-31 / :52-59: fewer than min_num neighbours -> missing; gridding_nearest (:99-103): also no neighbour at all
__global__ void apply_min_num_kernel(const long long* __restrict__ offsets, int nq, int min_num, bool empty_is_missing, float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= nq) return;
    const long long n = offsets[t + 1] - offsets[t];
    if((min_num > 0 && n < min_num) || (empty_is_missing && n == 0)) out[t] = NAN;
}

// KDTree::calc_distance, kdtree.cpp:107-133: great-circle distance from the spherical law of cosines in double (Geodetic), plane
// distance in float (Cartesian)
__device__ float calc_distance_dev(float lat1, float lon1, float lat2, float lon2, int type) {
    if(type == GPP_CARTESIAN) {
        const float dx = __fsub_rn(lon1, lon2), dy = __fsub_rn(lat1, lat2);
        return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
    if(!(is_valid(lat1) && is_valid(lat2) && is_valid(lon1) && is_valid(lon2))) return NAN;
    if(lat1 == lat2 && lon1 == lon2) return 0.f;
    // deg2rad returns a float (kdtree.cpp:195-197), promoted to double at the assignment
    const double lat1r = (double) (float) ((double) lat1 * M_PI / 180), lat2r = (double) (float) ((double) lat2 * M_PI / 180);
    const double lon1r = (double) (float) ((double) lon1 * M_PI / 180), lon2r = (double) (float) ((double) lon2 * M_PI / 180);
    const double ratio = cos(lat1r) * cos(lon1r) * cos(lat2r) * cos(lon2r) + cos(lat1r) * sin(lon1r) * cos(lat2r) * sin(lon2r) + sin(lat1r) * sin(lat2r);
    return (float) (acos(ratio) * 6.378137e6);
}
// distance.cpp: the largest calc_distance to the `num` closest points
__global__ void max_distance_kernel(const int* __restrict__ nn, int num, const float* __restrict__ qlat, const float* __restrict__ qlon, int nq,
                                    const float* __restrict__ plat, const float* __restrict__ plon, int type, bool query_first, float* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nq) return;
    float worst = 0.f;
    for(int k = 0; k < num; k++) {
        const int i = nn[(size_t) q * num + k];
        if(i < 0) continue;
        // the argument order differs between the overloads (distance.cpp:21 vs :52); it matters only for rounding
        const float d = query_first ? calc_distance_dev(qlat[q], qlon[q], plat[i], plon[i], type) : calc_distance_dev(plat[i], plon[i], qlat[q], qlon[q], type);
        if(d > worst) worst = d;
    }
    out[q] = worst;
}

// fill (fill.cpp:6-43): a cell inside any circle takes `value` (inside) or keeps its input (outside mode, the rest takes `value`).
// doping_circle (doping.cpp:52-93): the LAST point (highest index) whose circle holds the cell and passes the elevation test wins.
// One thread per point; winners are resolved with atomicMax on the point index.
__global__ void circle_winner_kernel(IndexView grid_ix, const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
                                     const float* __restrict__ radii, int np, const float* __restrict__ pelev, const float* __restrict__ gelev,
                                     float max_elev_diff, int* __restrict__ winner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= np) return;
    const bool check_elev = is_valid(max_elev_diff) && pelev != nullptr;
    const float e = check_elev ? pelev[i] : 0.f;
    for_each_neighbour(grid_ix, px[i], py[i], pz[i], radii[i], [&](int cell) {
        if(check_elev && fabsf(__fsub_rn(e, gelev[cell])) > max_elev_diff) return;   // doping.cpp:82-86 (a NaN difference passes, as there)
        atomicMax(&winner[cell], i);
    });
}
// doping_square (doping.cpp:5-51): the square of half-width hw[i] cells around the grid node nearest to point i
__global__ void square_winner_kernel(const int* __restrict__ nearest, const int* __restrict__ halfwidth, int np, int ny, int nx,
                                     const float* __restrict__ pelev, const float* __restrict__ gelev, float max_elev_diff, int* __restrict__ winner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= np) return;
    const int node = nearest[i];
    if(node < 0) return;
    const int cy = node / nx, cx = node - cy * nx, hw = halfwidth[i];
    const bool check_elev = is_valid(max_elev_diff);
    for(int yy = max(0, cy - hw); yy <= min(ny - 1, cy + hw); yy++)
        for(int xx = max(0, cx - hw); xx <= min(nx - 1, cx + hw); xx++) {
            if(check_elev && fabsf(__fsub_rn(pelev[i], gelev[(size_t) yy * nx + xx])) > max_elev_diff) continue;
            atomicMax(&winner[(size_t) yy * nx + xx], i);
        }
}
// mode 0: doping (winner's observation, else background); 1: fill inside (value where covered); 2: fill outside (input where covered, else value)
__global__ void apply_winner_kernel(const int* __restrict__ winner, size_t n, int mode, const float* __restrict__ input, const float* __restrict__ obs,
                                    float value, float* __restrict__ out) {
    const size_t c = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= n) return;
    const int w = winner[c];
    if(mode == 0) out[c] = w >= 0 ? obs[w] : input[c];
    else if(mode == 1) out[c] = w >= 0 ? value : input[c];
    else out[c] = w >= 0 ? input[c] : value;
}

// fill_missing (fill.cpp:44-134): linear interpolation between the valid values on either side along x, the same along y,
// and the mean of the two where both exist. One thread per row (pass 0) / per column (pass 1), as serial as the reference.
__global__ void fill_missing_line_kernel(const float* __restrict__ values, int ny, int nx, int along_x, float* __restrict__ result) {
    const int line = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_lines = along_x ? ny : nx, len = along_x ? nx : ny;
    if(line >= n_lines) return;
    const size_t stride = along_x ? 1 : (size_t) nx, base = along_x ? (size_t) line * nx : (size_t) line;
    int last = 0, next = -1;
    for(int i = 0; i < len; i++) {
        const float curr = values[base + i * stride];
        float r = NAN;
        if(!is_valid(curr)) {
            if(next < i)
                for(next = i; next < len; next++)
                    if(is_valid(values[base + next * stride])) break;
            if(next < len) {
                const float value_last = values[base + last * stride], value_next = values[base + next * stride];
                // (value_last) + (value_next - value_last) * (x - last) / (next - last): float * int -> float, / int -> float
                r = __fadd_rn(value_last, __fdiv_rn(__fmul_rn(__fsub_rn(value_next, value_last), (float) (i - last)), (float) (next - last)));
            }
        }
        else {
            last = i;
            r = curr;
        }
        result[base + i * stride] = r;
    }
}
__global__ void fill_missing_combine_kernel(const float* __restrict__ rx, const float* __restrict__ ry, size_t n, float* __restrict__ out) {
    const size_t c = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= n) return;
    int count = 0;
    float total = 0.f;
    // fill.cpp:117-130: results_y (the pass along x) first
    if(is_valid(rx[c])) { total = __fadd_rn(total, rx[c]); count++; }
    if(is_valid(ry[c])) { total = __fadd_rn(total, ry[c]); count++; }
    out[c] = count > 0 ? __fdiv_rn(total, (float) count) : NAN;
}

unsigned blocks_for(size_t n, int block = 128) { return (unsigned) ((n + block - 1) / block); }

// the statistic of the values of the neighbours of every output location, in chunks that bound the pair list
int gridding_radius(gpp_points* op, gpp_points* ip, const float* d_values, float radius, int min_num, int statistic, float* d_out) {
    const IndexView ix = view_of(ip);
    const long long nq = op->n;
    const long long max_pairs = 48LL << 20;   // 384 MB of keys per chunk
    DeviceBuffer<long long> counts;
    DeviceBuffer<unsigned long long> keys;
    long long q0 = 0;
    int chunk = (int) std::min<long long>(nq, 1 << 20);
    cudaStream_t stream = 0;
    auto policy = thrust::cuda::par.on(stream);
    while(q0 < nq) {
        int n = (int) std::min<long long>(chunk, nq - q0);
        GPP_TRY(counts.alloc((size_t) n + 1));
        GPP_LAUNCH(count_neighbours_kernel, blocks_for(n), 128, 0, stream, ix, op->dx.ptr, op->dy.ptr, op->dz.ptr, q0, n, radius, counts.ptr, nullptr);
        GPP_CUDA(cudaMemsetAsync(counts.ptr + n, 0, sizeof(long long), stream));
        thrust::exclusive_scan(policy, counts.ptr, counts.ptr + n + 1, counts.ptr);
        long long total = 0;
        GPP_CUDA(cudaMemcpyAsync(&total, counts.ptr + n, sizeof(long long), cudaMemcpyDeviceToHost, stream));
        GPP_CUDA(cudaStreamSynchronize(stream));
        if(total > max_pairs && n > 1) {   // too many pairs at once: halve the chunk and retry
            chunk = std::max(1, n / 2);
            continue;
        }
        GPP_TRY(keys.alloc((size_t) std::max<long long>(total, 1)));
        if(total > 0) {
            GPP_LAUNCH(list_neighbours_kernel, blocks_for(n), 128, 0, stream, ix, op->dx.ptr, op->dy.ptr, op->dz.ptr, q0, n, radius, counts.ptr, keys.ptr);
            thrust::sort(policy, keys.ptr, keys.ptr + total);   // by query, then ascending point index
            g_launches.fetch_add(2, std::memory_order_relaxed);
        }
        const SegmentRows R = {d_values, keys.ptr, counts.ptr};
        GPP_TRY(run_rows(R, n, statistic, NAN, nullptr, d_out + q0, stream, nullptr));
        GPP_LAUNCH(apply_min_num_kernel, blocks_for(n), 128, 0, stream, counts.ptr, n, min_num, false, d_out + q0);
        GPP_CUDA(cudaStreamSynchronize(stream));
        q0 += n;
    }
    return GPP_OK;
}

bool statistic_ok(int s) {
    return s == GPP_MEAN || s == GPP_MIN || s == GPP_MEDIAN || s == GPP_MAX || s == GPP_STD || s == GPP_VARIANCE || s == GPP_SUM || s == GPP_COUNT ||
           s == GPP_RANDOMCHOICE;
}

}  // namespace

extern "C" {

int gpp_gridding_host(const gpp_points* opoints, const gpp_points* ipoints, const float* values, float radius, int min_num, int statistic,
                      float* output) {
    if(!opoints || !ipoints) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(!is_valid(radius) || radius < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "radius must be >= 0");     // gridding.cpp:9-10
    if(min_num < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "min_num must be >= 0");                         // :11-12
    if(!statistic_ok(statistic)) return fail(GPP_ERR_RUNTIME, "Internal error. Cannot compute statistic");
    GPP_TRY(ensure_device());
    gpp_points *op = const_cast<gpp_points*>(opoints), *ip = const_cast<gpp_points*>(ipoints);
    if(op->n == 0) return GPP_OK;
    GPP_TRY(op->ensure_on_device());
    GPP_TRY(ip->ensure_index());
    DeviceBuffer<float> d_values, d_out;
    GPP_TRY(d_values.upload(values, (size_t) std::max(ip->n, 1)));
    GPP_TRY(d_out.alloc((size_t) op->n));
    GPP_TRY(gridding_radius(op, ip, d_values.ptr, radius, min_num, statistic, d_out.ptr));
    GPP_TRY(d_out.download(output, (size_t) op->n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_gridding_nearest_host(const gpp_points* opoints, const gpp_points* ipoints, const float* values, int min_num, int statistic,
                              float* output) {
    if(!opoints || !ipoints) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(min_num < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "min_num must be >= 0");                         // gridding.cpp:69-70
    if(!statistic_ok(statistic)) return fail(GPP_ERR_RUNTIME, "Internal error. Cannot compute statistic");
    GPP_TRY(ensure_device());
    gpp_points *op = const_cast<gpp_points*>(opoints), *ip = const_cast<gpp_points*>(ipoints);
    const int N = op->n, S = ip->n;
    if(N == 0) return GPP_OK;
    if(S == 0) {
        for(int i = 0; i < N; i++) output[i] = NAN;
        return GPP_OK;
    }
    // the output node nearest to every input point (gridding.cpp:85-90,118-121), then the input points grouped by node in
    // ascending order (the reference pushes them back in index order)
    std::vector<int> nn((size_t) S);
    GPP_TRY(gpp_points_nearest_host(op, ip->lats.data(), ip->lons.data(), S, 1, nn.data()));
    std::vector<unsigned long long> keys((size_t) S);
    for(int s = 0; s < S; s++) keys[s] = ((unsigned long long) (unsigned) nn[s] << 32) | (unsigned) s;
    std::sort(keys.begin(), keys.end());
    std::vector<long long> offsets((size_t) N + 1, 0);
    for(int s = 0; s < S; s++) offsets[(size_t) (keys[s] >> 32) + 1]++;
    for(int n = 0; n < N; n++) offsets[n + 1] += offsets[n];
    DeviceBuffer<float> d_values, d_out;
    DeviceBuffer<unsigned long long> d_keys;
    DeviceBuffer<long long> d_off;
    GPP_TRY(d_values.upload(values, (size_t) S));
    GPP_TRY(d_keys.upload(keys.data(), (size_t) S));
    GPP_TRY(d_off.upload(offsets.data(), (size_t) N + 1));
    GPP_TRY(d_out.alloc((size_t) N));
    const SegmentRows R = {d_values.ptr, d_keys.ptr, d_off.ptr};
    GPP_TRY(run_rows(R, N, statistic, NAN, nullptr, d_out.ptr, 0, nullptr));
    GPP_LAUNCH(apply_min_num_kernel, blocks_for(N), 128, 0, 0, d_off.ptr, N, min_num, true, d_out.ptr);
    GPP_TRY(d_out.download(output, (size_t) N));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_count_host(const gpp_points* ipoints, const gpp_points* opoints, float radius, float* output) {
    if(!opoints || !ipoints) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    GPP_TRY(ensure_device());
    gpp_points *op = const_cast<gpp_points*>(opoints), *ip = const_cast<gpp_points*>(ipoints);
    if(op->n == 0) return GPP_OK;
    GPP_TRY(op->ensure_on_device());
    GPP_TRY(ip->ensure_index());
    DeviceBuffer<float> d_out;
    GPP_TRY(d_out.alloc((size_t) op->n));
    GPP_LAUNCH(count_neighbours_kernel, blocks_for(op->n), 128, 0, 0, view_of(ip), op->dx.ptr, op->dy.ptr, op->dz.ptr, 0LL, op->n, radius, nullptr,
               d_out.ptr);
    GPP_TRY(d_out.download(output, (size_t) op->n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_distance_host(const gpp_points* ipoints, const gpp_points* opoints, int num, int query_first, float* output) {
    if(!opoints || !ipoints) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(ipoints->type != opoints->type) return fail(GPP_ERR_INVALID_ARGUMENT, "Incompatible coordinate types");   // distance.cpp:7-8
    if(num < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "num must be >= 0");
    GPP_TRY(ensure_device());
    const int nq = opoints->n;
    if(nq == 0) return GPP_OK;
    if(num == 0 || ipoints->n == 0) {
        for(int i = 0; i < nq; i++) output[i] = 0.f;
        return GPP_OK;
    }
    std::vector<int> nn((size_t) nq * num);
    GPP_TRY(gpp_points_closest_host(ipoints, opoints->lats.data(), opoints->lons.data(), nq, num, 1, nn.data()));
    DeviceBuffer<int> d_nn;
    DeviceBuffer<float> d_qlat, d_qlon, d_plat, d_plon, d_out;
    GPP_TRY(d_nn.upload(nn.data(), nn.size()));
    GPP_TRY(d_qlat.upload(opoints->lats.data(), (size_t) nq));
    GPP_TRY(d_qlon.upload(opoints->lons.data(), (size_t) nq));
    GPP_TRY(d_plat.upload(ipoints->lats.data(), (size_t) ipoints->n));
    GPP_TRY(d_plon.upload(ipoints->lons.data(), (size_t) ipoints->n));
    GPP_TRY(d_out.alloc((size_t) nq));
    GPP_LAUNCH(max_distance_kernel, blocks_for(nq), 128, 0, 0, d_nn.ptr, num, d_qlat.ptr, d_qlon.ptr, nq, d_plat.ptr, d_plon.ptr, ipoints->type,
               query_first != 0, d_out.ptr);
    GPP_TRY(d_out.download(output, (size_t) nq));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

namespace {
// shared by fill and doping_circle: winners of the circles of `points` over the nodes of `grid`
int circle_winners(gpp_points* grid, gpp_points* pts, const float* radii, bool with_elev, float max_elev_diff, DeviceBuffer<int>& winner) {
    GPP_TRY(grid->ensure_index());
    GPP_TRY(pts->ensure_on_device());
    GPP_TRY(winner.alloc((size_t) grid->n));
    GPP_CUDA(cudaMemsetAsync(winner.ptr, 0xff, sizeof(int) * (size_t) grid->n, 0));   // -1
    if(pts->n == 0) return GPP_OK;
    DeviceBuffer<float> d_radii;
    GPP_TRY(d_radii.upload(radii, (size_t) pts->n));
    GPP_LAUNCH(circle_winner_kernel, blocks_for(pts->n, 64), 64, 0, 0, view_of(grid), pts->dx.ptr, pts->dy.ptr, pts->dz.ptr, d_radii.ptr, pts->n,
               with_elev ? pts->delev.ptr : nullptr, grid->delev.ptr, with_elev ? max_elev_diff : NAN, winner.ptr);
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}
int apply_and_download(const DeviceBuffer<int>& winner, size_t n, int mode, const float* input, const float* obs, int n_obs, float value, float* output) {
    DeviceBuffer<float> d_in, d_obs, d_out;
    GPP_TRY(d_in.upload(input, n));
    if(obs) GPP_TRY(d_obs.upload(obs, (size_t) std::max(n_obs, 1)));
    GPP_TRY(d_out.alloc(n));
    GPP_LAUNCH(apply_winner_kernel, blocks_for(n, 256), 256, 0, 0, winner.ptr, n, mode, d_in.ptr, obs ? d_obs.ptr : nullptr, value, d_out.ptr);
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}
}  // namespace

int gpp_fill_host(const gpp_points* igrid, const float* input, const gpp_points* points, const float* radii, float value, int outside,
                  float* output) {
    if(!igrid || !points) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    for(int i = 0; i < points->n; i++)
        if(radii[i] < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "All radius sizes must be 0 or greater");   // fill.cpp:11-14
    GPP_TRY(ensure_device());
    if(igrid->n == 0) return GPP_OK;
    DeviceBuffer<int> winner;
    GPP_TRY(circle_winners(const_cast<gpp_points*>(igrid), const_cast<gpp_points*>(points), radii, false, NAN, winner));
    return apply_and_download(winner, (size_t) igrid->n, outside ? 2 : 1, input, nullptr, 0, value, output);
}

int gpp_doping_circle_host(const gpp_points* igrid, const float* background, const gpp_points* points, const float* observations,
                           const float* radii, float max_elev_diff, float* output) {
    if(!igrid || !points) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(is_valid(max_elev_diff) && max_elev_diff < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_elev_diff must be greater than or equal to 0");
    for(int i = 0; i < points->n; i++)
        if(radii[i] < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "radii must be greater than or equal to 0");   // doping.cpp:70-73
    GPP_TRY(ensure_device());
    if(igrid->n == 0) return GPP_OK;
    DeviceBuffer<int> winner;
    GPP_TRY(const_cast<gpp_points*>(igrid)->ensure_on_device());
    GPP_TRY(circle_winners(const_cast<gpp_points*>(igrid), const_cast<gpp_points*>(points), radii, true, max_elev_diff, winner));
    return apply_and_download(winner, (size_t) igrid->n, 0, background, observations, points->n, 0.f, output);
}

int gpp_doping_square_host(const gpp_points* igrid, const float* background, const gpp_points* points, const float* observations,
                           const int* halfwidth, float max_elev_diff, float* output) {
    if(!igrid || !points) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(is_valid(max_elev_diff) && max_elev_diff < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_elev_diff must be greater than or equal to 0");
    for(int i = 0; i < points->n; i++)
        if(halfwidth[i] < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "All halfwidth must be greater than or equal to 0");   // doping.cpp:25-28
    if(igrid->n > 0 && (igrid->shape_nx <= 0 || igrid->shape_ny <= 0)) return fail(GPP_ERR_INVALID_ARGUMENT, "doping_square needs a grid (gpp_points_set_shape)");
    GPP_TRY(ensure_device());
    if(igrid->n == 0) return GPP_OK;
    gpp_points *grid = const_cast<gpp_points*>(igrid), *pts = const_cast<gpp_points*>(points);
    GPP_TRY(grid->ensure_on_device());
    GPP_TRY(pts->ensure_on_device());
    DeviceBuffer<int> winner, d_nn, d_hw;
    GPP_TRY(winner.alloc((size_t) grid->n));
    GPP_CUDA(cudaMemsetAsync(winner.ptr, 0xff, sizeof(int) * (size_t) grid->n, 0));
    if(pts->n > 0) {
        std::vector<int> nn((size_t) pts->n);
        GPP_TRY(gpp_points_nearest_host(grid, pts->lats.data(), pts->lons.data(), pts->n, 1, nn.data()));   // doping.cpp:32
        GPP_TRY(d_nn.upload(nn.data(), nn.size()));
        GPP_TRY(d_hw.upload(halfwidth, (size_t) pts->n));
        GPP_LAUNCH(square_winner_kernel, blocks_for(pts->n, 64), 64, 0, 0, d_nn.ptr, d_hw.ptr, pts->n, grid->shape_ny, grid->shape_nx, pts->delev.ptr,
                   grid->delev.ptr, max_elev_diff, winner.ptr);
    }
    return apply_and_download(winner, (size_t) grid->n, 0, background, observations, pts->n, 0.f, output);
}

int gpp_fill_missing_host(const float* values, int ny, int nx, float* output) {
    if(ny < 0 || nx < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative size");
    GPP_TRY(ensure_device());
    if(ny == 0 || nx == 0) return GPP_OK;
    const size_t n = (size_t) ny * nx;
    DeviceBuffer<float> d_in, d_rx, d_ry, d_out;
    GPP_TRY(d_in.upload(values, n));
    GPP_TRY(d_rx.alloc(n));
    GPP_TRY(d_ry.alloc(n));
    GPP_TRY(d_out.alloc(n));
    GPP_LAUNCH(fill_missing_line_kernel, blocks_for(ny, 64), 64, 0, 0, d_in.ptr, ny, nx, 1, d_rx.ptr);
    GPP_LAUNCH(fill_missing_line_kernel, blocks_for(nx, 64), 64, 0, 0, d_in.ptr, ny, nx, 0, d_ry.ptr);
    GPP_LAUNCH(fill_missing_combine_kernel, blocks_for(n, 256), 256, 0, 0, d_rx.ptr, d_ry.ptr, n, d_out.ptr);
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

}  // extern "C"
