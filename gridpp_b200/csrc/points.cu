// Point sets, their device bucket index, and the nearest / radius / k-nearest query kernels.
// Replaces gridpp::KDTree (src/api/kdtree.cpp), gridpp::Points (points.cpp), the flattened gridpp::Grid
// (grid.cpp) and the per-point loops of gridpp::nearest (nearest.cpp).
#include "points.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

using namespace gpp;

namespace {

// gridpp::convert_coordinates, util.cpp:583-615 (+ is_valid_lat/lon :617-624). Done on the HOST with the same
// libm calls as the reference so that x/y/z agree with it bit for bit.
bool convert_one(float lat, float lon, int type, float& x, float& y, float& z) {
    bool ok_lat = type == GPP_CARTESIAN ? is_valid(lat) : (is_valid(lat) && (lat >= -90.001) && (lat <= 90.001));
    if(!ok_lat || !is_valid(lon)) return false;
    if(type == GPP_CARTESIAN) {
        x = lon;
        y = lat;
        z = 0;
    }
    else {
        const double radius_earth = 6.378137e6;   // gridpp.h:56
        double lonr = M_PI / 180 * lon;
        double latr = M_PI / 180 * lat;
        x = std::cos(latr) * std::cos(lonr) * radius_earth;
        y = std::cos(latr) * std::sin(lonr) * radius_earth;
        z = std::sin(latr) * radius_earth;
    }
    return true;
}

int convert_queries(const float* qlats, const float* qlons, int nq, int type, std::vector<float>& x, std::vector<float>& y,
                    std::vector<float>& z) {
    x.resize(nq); y.resize(nq); z.resize(nq);
    int bad = -1;
    #pragma omp parallel for
    for(int i = 0; i < nq; i++)
        if(!convert_one(qlats[i], qlons[i], type, x[i], y[i], z[i])) bad = i;
    if(bad >= 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Invalid coords: %g,%g", qlats[bad], qlons[bad]);
    return GPP_OK;
}

// ------------------------------------------------------------------ index build (K4) ------------------
__global__ void cell_count_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int n,
                                  CellGeom g, int* __restrict__ cell_of_point, int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    int c = (cell_coord(g, 2, z[i]) * g.n[1] + cell_coord(g, 1, y[i])) * g.n[0] + cell_coord(g, 0, x[i]);
    cell_of_point[i] = c;
    atomicAdd(&counts[c], 1);
}
// exclusive scan of counts[0..n) into start[0..n], single CTA (n is at most 2^24; one-off per point set)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ start) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if(threadIdx.x == 0) carry = 0;
    __syncthreads();
    for(int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < n ? counts[i] : 0;
        int incl = v;
        #pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if(lane_id() >= (unsigned) d) incl += t;
        }
        if(lane_id() == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if(threadIdx.x < 32) {
            int w = warp_sums[threadIdx.x];
            int wi = w;
            #pragma unroll
            for(int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, wi, d);
                if(lane_id() >= (unsigned) d) wi += t;
            }
            warp_sums[threadIdx.x] = wi - w;   // exclusive prefix of the warp totals
        }
        __syncthreads();
        int excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
        if(i < n) start[i] = excl;
        __syncthreads();
        if(threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if(threadIdx.x == 0) start[n] = carry;
}
__global__ void cell_scatter_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, int n,
                                    const int* __restrict__ cell_of_point, const int* __restrict__ start, int* __restrict__ cursor,
                                    int* __restrict__ order, float* __restrict__ sx, float* __restrict__ sy, float* __restrict__ sz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    int c = cell_of_point[i];
    int slot = start[c] + atomicAdd(&cursor[c], 1);   // order inside a cell is arbitrary; every query is order-independent
    order[slot] = i;
    sx[slot] = x[i];
    sy[slot] = y[i];
    sz[slot] = z[i];
}

// ------------------------------------------------------------------ queries (K5) ----------------------
// Squared distance in double with individually rounded operations: what Boost's rtree compares for float
// points (comparable distance promoted to double); see DESIGN.md, "Index lookups".
__device__ __forceinline__ double dist2_double(float qx, float qy, float qz, float px, float py, float pz) {
    double dx = __dsub_rn((double) qx, (double) px), dy = __dsub_rn((double) qy, (double) py), dz = __dsub_rn((double) qz, (double) pz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// k nearest (k <= KMAX) per query, one thread per query, expanding Chebyshev shells of cells.
// KDTree::get_closest_neighbours kdtree.cpp:82-103 (+ is_not_equal :262-270). Ties -> lowest index.
constexpr int KNN_MAX = 16;
template <int K>
__global__ void __launch_bounds__(128) knn_kernel(const float* __restrict__ qx, const float* __restrict__ qy,
                                                  const float* __restrict__ qz, int nq, CellGeom g,
                                                  const int* __restrict__ cell_start, const int* __restrict__ order,
                                                  const float* __restrict__ sx, const float* __restrict__ sy,
                                                  const float* __restrict__ sz, int k, int include_match,
                                                  int* __restrict__ out) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nq) return;
    float x = qx[q], y = qy[q], z = qz[q];
    double bd[K];
    int bi[K];
    #pragma unroll
    for(int i = 0; i < K; i++) { bd[i] = INFINITY; bi[i] = -1; }
    int cc[3] = {cell_coord(g, 0, x), cell_coord(g, 1, y), cell_coord(g, 2, z)};
    int maxr = max(g.n[0], max(g.n[1], g.n[2]));
    float max_edge = fmaxf(g.edge[0], fmaxf(g.edge[1], g.edge[2]));
    for(int r = 0; r <= maxr; r++) {
        int b0[3], b1[3];
        #pragma unroll
        for(int d = 0; d < 3; d++) { b0[d] = max(0, cc[d] - r); b1[d] = min(g.n[d] - 1, cc[d] + r); }
        for(int cz = b0[2]; cz <= b1[2]; cz++)
            for(int cy = b0[1]; cy <= b1[1]; cy++) {
                bool row_on_shell = abs(cz - cc[2]) == r || abs(cy - cc[1]) == r;
                int nx = row_on_shell ? (b1[0] - b0[0] + 1) : (r == 0 ? 1 : 2);
                for(int ix = 0; ix < nx; ix++) {
                    int cx = row_on_shell ? b0[0] + ix : (ix == 0 ? cc[0] - r : cc[0] + r);
                    if(cx < 0 || cx > g.n[0] - 1) continue;
                    int id = (cz * g.n[1] + cy) * g.n[0] + cx;
                    for(int s = cell_start[id]; s < cell_start[id + 1]; s++) {
                        float px = sx[s], py = sy[s], pz = sz[s];
                        if(!include_match && px == x && py == y && pz == z) continue;
                        double d2 = dist2_double(x, y, z, px, py, pz);
                        int idx = order[s];
                        // insert into the sorted (d2, idx) list; K is tiny
                        if(d2 < bd[K - 1] || (d2 == bd[K - 1] && idx < bi[K - 1]) || bi[K - 1] < 0) {
                            double cd = d2;
                            int ci = idx;
                            #pragma unroll
                            for(int j = 0; j < K; j++) {
                                bool better = bi[j] < 0 || cd < bd[j] || (cd == bd[j] && ci < bi[j]);
                                if(better) {
                                    double td = bd[j]; int ti = bi[j];
                                    bd[j] = cd; bi[j] = ci;
                                    cd = td; ci = ti;
                                    if(ci < 0) break;
                                }
                            }
                        }
                    }
                }
            }
        bool all = true;
        #pragma unroll
        for(int d = 0; d < 3; d++) if(b0[d] > 0 || b1[d] < g.n[d] - 1) all = false;
        if(all) break;
        int kk = min(k, K);
        if(bi[kk - 1] >= 0) {
            // smallest possible distance to anything outside the visited block of cells
            float bound = INFINITY;
            float qc[3] = {x, y, z};
            #pragma unroll
            for(int d = 0; d < 3; d++) {
                if(b0[d] > 0) bound = fminf(bound, qc[d] - (g.lo[d] + b0[d] * g.edge[d]));
                if(b1[d] < g.n[d] - 1) bound = fminf(bound, (g.lo[d] + (b1[d] + 1) * g.edge[d]) - qc[d]);
            }
            bound -= 1e-3f * max_edge;   // slack for the rounding in the cell assignment
            if(bound > 0.f && bd[kk - 1] < (double) bound * (double) bound) break;
        }
    }
    for(int i = 0; i < k; i++) out[(size_t) q * k + i] = i < K ? bi[i] : -1;
}

// The same search for any k: the sorted (d2, index) list of a query lives in global memory (bd scratch, the output row itself
// holds the indices) instead of registers. Used above KNN_MAX neighbours, where insertion cost no longer matters next to the
// number of cells a query has to visit.
__global__ void __launch_bounds__(128) knn_any_kernel(const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz,
                                                      int nq, CellGeom g, const int* __restrict__ cell_start, const int* __restrict__ order,
                                                      const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sz,
                                                      int k, int include_match, double* __restrict__ scratch, int* __restrict__ out) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nq) return;
    float x = qx[q], y = qy[q], z = qz[q];
    double* bd = scratch + (size_t) q * k;
    int* bi = out + (size_t) q * k;
    for(int i = 0; i < k; i++) { bd[i] = INFINITY; bi[i] = -1; }
    int have = 0;
    int cc[3] = {cell_coord(g, 0, x), cell_coord(g, 1, y), cell_coord(g, 2, z)};
    int maxr = max(g.n[0], max(g.n[1], g.n[2]));
    float max_edge = fmaxf(g.edge[0], fmaxf(g.edge[1], g.edge[2]));
    for(int r = 0; r <= maxr; r++) {
        int b0[3], b1[3];
        #pragma unroll
        for(int d = 0; d < 3; d++) { b0[d] = max(0, cc[d] - r); b1[d] = min(g.n[d] - 1, cc[d] + r); }
        for(int cz = b0[2]; cz <= b1[2]; cz++)
            for(int cy = b0[1]; cy <= b1[1]; cy++) {
                bool row_on_shell = abs(cz - cc[2]) == r || abs(cy - cc[1]) == r;
                int nx = row_on_shell ? (b1[0] - b0[0] + 1) : (r == 0 ? 1 : 2);
                for(int ix = 0; ix < nx; ix++) {
                    int cx = row_on_shell ? b0[0] + ix : (ix == 0 ? cc[0] - r : cc[0] + r);
                    if(cx < 0 || cx > g.n[0] - 1) continue;
                    int id = (cz * g.n[1] + cy) * g.n[0] + cx;
                    for(int s = cell_start[id]; s < cell_start[id + 1]; s++) {
                        float px = sx[s], py = sy[s], pz = sz[s];
                        if(!include_match && px == x && py == y && pz == z) continue;
                        double d2 = dist2_double(x, y, z, px, py, pz);
                        int idx = order[s];
                        if(have == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && idx < bi[k - 1]))) continue;
                        int j = have < k ? have : k - 1;     // shift the worse entries down, insert
                        while(j > 0 && (d2 < bd[j - 1] || (d2 == bd[j - 1] && idx < bi[j - 1]))) {
                            bd[j] = bd[j - 1];
                            bi[j] = bi[j - 1];
                            j--;
                        }
                        bd[j] = d2;
                        bi[j] = idx;
                        if(have < k) have++;
                    }
                }
            }
        bool all = true;
        #pragma unroll
        for(int d = 0; d < 3; d++) if(b0[d] > 0 || b1[d] < g.n[d] - 1) all = false;
        if(all) break;
        if(have == k) {
            float bound = INFINITY;
            float qc[3] = {x, y, z};
            #pragma unroll
            for(int d = 0; d < 3; d++) {
                if(b0[d] > 0) bound = fminf(bound, qc[d] - (g.lo[d] + b0[d] * g.edge[d]));
                if(b1[d] < g.n[d] - 1) bound = fminf(bound, (g.lo[d] + (b1[d] + 1) * g.edge[d]) - qc[d]);
            }
            bound -= 1e-3f * max_edge;
            if(bound > 0.f && bd[k - 1] < (double) bound * (double) bound) break;
        }
    }
}

// Radius query, one thread per query. KDTree::get_neighbours kdtree.cpp:39-62 with within_radius :247-260:
// STRICTLY inside the box [q-r, q+r]^3 (Boost within()), straight distance <= r, > 0 unless include_match.
// Pass 1 (out_index == NULL) counts; pass 2 stores up to `capacity` indices and sorts them ascending.
__global__ void __launch_bounds__(128) radius_kernel(const float* __restrict__ qx, const float* __restrict__ qy,
                                                     const float* __restrict__ qz, const float* __restrict__ radii, int nq,
                                                     CellGeom g, const int* __restrict__ cell_start,
                                                     const int* __restrict__ order, const float* __restrict__ sx,
                                                     const float* __restrict__ sy, const float* __restrict__ sz,
                                                     int include_match, int capacity, int* __restrict__ out_index,
                                                     float* __restrict__ out_dist, int* __restrict__ out_count) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nq) return;
    float x = qx[q], y = qy[q], z = qz[q], radius = radii[q];
    float lo[3] = {__fsub_rn(x, radius), __fsub_rn(y, radius), __fsub_rn(z, radius)};
    float hi[3] = {__fadd_rn(x, radius), __fadd_rn(y, radius), __fadd_rn(z, radius)};
    int n = 0;
    if(lo[0] < hi[0] && lo[1] < hi[1] && lo[2] < hi[2]) {
        int c0[3], c1[3];
        #pragma unroll
        for(int d = 0; d < 3; d++) { c0[d] = cell_coord(g, d, lo[d]); c1[d] = cell_coord(g, d, hi[d]); }
        int* mine = out_index ? out_index + (size_t) q * capacity : nullptr;
        for(int cz = c0[2]; cz <= c1[2]; cz++)
            for(int cy = c0[1]; cy <= c1[1]; cy++) {
                int base = (cz * g.n[1] + cy) * g.n[0];
                for(int s = cell_start[base + c0[0]]; s < cell_start[base + c1[0] + 1]; s++) {
                    float px = sx[s], py = sy[s], pz = sz[s];
                    if(!(px > lo[0] && px < hi[0] && py > lo[1] && py < hi[1] && pz > lo[2] && pz < hi[2])) continue;
                    float dist = straight_distance(px, py, pz, x, y, z);
                    bool ok = include_match ? (dist <= radius) : (dist <= radius && dist > 0.f);
                    if(!ok) continue;
                    if(mine) {
                        // keep the `capacity` smallest indices, sorted ascending (insertion sort)
                        int idx = order[s];
                        int m = min(n, capacity);
                        if(m < capacity || idx < mine[m - 1]) {
                            int pos = m < capacity ? m : m - 1;
                            while(pos > 0 && mine[pos - 1] > idx) { mine[pos] = mine[pos - 1]; pos--; }
                            mine[pos] = idx;
                        }
                    }
                    n++;
                }
            }
    }
    out_count[q] = n;
    (void) out_dist;
}
// distances for the stored neighbours (kdtree.cpp:23-34), from the original-order coordinates
__global__ void radius_dist_kernel(const float* __restrict__ qx, const float* __restrict__ qy, const float* __restrict__ qz, int nq,
                                   const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
                                   int capacity, const int* __restrict__ index, const int* __restrict__ count,
                                   float* __restrict__ out_dist) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= (size_t) nq * capacity) return;
    int q = (int) (i / capacity), j = (int) (i % capacity);
    if(j >= count[q]) return;
    int p = index[i];
    out_dist[i] = straight_distance(qx[q], qy[q], qz[q], px[p], py[p], pz[p]);
}

// gridpp::nearest gather (nearest.cpp): out[f][q] = ivalues[f][index[q]]
__global__ void gather_kernel(const float* __restrict__ ivalues, int n_in, const int* __restrict__ index, int nq, int n_fields,
                              float* __restrict__ out) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= (size_t) nq * n_fields) return;
    int f = (int) (i / nq), q = (int) (i % nq);
    int p = index[q];
    out[i] = p >= 0 ? ivalues[(size_t) f * n_in + p] : NAN;
}

inline unsigned blocks_for(size_t n, int block) { return (unsigned) ((n + block - 1) / block); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------
int gpp_points::ensure_on_device() {
    std::lock_guard<std::mutex> lock(mutex);
    if(on_device) return GPP_OK;
    GPP_TRY(ensure_device());
    GPP_CUDA(cudaGetDevice(&device));
    GPP_TRY(dx.upload(x.data(), n));
    GPP_TRY(dy.upload(y.data(), n));
    GPP_TRY(dz.upload(z.data(), n));
    GPP_TRY(delev.upload(elevs.data(), n));
    GPP_TRY(dlaf.upload(lafs.data(), n));
    GPP_CUDA(cudaStreamSynchronize(0));
    on_device = true;
    return GPP_OK;
}

int gpp_points::ensure_index() {
    GPP_TRY(ensure_on_device());
    std::lock_guard<std::mutex> lock(mutex);
    if(index.built) return GPP_OK;
    CellGeom& g = index.geom;
    // about 3 points per cell over the non-degenerate dimensions
    double ext[3], vol = 1;
    int nd = 0;
    for(int d = 0; d < 3; d++) {
        ext[d] = n > 0 ? (double) hi[d] - (double) lo[d] : 0;
        if(ext[d] > 0) { nd++; vol *= ext[d]; }
    }
    double edge = (n > 16 && nd > 0) ? std::pow(vol / (n / 3.0), 1.0 / nd) : 0;
    long long total = 1;
    for(int d = 0; d < 3; d++) {
        int c = 1;
        if(edge > 0 && ext[d] > 0) c = (int) std::min(4096.0, std::max(1.0, std::ceil(ext[d] / edge)));
        g.n[d] = c;
        total *= c;
    }
    while(total > (1LL << 24)) {
        int dmax = 0;
        for(int d = 1; d < 3; d++) if(g.n[d] > g.n[dmax]) dmax = d;
        total /= g.n[dmax];
        g.n[dmax] = (g.n[dmax] + 1) / 2;
        total *= g.n[dmax];
    }
    for(int d = 0; d < 3; d++) {
        g.lo[d] = lo[d];
        g.edge[d] = ext[d] > 0 ? (float) (ext[d] / g.n[d]) : 1.f;
        g.inv[d] = ext[d] > 0 ? (float) (g.n[d] / ext[d]) : 0.f;
    }
    index.ncells = (int) total;
    GPP_TRY(index.cell_start.alloc(total + 1));
    GPP_TRY(index.order.alloc(n));
    GPP_TRY(index.sx.alloc(n));
    GPP_TRY(index.sy.alloc(n));
    GPP_TRY(index.sz.alloc(n));
    DeviceBuffer<int> cell_of_point, counts;
    GPP_TRY(cell_of_point.alloc(n));
    GPP_TRY(counts.alloc(total));
    GPP_CUDA(cudaMemsetAsync(counts.ptr, 0, sizeof(int) * total, 0));
    if(n > 0) GPP_LAUNCH(cell_count_kernel, blocks_for(n, 256), 256, 0, 0, dx.ptr, dy.ptr, dz.ptr, n, g, cell_of_point.ptr, counts.ptr);
    GPP_LAUNCH(exclusive_scan_kernel, 1, 1024, 0, 0, counts.ptr, (int) total, index.cell_start.ptr);
    GPP_CUDA(cudaMemsetAsync(counts.ptr, 0, sizeof(int) * total, 0));
    if(n > 0)
        GPP_LAUNCH(cell_scatter_kernel, blocks_for(n, 256), 256, 0, 0, dx.ptr, dy.ptr, dz.ptr, n, cell_of_point.ptr,
                   index.cell_start.ptr, counts.ptr, index.order.ptr, index.sx.ptr, index.sy.ptr, index.sz.ptr);
    GPP_CUDA(cudaStreamSynchronize(0));
    index.built = true;
    return GPP_OK;
}

namespace {
// shared driver for the k-nearest queries; d_out (nq*k ints) stays on the device
int run_knn(gpp_points* p, const float* qlats, const float* qlons, int nq, int k, int include_match, DeviceBuffer<int>& d_out) {
    GPP_TRY(p->ensure_index());
    std::vector<float> qx, qy, qz;
    GPP_TRY(convert_queries(qlats, qlons, nq, p->type, qx, qy, qz));
    DeviceBuffer<float> dqx, dqy, dqz;
    GPP_TRY(dqx.upload(qx.data(), nq));
    GPP_TRY(dqy.upload(qy.data(), nq));
    GPP_TRY(dqz.upload(qz.data(), nq));
    GPP_TRY(d_out.alloc((size_t) nq * k));
    const CellIndex& ix = p->index;
    if(nq > 0) {
        if(k == 1)
            GPP_LAUNCH(knn_kernel<1>, blocks_for(nq, 128), 128, 0, 0, dqx.ptr, dqy.ptr, dqz.ptr, nq, ix.geom, ix.cell_start.ptr,
                       ix.order.ptr, ix.sx.ptr, ix.sy.ptr, ix.sz.ptr, k, include_match, d_out.ptr);
        else if(k <= KNN_MAX)
            GPP_LAUNCH(knn_kernel<KNN_MAX>, blocks_for(nq, 128), 128, 0, 0, dqx.ptr, dqy.ptr, dqz.ptr, nq, ix.geom,
                       ix.cell_start.ptr, ix.order.ptr, ix.sx.ptr, ix.sy.ptr, ix.sz.ptr, k, include_match, d_out.ptr);
        else {
            DeviceBuffer<double> scratch;
            GPP_TRY(scratch.alloc((size_t) nq * k));
            GPP_LAUNCH(knn_any_kernel, blocks_for(nq, 128), 128, 0, 0, dqx.ptr, dqy.ptr, dqz.ptr, nq, ix.geom, ix.cell_start.ptr, ix.order.ptr,
                       ix.sx.ptr, ix.sy.ptr, ix.sz.ptr, k, include_match, scratch.ptr, d_out.ptr);
            GPP_CUDA(cudaStreamSynchronize(0));   // the scratch goes out of scope
        }
    }
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}
}  // namespace

extern "C" {

int gpp_points_create(const float* lats, const float* lons, const float* elevs, const float* lafs, int n, int coordinate_type,
                      gpp_points** out) {
    if(!out) return fail(GPP_ERR_INVALID_ARGUMENT, "out must not be NULL");
    *out = nullptr;
    if(n < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative number of points");
    if(coordinate_type != GPP_GEODETIC && coordinate_type != GPP_CARTESIAN)
        return fail(GPP_ERR_INVALID_ARGUMENT, "unknown coordinate type %d", coordinate_type);
    gpp_points* p = new(std::nothrow) gpp_points();
    if(!p) return fail(GPP_ERR_RUNTIME, "out of memory");
    p->n = n;
    p->type = coordinate_type;
    p->lats.assign(lats, lats + n);
    p->lons.assign(lons, lons + n);
    // points.cpp:23-30 / grid.cpp:41-54: missing elevations and land fractions are NaN
    if(elevs) p->elevs.assign(elevs, elevs + n); else p->elevs.assign(n, NAN);
    if(lafs) p->lafs.assign(lafs, lafs + n); else p->lafs.assign(n, NAN);
    p->has_elevs = elevs != nullptr;
    p->has_lafs = lafs != nullptr;
    p->x.resize(n); p->y.resize(n); p->z.resize(n);
    int bad = -1;
    #pragma omp parallel for
    for(int i = 0; i < n; i++)
        if(!convert_one(lats[i], lons[i], coordinate_type, p->x[i], p->y[i], p->z[i])) bad = i;
    if(bad >= 0) {
        int rc = fail(GPP_ERR_INVALID_ARGUMENT, "Invalid coords: %g,%g", lats[bad], lons[bad]);   // util.cpp:596-600
        delete p;
        return rc;
    }
    const std::vector<float>* co[3] = {&p->x, &p->y, &p->z};
    for(int d = 0; d < 3; d++) {
        float l = INFINITY, h = -INFINITY;
        #pragma omp parallel for reduction(min : l) reduction(max : h)
        for(int i = 0; i < n; i++) { l = std::min(l, (*co[d])[i]); h = std::max(h, (*co[d])[i]); }
        p->lo[d] = n > 0 ? l : 0.f;
        p->hi[d] = n > 0 ? h : 0.f;
    }
    *out = p;
    return GPP_OK;
}

void gpp_points_destroy(gpp_points* p) { delete p; }
int gpp_convert_coordinates(const float* lats, const float* lons, int n, int coordinate_type, float* x, float* y, float* z) {
    if(n < 0 || (n > 0 && (!lats || !lons || !x || !y || !z))) return fail(GPP_ERR_INVALID_ARGUMENT, "invalid arguments");
    if(coordinate_type != GPP_GEODETIC && coordinate_type != GPP_CARTESIAN)
        return fail(GPP_ERR_INVALID_ARGUMENT, "unknown coordinate type %d", coordinate_type);
    for(int i = 0; i < n; i++)
        if(!convert_one(lats[i], lons[i], coordinate_type, x[i], y[i], z[i]))
            return fail(GPP_ERR_INVALID_ARGUMENT, "Invalid coords: %g,%g", lats[i], lons[i]);   // util.cpp:596-600
    return GPP_OK;
}
int gpp_points_set_shape(gpp_points* p, int ny, int nx) {
    if(!p) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    if(ny < 0 || nx < 0 || (long long) ny * nx != p->n) return fail(GPP_ERR_INVALID_ARGUMENT, "shape %d x %d does not match %d points", ny, nx, p->n);
    p->shape_ny = ny;
    p->shape_nx = nx;
    return GPP_OK;
}
int gpp_points_size(const gpp_points* p) { return p ? p->n : 0; }
int gpp_points_coordinate_type(const gpp_points* p) { return p ? p->type : GPP_GEODETIC; }

int gpp_points_get_xyz(const gpp_points* p, float* x, float* y, float* z) {
    if(!p) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    if(x) std::memcpy(x, p->x.data(), sizeof(float) * p->n);
    if(y) std::memcpy(y, p->y.data(), sizeof(float) * p->n);
    if(z) std::memcpy(z, p->z.data(), sizeof(float) * p->n);
    return GPP_OK;
}

int gpp_points_nearest_host(const gpp_points* cp, const float* qlats, const float* qlons, int nq, int include_match,
                            int* out_index) {
    if(!cp) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    gpp_points* p = const_cast<gpp_points*>(cp);
    GPP_TRY(ensure_device());
    if(p->n == 0) {   // points.cpp:56-62: no neighbours -> -1
        std::vector<float> x, y, z;
        GPP_TRY(convert_queries(qlats, qlons, nq, p->type, x, y, z));
        for(int q = 0; q < nq; q++) out_index[q] = -1;
        return GPP_OK;
    }
    DeviceBuffer<int> d_out;
    GPP_TRY(run_knn(p, qlats, qlons, nq, 1, include_match, d_out));
    GPP_TRY(d_out.download(out_index, nq));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_points_closest_host(const gpp_points* cp, const float* qlats, const float* qlons, int nq, int num, int include_match,
                            int* out_index) {
    if(!cp) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    if(num < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "num must be >= 0");
    gpp_points* p = const_cast<gpp_points*>(cp);
    GPP_TRY(ensure_device());
    if(num == 0 || nq == 0) return GPP_OK;
    if(p->n == 0) {
        for(size_t i = 0; i < (size_t) nq * num; i++) out_index[i] = -1;
        return GPP_OK;
    }
    DeviceBuffer<int> d_out;
    GPP_TRY(run_knn(p, qlats, qlons, nq, num, include_match, d_out));
    GPP_TRY(d_out.download(out_index, (size_t) nq * num));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_points_neighbours_host(const gpp_points* cp, const float* qlats, const float* qlons, const float* radii, int nq,
                               int include_match, int capacity, int* out_index, float* out_dist, int* out_count) {
    if(!cp) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    if(capacity < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "capacity must be >= 0");
    gpp_points* p = const_cast<gpp_points*>(cp);
    GPP_TRY(ensure_device());
    std::vector<float> qx, qy, qz;
    GPP_TRY(convert_queries(qlats, qlons, nq, p->type, qx, qy, qz));
    if(nq == 0) return GPP_OK;
    if(p->n == 0) {
        for(int q = 0; q < nq; q++) out_count[q] = 0;
        return GPP_OK;
    }
    GPP_TRY(p->ensure_index());
    DeviceBuffer<float> dqx, dqy, dqz, dr, ddist;
    DeviceBuffer<int> dindex, dcount;
    GPP_TRY(dqx.upload(qx.data(), nq));
    GPP_TRY(dqy.upload(qy.data(), nq));
    GPP_TRY(dqz.upload(qz.data(), nq));
    GPP_TRY(dr.upload(radii, nq));
    GPP_TRY(dcount.alloc(nq));
    bool store = capacity > 0 && out_index;
    if(store) {
        GPP_TRY(dindex.alloc((size_t) nq * capacity));
        GPP_CUDA(cudaMemsetAsync(dindex.ptr, 0xFF, sizeof(int) * (size_t) nq * capacity, 0));   // unused slots read -1
    }
    const CellIndex& ix = p->index;
    GPP_LAUNCH(radius_kernel, blocks_for(nq, 128), 128, 0, 0, dqx.ptr, dqy.ptr, dqz.ptr, dr.ptr, nq, ix.geom, ix.cell_start.ptr,
               ix.order.ptr, ix.sx.ptr, ix.sy.ptr, ix.sz.ptr, include_match, capacity, store ? dindex.ptr : nullptr, nullptr,
               dcount.ptr);
    GPP_TRY(dcount.download(out_count, nq));
    if(store) {
        GPP_TRY(dindex.download(out_index, (size_t) nq * capacity));
        if(out_dist) {
            GPP_TRY(ddist.alloc((size_t) nq * capacity));
            GPP_CUDA(cudaMemsetAsync(ddist.ptr, 0xFF, sizeof(float) * (size_t) nq * capacity, 0));   // unused slots read NaN
            GPP_LAUNCH(radius_dist_kernel, blocks_for((size_t) nq * capacity, 256), 256, 0, 0, dqx.ptr, dqy.ptr, dqz.ptr, nq,
                       p->dx.ptr, p->dy.ptr, p->dz.ptr, capacity, dindex.ptr, dcount.ptr, ddist.ptr);
            GPP_TRY(ddist.download(out_dist, (size_t) nq * capacity));
        }
    }
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_nearest_host(const gpp_points* cp, const float* qlats, const float* qlons, int nq, const float* ivalues, int n_fields,
                     float* out) {
    if(!cp) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    if(n_fields < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "n_fields must be >= 0");
    gpp_points* p = const_cast<gpp_points*>(cp);
    GPP_TRY(ensure_device());
    if(p->n == 0) {   // nearest.cpp:17-18 etc.: an empty input set gives missing values
        std::vector<float> x, y, z;
        GPP_TRY(convert_queries(qlats, qlons, nq, p->type, x, y, z));
        for(size_t i = 0; i < (size_t) nq * n_fields; i++) out[i] = NAN;
        return GPP_OK;
    }
    if(nq == 0 || n_fields == 0) return GPP_OK;
    DeviceBuffer<int> d_index;
    GPP_TRY(run_knn(p, qlats, qlons, nq, 1, 1, d_index));
    DeviceBuffer<float> d_values, d_out;
    GPP_TRY(d_values.upload(ivalues, (size_t) n_fields * p->n));
    GPP_TRY(d_out.alloc((size_t) n_fields * nq));
    GPP_LAUNCH(gather_kernel, blocks_for((size_t) nq * n_fields, 256), 256, 0, 0, d_values.ptr, p->n, d_index.ptr, nq, n_fields,
               d_out.ptr);
    GPP_TRY(d_out.download(out, (size_t) n_fields * nq));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

}  // extern "C"
