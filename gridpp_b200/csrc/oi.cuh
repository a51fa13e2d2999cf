// Observation-side state shared by the OI kernels (deterministic OI in oi.cu, EnSI in ensi.cu).
#pragma once

#include "points.cuh"
#include "structure.cuh"

namespace gpp {

// Device view of the observation table: VALID observations only, sorted by bucket-grid cell (stable, so
// ascending original index inside a cell). Passed to kernels by value.
struct ObsView {
    CellGeom geom;
    const int* cell_start;   // ncells + 1
    const float *x, *y, *z, *elev, *laf;
    const double* innov;     // (double) pobs - (double) pbackground            (oi.cpp:301-302,316)
    const float* ratio;      // obs_variance / bvariance_at_points, float       (oi.cpp:192-195)
    const int* orig;         // original observation index (selection tie-break, EnSI member lookup)
    const float *sh, *sv, *sw;   // spatially varying structure function: the scales at each observation (else NULL)
    int n;                   // number of valid observations
};

// Per-warp candidate buffer in shared memory (capacity 64). A candidate is a 64-bit key
//   (float bits of rho) << 32 | (0x7fffffff - original index)
// so that "better" (higher rho, then lower original index -- the streaming form of oi.cpp:262-273) is simply a
// larger key (rho > 0, so its bit pattern orders like the value), plus its slot in the sorted observation table.
struct CandBuf {
    unsigned long long* key;   // [64]
    int* pos;                  // [64]
};
__device__ __forceinline__ unsigned long long cand_key(float rho, int orig) {
    return ((unsigned long long) __float_as_uint(rho) << 32) | (unsigned) (0x7fffffff - orig);
}
__device__ __forceinline__ float cand_key_rho(unsigned long long key) { return __uint_as_float((unsigned) (key >> 32)); }
__device__ __forceinline__ int cand_key_orig(unsigned long long key) { return 0x7fffffff - (int) (unsigned) key; }
__device__ __forceinline__ bool cand_better(float r1, int o1, float r2, int o2) { return r1 > r2 || (r1 == r2 && o1 < o2); }

// Rank-select the best min(n, k) of the first n entries (n <= 32 * NSLOT), best first. Each lane owns entries
// lane, lane + 32, ...
template <int NSLOT>
__device__ __forceinline__ int cand_prune(const CandBuf& b, int n, int k) {
    const unsigned lane = lane_id();
    unsigned long long key[NSLOT];
    int pos[NSLOT], rank[NSLOT];
    #pragma unroll
    for(int t = 0; t < NSLOT; t++) {
        const bool has = (int) lane + 32 * t < n;
        key[t] = has ? b.key[lane + 32 * t] : 0ull;
        pos[t] = has ? b.pos[lane + 32 * t] : 0;
        rank[t] = 0;
    }
    if(n <= 32) {
        #pragma unroll 4
        for(int j = 0; j < n; j++) rank[0] += b.key[j] > key[0];
    }
    else {
        #pragma unroll 2
        for(int j = 0; j < n; j++) {
            const unsigned long long kj = b.key[j];
            #pragma unroll
            for(int t = 0; t < NSLOT; t++) rank[t] += kj > key[t];
        }
    }
    __syncwarp();
    #pragma unroll
    for(int t = 0; t < NSLOT; t++)
        if((int) lane + 32 * t < n && rank[t] < k) { b.key[rank[t]] = key[t]; b.pos[rank[t]] = pos[t]; }
    __syncwarp();
    return min(n, k);
}

// Scans the bucket-grid cells overlapping the localization box of p1 and keeps the k best observations
// (k <= 32 * (NSLOT - 1)). Warp-synchronous; every lane returns the same count. The buffer (capacity 32 * NSLOT)
// holds the selection on return, best-first when a prune ran. Invalid observations are not in the table.
template <int SMODE, int NSLOT>
__device__ __forceinline__ int gather_candidates(const ObsView& obs, const gpp_structure& s, const Pt& p1, float R, int k,
                                                 const CandBuf& b, bool* cut = nullptr) {
    const unsigned lane = lane_id();
    bool did_prune = false;   // more candidates than k: the selection was cut and is sorted best-first
    // kdtree.cpp:46-47: box corners are computed in float
    const float lo0 = __fsub_rn(p1.x, R), lo1 = __fsub_rn(p1.y, R), lo2 = __fsub_rn(p1.z, R);
    const float hi0 = __fadd_rn(p1.x, R), hi1 = __fadd_rn(p1.y, R), hi2 = __fadd_rn(p1.z, R);
    if(cut) *cut = false;
    if(!(lo0 < hi0 && lo1 < hi1 && lo2 < hi2)) return 0;
    const int cx0 = cell_coord(obs.geom, 0, lo0), cx1 = cell_coord(obs.geom, 0, hi0);
    const int cy0 = cell_coord(obs.geom, 1, lo1), cy1 = cell_coord(obs.geom, 1, hi1);
    const int cz0 = cell_coord(obs.geom, 2, lo2), cz1 = cell_coord(obs.geom, 2, hi2);
    int n = 0;
    for(int cz = cz0; cz <= cz1; cz++)
        for(int cy = cy0; cy <= cy1; cy++) {
            const int base = (cz * obs.geom.n[1] + cy) * obs.geom.n[0];
            const int s0 = obs.cell_start[base + cx0], s1 = obs.cell_start[base + cx1 + 1];
            for(int chunk = s0; chunk < s1; chunk += 32) {
                const int i = chunk + (int) lane;
                float rho = 0.f;
                bool ok = i < s1;
                if(ok) {
                    const float ox = obs.x[i], oy = obs.y[i], oz = obs.z[i];
                    // Boost within(): strictly inside the box (kdtree.cpp:46-53)
                    ok = ox > lo0 && ox < hi0 && oy > lo1 && oy < hi1 && oz > lo2 && oz < hi2;
                    if(ok) {
                        const float dist = straight_distance(ox, oy, oz, p1.x, p1.y, p1.z);
                        ok = dist <= R;   // within_radius, kdtree.cpp:247-260
                        if(ok) {
                            const Pt p2 = {ox, oy, oz, obs.elev[i], obs.laf[i]};
                            rho = corr_background_call<SMODE>(s, p1, p2, dist);   // oi.cpp:250
                            ok = rho > 0.f;                                        // oi.cpp:253
                        }
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, ok);
                if(ok) {
                    const int slot = n + __popc(mask & ((1u << lane) - 1u));
                    b.key[slot] = cand_key(rho, obs.orig[i]);
                    b.pos[slot] = i;
                }
                n += __popc(mask);
                __syncwarp();
                if(n > 32 * (NSLOT - 1)) { n = cand_prune<NSLOT>(b, n, k); did_prune = true; }   // keep room for the next 32
            }
        }
    if(n > k) { n = cand_prune<NSLOT>(b, n, k); did_prune = true; }   // final selection (oi.cpp:262-273)
    if(cut) *cut = did_prune;
    return n;
}

// calc_statistic(Mean) and (Std) of n valid floats (util.cpp:19-73): sequential float accumulation, the variance about the
// first value. Every lane may call it on the same shared array.
__device__ __forceinline__ void seq_mean_std(const float* v, int n, float* mean_out, float* std_out) {
    float total = 0.f;
    for(int e = 0; e < n; e++) total = __fadd_rn(total, v[e]);
    *mean_out = __fdiv_rn(total, (float) n);
    const float K = v[0];
    float t1 = 0.f, t2 = 0.f;
    for(int e = 0; e < n; e++) {
        const float d = __fsub_rn(v[e], K);
        t1 = __fadd_rn(t1, d);
        t2 = __fadd_rn(t2, __fmul_rn(d, d));
    }
    const float m1 = __fdiv_rn(t1, (float) n), m2 = __fdiv_rn(t2, (float) n);
    float var = __fsub_rn(m2, __fmul_rn(m1, m1));
    if(var < 0.f) var = 0.f;
    *std_out = __fsqrt_rn(var);
}


__device__ __forceinline__ double shfl_f64(double v, int src) {
    return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}

// Solves M X = B in place for one warp: M is k x k, row-major with leading dimension ld, in shared or global memory; the nrhs
// right-hand sides are columns k .. k + nrhs - 1 of the same rows and hold X on return. Gaussian elimination with partial
// pivoting (the largest |entry| of the column, lowest row on ties): lanes own columns, the pivot row is held in registers, the
// multipliers of a column are formed by all lanes at once and the rows below the pivot are then updated without any
// synchronisation between them; back substitution with the lanes on rows (few right-hand sides) or one column per lane (many).
// NT * 32 >= k + nrhs. Returns false when a pivot is
// zero or not finite (the reference's inv() throws).
template <int NT>
__device__ __forceinline__ bool ge_solve(double* M, int k, int ld, int nrhs, int lane) {
    const int last = k + nrhs - 1;                     // last column in use
    for(int c = 0; c < k; c++) {
        double best = -1.0;
        int bi = c;
        for(int i = c + lane; i < k; i += 32) {
            const double a = fabs(M[(size_t) i * ld + c]);
            if(a > best) { best = a; bi = i; }
        }
        #pragma unroll
        for(int off = 16; off > 0; off >>= 1) {
            const double ob = shfl_f64(best, lane ^ off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if(ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if(!(best > 0.0) || isinf(best)) return false;
        if(bi != c)
            for(int j = c + lane; j <= last; j += 32) {
                const double a = M[(size_t) c * ld + j], b = M[(size_t) bi * ld + j];
                M[(size_t) c * ld + j] = b;
                M[(size_t) bi * ld + j] = a;
            }
        __syncwarp();
        const double piv = M[(size_t) c * ld + c];
        for(int i = c + 1 + lane; i < k; i += 32) M[(size_t) i * ld + c] = M[(size_t) i * ld + c] / piv;   // the multipliers
        double prow[NT];
        #pragma unroll
        for(int t = 0; t < NT; t++) {
            const int j = c + 1 + lane + 32 * t;
            prow[t] = j <= last ? M[(size_t) c * ld + j] : 0.0;
        }
        __syncwarp();
        const int nt = (last - c + 31) / 32;           // column chunks still alive (columns c + 1 .. last)
        #pragma unroll 2
        for(int i = c + 1; i < k; i++) {
            const double f = M[(size_t) i * ld + c];
            #pragma unroll
            for(int t = 0; t < NT; t++) {
                const int j = c + 1 + lane + 32 * t;
                if(t < nt && j <= last) M[(size_t) i * ld + j] = fma(-f, prow[t], M[(size_t) i * ld + j]);
            }
        }
        __syncwarp();
    }
    if(nrhs >= 8) {
        // many right-hand sides: a lane takes a whole column and substitutes it back on its own -- no synchronisation, the row
        // of U is a broadcast read and the column entries of neighbouring lanes are neighbours in memory
        for(int r0 = 0; r0 < nrhs; r0 += 32) {
            const int r = r0 + lane;
            if(r < nrhs)
                for(int c = k - 1; c >= 0; c--) {
                    double acc = M[(size_t) c * ld + k + r];
                    for(int j = c + 1; j < k; j++) acc = fma(-M[(size_t) c * ld + j], M[(size_t) j * ld + k + r], acc);
                    M[(size_t) c * ld + k + r] = acc / M[(size_t) c * ld + c];
                }
        }
        __syncwarp();
        return true;
    }
    for(int c = k - 1; c >= 0; c--) {
        const double d = M[(size_t) c * ld + c];
        for(int r = 0; r < nrhs; r++) {
            const double xc = M[(size_t) c * ld + k + r] / d;
            __syncwarp();
            if(lane == 0) M[(size_t) c * ld + k + r] = xc;
            for(int i = lane; i < c; i += 32) M[(size_t) i * ld + k + r] = fma(-M[(size_t) i * ld + c], xc, M[(size_t) i * ld + k + r]);
        }
        __syncwarp();
    }
    return true;
}

}  // namespace gpp

constexpr unsigned OI_WORK_SLOTS = 4;

// Host-side owner of the observation table.
struct gpp_oi_obs {
    int n_total = 0;          // observations given
    int n_valid = 0;          // observations in the table
    float loc_dist = 0;       // localization distance the bucket grid was sized for
    gpp::CellGeom geom;
    int ncells = 0;
    gpp::DeviceBuffer<int> cell_start, orig;
    gpp::DeviceBuffer<float> x, y, z, elev, laf, ratio;
    gpp::DeviceBuffer<float> sh, sv, sw;   // only for spatially varying structure functions
    bool has_scales = false;
    gpp::DeviceBuffer<double> innov;
    // launch workspaces of the register OI kernel (gpp_oi_obs_create): OI_WORK_SLOTS slots handed out round-robin, so at most
    // that many analyses sharing this state may be in flight at once (more: gpp_optimal_interpolation_device_ws)
    gpp::DeviceBuffer<unsigned char> work;
    size_t work_slot_bytes = 0;
    mutable std::atomic<unsigned> work_next{0};
    gpp::ObsView view() const {
        gpp::ObsView v;
        v.geom = geom;
        v.cell_start = cell_start.ptr;
        v.x = x.ptr; v.y = y.ptr; v.z = z.ptr; v.elev = elev.ptr; v.laf = laf.ptr;
        v.innov = innov.ptr;
        v.ratio = ratio.ptr;
        v.orig = orig.ptr;
        v.sh = has_scales ? sh.ptr : nullptr; v.sv = has_scales ? sv.ptr : nullptr; v.sw = has_scales ? sw.ptr : nullptr;
        v.n = n_valid;
        return v;
    }
};

#include <functional>

namespace gpp {
// oi.cu: blocks [bounds[c], bounds[c+1]) of d_out are produced by launch(c) on the default stream and returned to host_out
// through a pinned staging buffer while later blocks are still being computed.
int pipelined_download(const std::vector<size_t>& bounds, const std::function<int(int, cudaStream_t)>& launch, const float* d_out, float* host_out,
                       bool overlap_blocks);
// Builds the table from host arrays. `valid[i]` selects the observations that enter the table.
// Largest number of table observations inside the localization radius of any of the background points
// [first, first+count) (points with an invalid d_background value are skipped when d_background is given).
int count_max_candidates(gpp_points* bp, int first, int count, const float* d_background, const ObsView& obs, float R,
                         cudaStream_t stream, int* out);
// `order`, when given, receives the original index of every table slot (what `orig` holds on the device).
int build_obs_table(const gpp_points* opoints, const std::vector<char>& valid, const std::vector<double>& innov,
                    const std::vector<float>& ratio, float loc_dist, gpp_oi_obs* out, std::vector<int>* order = nullptr,
                    const std::vector<float>* scales = nullptr /* h, v, w per observation, original order */);
}
