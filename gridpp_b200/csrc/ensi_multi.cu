// The "multi" ensemble statistical interpolation variants with member-by-member increments, and the static correlations
// between two point sets:
//   gridpp::optimal_interpolation_ensi_multi_ebe    src/api/oi_ensi_multi.cpp:329-627   (ensemble-based correlations)
//   gridpp::optimal_interpolation_ensi_multi_ebesc  src/api/oi_ensi_multi.cpp:630-859   (static correlations)
//   gridpp::staticcorr_points                       src/api/corr_points.cpp:26-131
// (the third variant, _utem, is a mode of the EnSI kernel: ensi.cu.)
//
// One warp per background point, like the other OI kernels: the observations inside the localization radius are selected
// from the bucket grid (gather_candidates) and put in index order. The k x k matrix  A = C o (Z Z') + diag(pratios)  (C:
// structure function between the observations; Z: the standardised perturbations, ebe only; not assumed symmetric:
// MultipleStructure::corr is not) depends on the selected SET only, so  A^-1 Innov  (k x E) is computed once per set -- Gaussian
// elimination with partial pivoting on [A | Innov] in shared memory, oi.cuh ge_solve -- and kept while consecutive points select
// the same set; every point then forms  background + bratio * r (A^-1 Innov)  with its own correlations r (the reference's
// K = r inv(A), dx = bratio * K Innov, re-associated).
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "oi.cuh"

using namespace gpp;

extern "C" int gpp_ensi_valid_members_device(const float* d_background, long long n_points, int nE, int* member_valid, void* stream);

namespace {

constexpr int EM_NSLOT = 5;                       // candidate buffer = 160 entries
constexpr int EM_KMAX = 32 * (EM_NSLOT - 1);      // observations per point
constexpr int EM_GRAB = 32;                       // consecutive points a warp takes per grab of the work counter
constexpr int EM_CHUNKS = 8;                       // blocks of a large field in flight through upload / analysis / download
constexpr int EM_MAX_COLS = 256;                  // observations per point + valid members (the columns of the augmented matrix)

struct EmParams {
    const float *gx, *gy, *gz, *gelev, *glaf;     // background points
    const float* bratios;                          // [nB]
    const float* background;                       // [nB][nE]
    const float* background_corr;                  // [nB][nE], ebe only
    float* analysis;                               // [nB][nE], holds the background on entry
    const int* valid_ens;                          // [E]
    int nB, nE, E;
    int first, count;                              // the points analysed by this launch: [first, first + count)
    ObsView obs;                                   // ratio = pratios
    const float* gI;                               // [slot][E]: pobs - pbackground of the valid members (float, :557,:806)
    const float* gZ;                               // [slot][E]: standardised perturbations at the observations (ebe, :421-445)
    gpp_structure s;
    float R;
    int k;                                         // observations per point, <= EM_KMAX
    int ldm, ldz;                                  // leading dimensions of M (doubles) and Z (floats)
    int smem_per_warp;
    int o_pos, o_spos, o_prev, o_pt, o_x, o_xl, o_val, o_z, o_mx, o_m;   // byte offsets into a warp's shared memory (key at 0)
    int allow_extrapolation;
    int* work_counter;
    int* singular;                                 // set when a pivot vanishes (the reference's inv() throws)
};

struct EmSmem {
    unsigned long long* key;
    int *pos, *spos, *prev;
    float *sx, *sy, *sz, *selev, *slaf;
    double *x, *xl, *mx, *mn;
    float *val, *Z;
    double* M;
    __device__ void bind(unsigned char* base, const EmParams& P) {
        key = reinterpret_cast<unsigned long long*>(base);
        pos = reinterpret_cast<int*>(base + P.o_pos);
        spos = reinterpret_cast<int*>(base + P.o_spos);
        prev = reinterpret_cast<int*>(base + P.o_prev);
        sx = reinterpret_cast<float*>(base + P.o_pt);
        sy = sx + P.k; sz = sy + P.k; selev = sz + P.k; slaf = selev + P.k;
        x = reinterpret_cast<double*>(base + P.o_x);
        xl = reinterpret_cast<double*>(base + P.o_xl);
        val = reinterpret_cast<float*>(base + P.o_val);
        Z = reinterpret_cast<float*>(base + P.o_z);
        mx = reinterpret_cast<double*>(base + P.o_mx);
        mn = mx + P.E;
        M = reinterpret_cast<double*>(base + P.o_m);
    }
};

size_t em_layout(EmParams& P, bool with_ens) {
    auto up = [](size_t v) { return (v + 15) & ~(size_t) 15; };
    const size_t K = (size_t) P.k, E = (size_t) P.E;
    size_t o = up(sizeof(unsigned long long) * 32 * EM_NSLOT);
    P.o_pos = (int) o; o = up(o + sizeof(int) * 32 * EM_NSLOT);
    P.o_spos = (int) o; o = up(o + sizeof(int) * K);
    P.o_prev = (int) o; o = up(o + sizeof(int) * K);
    P.o_pt = (int) o; o = up(o + sizeof(float) * 5 * K);
    P.o_x = (int) o; o = up(o + sizeof(double) * K);
    P.o_xl = (int) o; o = up(o + sizeof(double) * (with_ens ? E : 0));
    P.o_val = (int) o; o = up(o + sizeof(float) * (with_ens ? E : 0));
    P.ldz = (int) (E | 1);
    P.o_z = (int) o; o = up(o + sizeof(float) * (with_ens ? K * P.ldz : 0));
    P.o_mx = (int) o; o = up(o + sizeof(double) * 2 * E);
    P.ldm = (int) ((K + E) | 1);
    P.o_m = (int) o; o = up(o + sizeof(double) * K * P.ldm);
    return o;
}

// candidates not cut to max_points keep the order of the radius query (ascending original index); a cut leaves them best first
template <int NSLOT>
__device__ __forceinline__ void order_by_index(unsigned long long* key, int* pos, int k, int lane) {
    unsigned long long kk[NSLOT - 1];
    int pp[NSLOT - 1], rr[NSLOT - 1];
    #pragma unroll
    for(int t = 0; t < NSLOT - 1; t++) {
        const int i = lane + 32 * t;
        kk[t] = i < k ? key[i] : 0ull;
        pp[t] = i < k ? pos[i] : 0;
        rr[t] = 0;
    }
    for(int j = 0; j < k; j++) {
        const unsigned lw = (unsigned) key[j];   // 0x7fffffff - original index
        #pragma unroll
        for(int t = 0; t < NSLOT - 1; t++) rr[t] += lw > (unsigned) kk[t];
    }
    __syncwarp();
    #pragma unroll
    for(int t = 0; t < NSLOT - 1; t++)
        if(lane + 32 * t < k) { key[rr[t]] = kk[t]; pos[rr[t]] = pp[t]; }
    __syncwarp();
}

// ge_solve with the number of 32-column register chunks the k + nrhs columns need
__device__ __forceinline__ bool solve_columns(double* M, int k, int ld, int nrhs, int lane) {
    const int cols = k + nrhs;
    if(cols <= 32) return ge_solve<1>(M, k, ld, nrhs, lane);
    if(cols <= 64) return ge_solve<2>(M, k, ld, nrhs, lane);
    if(cols <= 96) return ge_solve<3>(M, k, ld, nrhs, lane);
    if(cols <= 160) return ge_solve<5>(M, k, ld, nrhs, lane);
    return ge_solve<EM_MAX_COLS / 32>(M, k, ld, nrhs, lane);
}

// A = C [o Z Z'] + diag(pratios) depends on the SET of selected observations only, and so does A^-1 Innov (k x E). The gain of a
// point is K = r A^-1 with r its own correlations, hence dx = bratio * r (A^-1 Innov): a k-term dot product per member once the
// set's system has been solved with the E innovation columns as right-hand sides. Neighbouring points mostly select the same
// set (the selection is put in index order so that it can be compared), so most points skip the assembly and the solve --
// the reuse of the deterministic OI kernel with E right-hand sides.
// NSLOT: the candidate buffer holds 32 NSLOT entries and the selection at most 32 (NSLOT - 1): 2 for the usual max_points <= 32
// (a rank selection over 64 entries instead of 160), 3 up to 64, EM_NSLOT up to 128
template <int SMODE, bool WITH_ENS, int NSLOT>
__global__ void __launch_bounds__(64) ensi_multi_kernel(const __grid_constant__ EmParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmSmem S;
    S.bind(smem_raw + (size_t) (threadIdx.x >> 5) * P.smem_per_warp, P);
    const int lane = (int) lane_id();
    const CandBuf cb = {S.key, S.pos};
    const int E = P.E, ld = P.ldm, ldz = P.ldz;
    const int n_blk = (P.count + EM_GRAB - 1) / EM_GRAB;
    int prev_k = -1;   // S.prev holds the set whose A^-1 Innov is in the right-hand-side columns of S.M
    for(;;) {
        int blk = 0;
        if(lane == 0) blk = atomicAdd(P.work_counter, 1);
        blk = __shfl_sync(0xffffffffu, blk, 0);
        if(blk >= n_blk) break;
        const int g_end = P.first + min((blk + 1) * EM_GRAB, P.count);
        for(int g = P.first + blk * EM_GRAB; g < g_end; g++) {
            const Pt p1 = {P.gx[g], P.gy[g], P.gz[g], P.gelev[g], P.glaf[g]};
            const int k = gather_candidates<SMODE, NSLOT>(P.obs, P.s, p1, P.R, P.k, cb);
            if(k == 0) continue;   // :462-465,:511-514: too few observations, keep the background
            order_by_index<NSLOT>(S.key, S.pos, k, lane);
            bool same = k == prev_k;
            if(same) {
                bool eq = true;
                for(int i = lane; i < k; i += 32) eq = eq && cand_key_orig(S.key[i]) == S.prev[i];
                same = __all_sync(0xffffffffu, eq);
            }
            if(!same) {
                prev_k = -1;
                for(int i = lane; i < k; i += 32) {
                    const int pos = S.pos[i];
                    S.spos[i] = pos;
                    S.prev[i] = cand_key_orig(S.key[i]);
                    S.sx[i] = P.obs.x[pos]; S.sy[i] = P.obs.y[pos]; S.sz[i] = P.obs.z[pos];
                    S.selev[i] = P.obs.elev[pos]; S.slaf[i] = P.obs.laf[pos];
                }
                __syncwarp();
                if(WITH_ENS)
                    for(int idx = lane; idx < k * E; idx += 32) {
                        const int i = idx / E, e = idx - i * E;
                        S.Z[i * ldz + e] = P.gZ[(size_t) S.spos[i] * E + e];
                    }
                __syncwarp();
                // ---- A(i, j) = corr(p_i, p_j) [* (Z Z')(i, j)] + (i == j) pratios_i  (:547-575, :797-818)
                for(int idx = lane; idx < k * k; idx += 32) {
                    const int i = idx / k, j = idx - i * k;
                    const Pt a = {S.sx[i], S.sy[i], S.sz[i], S.selev[i], S.slaf[i]};
                    const Pt b = {S.sx[j], S.sy[j], S.sz[j], S.selev[j], S.slaf[j]};
                    const float hdist = straight_distance(a.x, a.y, a.z, b.x, b.y, b.z);
                    double v = (double) corr_call<SMODE>(P.s, a, b, hdist);
                    if(WITH_ENS) {
                        double zz = 0.0;
                        for(int e = 0; e < E; e++) zz = __dadd_rn(zz, __dmul_rn((double) S.Z[i * ldz + e], (double) S.Z[j * ldz + e]));
                        v = __dmul_rn(v, zz);
                    }
                    if(i == j) v = __dadd_rn(v, (double) P.obs.ratio[S.spos[i]]);
                    S.M[i * ld + j] = v;
                }
                // ---- the right-hand sides: lInnov (k x E, :557, :806), and its column extremes for the clamp (:584-585)
                for(int e = lane; e < E; e += 32) {
                    double mx = -INFINITY, mn = INFINITY;
                    for(int i = 0; i < k; i++) {
                        const double innov = (double) P.gI[(size_t) S.spos[i] * E + e];
                        S.M[i * ld + k + e] = innov;
                        mx = innov > mx ? innov : mx;
                        mn = innov < mn ? innov : mn;
                    }
                    S.mx[e] = mx;
                    S.mn[e] = mn;
                }
                __syncwarp();
                if(!solve_columns(S.M, k, ld, E, lane)) {
                    if(lane == 0) atomicExch(P.singular, 1);
                    continue;
                }
                prev_k = k;
            }
            // ---- r: the point's correlations with the selection, [o X_L Z'] (:552, :572, :802)
            if(WITH_ENS) {
                // lX_L (:524-536): the standardised perturbations of background_corr at this point
                for(int e = lane; e < E; e += 32) S.val[e] = P.background_corr[(size_t) g * P.nE + P.valid_ens[e]];
                __syncwarp();
                float mean, sd;
                seq_mean_std(S.val, E, &mean, &sd);
                const bool use = is_valid(mean) && is_valid(sd) && sd > 0.0013f;
                const double cf = 1.0 / sqrt((double) (E - 1));
                for(int e = lane; e < E; e += 32)
                    S.xl[e] = use ? __ddiv_rn(__dmul_rn(cf, (double) __fsub_rn(S.val[e], mean)), (double) sd) : 0.0;
                __syncwarp();
            }
            for(int i = lane; i < k; i += 32) {
                double r = (double) cand_key_rho(S.key[i]);
                if(WITH_ENS) {
                    double xz = 0.0;
                    for(int e = 0; e < E; e++) xz = __dadd_rn(xz, __dmul_rn(S.xl[e], (double) S.Z[i * ldz + e]));
                    r = __dmul_rn(r, xz);
                }
                S.x[i] = r;
            }
            __syncwarp();
            // ---- dx = bratio * r (A^-1 lInnov) per member, the anti-extrapolation filter, analysis = background + dx (:577-613)
            const double ratio = (double) P.bratios[g];
            for(int e = lane; e < E; e += 32) {
                double acc = 0.0;
                for(int i = 0; i < k; i++) acc = fma(S.x[i], S.M[i * ld + k + e], acc);
                double dx = __dmul_rn(ratio, acc);
                if(!P.allow_extrapolation) {
                    float increment = (float) dx;
                    const float maxInc = (float) S.mx[e], minInc = (float) S.mn[e];
                    if(maxInc > 0 && increment > maxInc) increment = maxInc;
                    else if(maxInc < 0 && increment > 0) increment = 0;
                    else if(minInc < 0 && increment < minInc) increment = minInc;
                    else if(minInc > 0 && increment < 0) increment = 0;
                    dx = (double) increment;
                }
                const size_t o = (size_t) g * P.nE + P.valid_ens[e];
                P.analysis[o] = (float) __dadd_rn((double) P.background[o], dx);
            }
            __syncwarp();
        }
    }
}

// corr_points.cpp:61-125: one warp per point, the selected knots' correlations scattered into the (zeroed) row
template <int SMODE>
__global__ void __launch_bounds__(128) staticcorr_kernel(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                                                         const float* __restrict__ gelev, const float* __restrict__ glaf, int nY, ObsView knots,
                                                         const __grid_constant__ gpp_structure s, float R, int k, float* __restrict__ out, int nS) {
    __shared__ unsigned long long key[4][32 * EM_NSLOT];
    __shared__ int pos[4][32 * EM_NSLOT];
    const int w = threadIdx.x >> 5, lane = (int) lane_id();
    const CandBuf cb = {key[w], pos[w]};
    for(int y = blockIdx.x * 4 + w; y < nY; y += gridDim.x * 4) {
        const Pt p1 = {gx[y], gy[y], gz[y], gelev[y], glaf[y]};
        const int n = gather_candidates<SMODE, EM_NSLOT>(knots, s, p1, R, k, cb);
        for(int i = lane; i < n; i += 32) out[(size_t) y * nS + cand_key_orig(key[w][i])] = cand_key_rho(key[w][i]);
        __syncwarp();
    }
}

// max_points == 0: no selection, every knot inside the radius with rho > 0 is written (any number of them)
template <int SMODE>
__global__ void __launch_bounds__(128) staticcorr_all_kernel(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                                                             const float* __restrict__ gelev, const float* __restrict__ glaf, int nY, ObsView obs,
                                                             const __grid_constant__ gpp_structure s, float R, float* __restrict__ out, int nS) {
    const int w = threadIdx.x >> 5, lane = (int) lane_id();
    for(int yy = blockIdx.x * 4 + w; yy < nY; yy += gridDim.x * 4) {
        const Pt p1 = {gx[yy], gy[yy], gz[yy], gelev[yy], glaf[yy]};
        const float lo0 = __fsub_rn(p1.x, R), lo1 = __fsub_rn(p1.y, R), lo2 = __fsub_rn(p1.z, R);
        const float hi0 = __fadd_rn(p1.x, R), hi1 = __fadd_rn(p1.y, R), hi2 = __fadd_rn(p1.z, R);
        if(!(lo0 < hi0 && lo1 < hi1 && lo2 < hi2)) continue;
        const int cx0 = cell_coord(obs.geom, 0, lo0), cx1 = cell_coord(obs.geom, 0, hi0);
        const int cy0 = cell_coord(obs.geom, 1, lo1), cy1 = cell_coord(obs.geom, 1, hi1);
        const int cz0 = cell_coord(obs.geom, 2, lo2), cz1 = cell_coord(obs.geom, 2, hi2);
        for(int cz = cz0; cz <= cz1; cz++)
            for(int cy = cy0; cy <= cy1; cy++) {
                const int base = (cz * obs.geom.n[1] + cy) * obs.geom.n[0];
                const int s0 = obs.cell_start[base + cx0], s1 = obs.cell_start[base + cx1 + 1];
                for(int i = s0 + lane; i < s1; i += 32) {
                    const float ox = obs.x[i], oy = obs.y[i], oz = obs.z[i];
                    if(!(ox > lo0 && ox < hi0 && oy > lo1 && oy < hi1 && oz > lo2 && oz < hi2)) continue;
                    const float dist = straight_distance(ox, oy, oz, p1.x, p1.y, p1.z);
                    if(!(dist <= R)) continue;
                    const Pt p2 = {ox, oy, oz, obs.elev[i], obs.laf[i]};
                    const float rho = corr_background_mode<SMODE>(s, p1, p2, dist);
                    if(rho > 0.f) out[(size_t) yy * nS + obs.orig[i]] = rho;
                }
            }
    }
}

// calc_statistic Mean / Std on the host (util.cpp:19-73) for the observation-side tables
void host_mean_std(const float* v, int n, float* mean, float* sd) {
    float total = 0;
    int count = 0;
    for(int i = 0; i < n; i++)
        if(is_valid(v[i])) { total += v[i]; count++; }
    *mean = count > 0 ? total / count : NAN;
    float t1 = 0, t2 = 0, K = NAN;
    count = 0;
    for(int i = 0; i < n; i++) {
        if(!is_valid(v[i])) continue;
        if(!is_valid(K)) K = v[i];
        const float d = v[i] - K;
        t1 += d;
        t2 += d * d;
        count++;
    }
    *sd = NAN;
    if(count > 0) {
        const float m1 = t1 / count, m2 = t2 / count;
        float var = m2 - m1 * m1;
        if(var < 0) var = 0;
        *sd = std::sqrt(var);
    }
}

int multi_host(const gpp_points* cbp, const float* bratios, const float* background, const float* background_corr, int nE,
               const gpp_points* opoints, const float* pobs, const float* pratios, const float* pbackground, const float* pbackground_corr,
               const gpp_structure* structure, int max_points, int allow_extrapolation, float* analysis, bool with_ens) {
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");   // :341-342,:640-641
    if(!cbp || !opoints || !structure) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(nE < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative ensemble size");
    gpp_points* bp = const_cast<gpp_points*>(cbp);
    const int nB = bp->n, nS = opoints->n;
    const size_t nBE = (size_t) nB * nE;
    if(bp->type != opoints->type)
        return fail(GPP_ERR_INVALID_ARGUMENT, "Both background and observations points must be of same coorindate type (lat/lon or x/y)");
    if(nS == 0 || nBE == 0) {   // :362-364,:656-658
        if(nBE) std::memcpy(analysis, background, sizeof(float) * nBE);
        return GPP_OK;
    }
    if(!bratios || !background || !pobs || !pratios || !pbackground || (with_ens && (!background_corr || !pbackground_corr)))
        return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    GPP_TRY(reject_unset_scales(structure));
    GPP_TRY(ensure_device());
    Trace trace("optimal_interpolation_ensi_multi_host");
    // ---- members valid everywhere (:395-419,:687-709). Large fields are scanned by host threads so that their upload can be
    // pipelined with the analysis block by block; small ones are uploaded at once and scanned on the device.
    DeviceBuffer<float> d_bg, d_bgc, d_out, d_br;
    const int n_chunks = nB >= (1 << 18) ? EM_CHUNKS : 1;
    std::vector<int> ok(nE, 1), ok2(nE, 1);
    if(n_chunks > 1) {
        GPP_TRY(d_bg.alloc(nBE));
        if(with_ens) GPP_TRY(d_bgc.alloc(nBE));
        std::vector<unsigned char> bad((size_t) nE, 0), bad2((size_t) nE, 0);
        const int scan_threads = std::max(1, std::min(8, omp_get_num_procs()));   // (explicit: launchers export OMP_NUM_THREADS=1)
        #pragma omp parallel num_threads(scan_threads)
        {
            std::vector<unsigned char> mine((size_t) nE, 0), mine2((size_t) nE, 0);
            #pragma omp for schedule(static) nowait
            for(long long p = 0; p < (long long) nB; p++) {
                const float* row = background + (size_t) p * nE;
                for(int e = 0; e < nE; e++) mine[e] |= (unsigned char) !is_valid(row[e]);
                if(with_ens) {
                    const float* row2 = background_corr + (size_t) p * nE;
                    for(int e = 0; e < nE; e++) mine2[e] |= (unsigned char) !is_valid(row2[e]);
                }
            }
            #pragma omp critical
            for(int e = 0; e < nE; e++) { bad[e] |= mine[e]; bad2[e] |= mine2[e]; }
        }
        for(int e = 0; e < nE; e++) { ok[e] = !bad[e]; ok2[e] = !bad2[e]; }
    }
    else {
        GPP_TRY(d_bg.upload(background, nBE));
        GPP_TRY(gpp_ensi_valid_members_device(d_bg.ptr, nB, nE, ok.data(), nullptr));
        if(with_ens) {
            GPP_TRY(d_bgc.upload(background_corr, nBE));
            GPP_TRY(gpp_ensi_valid_members_device(d_bgc.ptr, nB, nE, ok2.data(), nullptr));
        }
    }
    std::vector<int> valid_ens;
    for(int e = 0; e < nE; e++) {
        bool good = ok[e] && ok2[e];
        for(int i = 0; i < nS && good; i++)
            good = is_valid(pbackground[(size_t) i * nE + e]) && (!with_ens || is_valid(pbackground_corr[(size_t) i * nE + e]));
        if(good) valid_ens.push_back(e);
    }
    const int E = (int) valid_ens.size();
    if(trace.on) { cudaStreamSynchronize(0); trace.lap("H2D + valid-member scan"); }
    if(E == 0) {
        std::memcpy(analysis, background, sizeof(float) * nBE);
        return GPP_OK;
    }
    // The reference addresses its innovation matrix (lS x nValidEns) by the ORIGINAL member index (:557,:806), which is only
    // inside the matrix when the invalid members are the last ones; otherwise Armadillo's bounds check throws.
    for(int e = 0; e < E; e++)
        if(valid_ens[e] >= E) return fail(GPP_ERR_RUNTIME, "Mat::operator(): index out of bounds");
    // ---- observation table: pobs[index][0] decides (:480,:745)
    std::vector<char> valid(nS);
    std::vector<double> unused(nS, 0.0);
    std::vector<float> ratio(pratios, pratios + nS);
    for(int i = 0; i < nS; i++) valid[i] = is_valid(pobs[(size_t) i * nE]);
    gpp_oi_obs table;
    std::vector<int> order;
    GPP_TRY(build_obs_table(opoints, valid, unused, ratio, structure->term[0].loc_dist, &table, &order));
    if(table.n_valid == 0) {
        std::memcpy(analysis, background, sizeof(float) * nBE);
        return GPP_OK;
    }
    const size_t nV = order.size();
    std::vector<float> gI(nV * E), gZ(with_ens ? nV * E : 1), row(E);
    for(size_t slot = 0; slot < nV; slot++) {
        const size_t i = (size_t) order[slot];
        for(int e = 0; e < E; e++) gI[slot * E + e] = pobs[i * nE + valid_ens[e]] - pbackground[i * nE + valid_ens[e]];
        if(with_ens) {   // gZ_R, :421-445 (a float table in the reference)
            for(int e = 0; e < E; e++) row[e] = pbackground_corr[i * nE + valid_ens[e]];
            float mean, sd;
            host_mean_std(row.data(), E, &mean, &sd);
            const bool use = is_valid(mean) && is_valid(sd) && sd > 0.0013f;
            for(int e = 0; e < E; e++) gZ[slot * E + e] = use ? (float) (1 / std::sqrt((double) (E - 1)) * (row[e] - mean) / sd) : 0.f;
        }
    }
    DeviceBuffer<float> d_gI, d_gZ;
    DeviceBuffer<int> d_valid, d_flags;
    GPP_TRY(d_gI.upload(gI.data(), gI.size()));
    if(with_ens) GPP_TRY(d_gZ.upload(gZ.data(), gZ.size()));
    GPP_TRY(d_valid.upload(valid_ens.data(), (size_t) E));
    GPP_TRY(d_br.upload(bratios, (size_t) nB));
    GPP_TRY(d_out.alloc(nBE));
    if(n_chunks == 1) GPP_CUDA(cudaMemcpyAsync(d_out.ptr, d_bg.ptr, sizeof(float) * nBE, cudaMemcpyDeviceToDevice, 0));
    GPP_TRY(d_flags.alloc(1 + EM_CHUNKS));   // [0]: a pivot vanished somewhere; [1 + c]: the work counter of block c
    GPP_CUDA(cudaMemsetAsync(d_flags.ptr, 0, sizeof(int) * (1 + EM_CHUNKS), 0));
    GPP_TRY(bp->ensure_on_device());
    int kcap = max_points > 0 ? std::min(max_points, table.n_valid) : table.n_valid;
    if(max_points == 0 && kcap > EM_KMAX) {
        int hmax = 0;
        GPP_TRY(count_max_candidates(bp, 0, nB, nullptr, table.view(), structure->term[0].loc_dist, 0, &hmax));
        kcap = std::min(kcap, std::max(hmax, 1));
    }
    if(kcap > EM_KMAX)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "optimal_interpolation_ensi_multi supports at most %d observations per point on the device (got %d)", EM_KMAX, kcap);
    if(kcap + E > EM_MAX_COLS)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "optimal_interpolation_ensi_multi: observations per point + valid members must not exceed %d on the device (got %d + %d)",
                    EM_MAX_COLS, kcap, E);
    EmParams P;
    std::memset(&P, 0, sizeof(P));
    P.gx = bp->dx.ptr; P.gy = bp->dy.ptr; P.gz = bp->dz.ptr; P.gelev = bp->delev.ptr; P.glaf = bp->dlaf.ptr;
    P.bratios = d_br.ptr;
    P.background = d_bg.ptr;
    P.background_corr = d_bgc.ptr;
    P.analysis = d_out.ptr;
    P.valid_ens = d_valid.ptr;
    P.nB = nB; P.nE = nE; P.E = E;
    P.obs = table.view();
    P.gI = d_gI.ptr; P.gZ = d_gZ.ptr;
    P.s = *structure;
    P.R = structure->term[0].loc_dist;
    P.k = kcap;
    P.allow_extrapolation = allow_extrapolation;
    P.singular = d_flags.ptr;
    const size_t per_warp = em_layout(P, with_ens);
    P.smem_per_warp = (int) per_warp;
    const int warps = per_warp * 2 <= 200 * 1024 ? 2 : 1;
    const size_t smem = per_warp * warps;
    if(smem > 227 * 1024)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "optimal_interpolation_ensi_multi: %d observations per point x %d members do not fit in shared memory", kcap, E);
    const int mode = structure_mode(*structure);
    void (*kernel)(EmParams) = nullptr;
#define EM_PICK(NS) (with_ens ? (mode == 1 ? ensi_multi_kernel<1, true, NS> : ensi_multi_kernel<0, true, NS>) \
                              : (mode == 1 ? ensi_multi_kernel<1, false, NS> : ensi_multi_kernel<0, false, NS>))
    if(kcap <= 32) kernel = EM_PICK(2);
    else if(kcap <= 64) kernel = EM_PICK(3);
    else kernel = EM_PICK(EM_NSLOT);
#undef EM_PICK
    GPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 1;
    GPP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
    trace.lap("observation tables");
    // blocks of points: uploaded, analysed and returned as a pipeline (oi.cu pipelined_download); every block has its own work counter
    std::vector<size_t> bounds(n_chunks + 1);
    for(int c = 0; c <= n_chunks; c++) bounds[c] = (size_t) ((long long) nB * c / n_chunks) * nE;
    auto launch = [&](int c, cudaStream_t stream) {
        EmParams Q = P;
        Q.first = (int) (bounds[c] / nE);
        Q.count = (int) ((bounds[c + 1] - bounds[c]) / nE);
        Q.work_counter = d_flags.ptr + 1 + c;
        if(Q.count == 0) return (int) GPP_OK;
        if(n_chunks > 1) {
            const size_t n = bounds[c + 1] - bounds[c];
            GPP_CUDA(cudaMemcpyAsync(d_bg.ptr + bounds[c], background + bounds[c], sizeof(float) * n, cudaMemcpyHostToDevice, stream));
            GPP_CUDA(cudaMemcpyAsync(d_out.ptr + bounds[c], d_bg.ptr + bounds[c], sizeof(float) * n, cudaMemcpyDeviceToDevice, stream));
            if(with_ens)
                GPP_CUDA(cudaMemcpyAsync(d_bgc.ptr + bounds[c], background_corr + bounds[c], sizeof(float) * n, cudaMemcpyHostToDevice, stream));
        }
        const long long want = ((long long) (Q.count + EM_GRAB - 1) / EM_GRAB + warps - 1) / warps;
        const unsigned grid = (unsigned) std::max<long long>(1, std::min<long long>(want, (long long) sm_count() * std::max(per_sm, 1)));
        GPP_LAUNCH(kernel, grid, warps * 32, smem, stream, Q);
        return (int) GPP_OK;
    };
    if(n_chunks > 1) GPP_TRY(pipelined_download(bounds, launch, d_out.ptr, analysis, true));
    else {
        GPP_TRY(launch(0, 0));
        if(trace.on) { cudaStreamSynchronize(0); trace.lap("kernel"); }
        GPP_TRY(d_out.download(analysis, nBE));
    }
    int flags[1] = {0};
    GPP_CUDA(cudaMemcpyAsync(flags, d_flags.ptr, sizeof(flags), cudaMemcpyDeviceToHost, 0));
    GPP_CUDA(cudaStreamSynchronize(0));
    trace.lap(n_chunks > 1 ? "H2D + kernels + D2H (pipelined)" : "D2H");
    if(flags[0]) return fail(GPP_ERR_RUNTIME, "inv(): matrix is singular");
    return GPP_OK;
}

}  // namespace

extern "C" {

int gpp_optimal_interpolation_ensi_multi_ebe_host(const gpp_points* bpoints, const float* bratios, const float* background,
                                                  const float* background_corr, int nE, const gpp_points* opoints, const float* pobs,
                                                  const float* pratios, const float* pbackground, const float* pbackground_corr,
                                                  const gpp_structure* structure, int max_points, int allow_extrapolation, float* analysis) {
    return multi_host(bpoints, bratios, background, background_corr, nE, opoints, pobs, pratios, pbackground, pbackground_corr, structure,
                      max_points, allow_extrapolation, analysis, true);
}

int gpp_optimal_interpolation_ensi_multi_ebesc_host(const gpp_points* bpoints, const float* bratios, const float* background, int nE,
                                                    const gpp_points* opoints, const float* pobs, const float* pratios,
                                                    const float* pbackground, const gpp_structure* structure, int max_points,
                                                    int allow_extrapolation, float* analysis) {
    return multi_host(bpoints, bratios, background, nullptr, nE, opoints, pobs, pratios, pbackground, nullptr, structure, max_points,
                      allow_extrapolation, analysis, false);
}

int gpp_staticcorr_points_host(const gpp_points* cpoints, const gpp_points* knots, const gpp_structure* structure, int max_points, float* output) {
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");   // corr_points.cpp:34-35
    if(!cpoints || !knots || !structure) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(cpoints->type != knots->type)
        return fail(GPP_ERR_INVALID_ARGUMENT, "Both background grid and observations points must be of same coordinate type (lat/lon or x/y)");
    GPP_TRY(reject_unset_scales(structure));
    GPP_TRY(ensure_device());
    gpp_points* bp = const_cast<gpp_points*>(cpoints);
    const int nY = bp->n, nS = knots->n;
    const size_t n = (size_t) nY * nS;
    if(n == 0) return GPP_OK;
    if(max_points > EM_KMAX)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "staticcorr_points supports max_points of 0 (no limit) or at most %d on the device", EM_KMAX);
    std::vector<char> valid(nS, 1);
    std::vector<double> unused(nS, 0.0);
    std::vector<float> ratio(nS, 0.f);
    gpp_oi_obs table;
    GPP_TRY(build_obs_table(knots, valid, unused, ratio, structure->term[0].loc_dist, &table));
    GPP_TRY(bp->ensure_on_device());
    DeviceBuffer<float> d_out;
    GPP_TRY(d_out.alloc(n));
    GPP_CUDA(cudaMemsetAsync(d_out.ptr, 0, sizeof(float) * n, 0));
    const int mode = structure_mode(*structure);
    const float R = structure->term[0].loc_dist;
    const unsigned grid = (unsigned) std::max(1, std::min((nY + 3) / 4, sm_count() * 8));
    if(max_points == 0) {
        if(mode == 1) GPP_LAUNCH(staticcorr_all_kernel<1>, grid, 128, 0, 0, bp->dx.ptr, bp->dy.ptr, bp->dz.ptr, bp->delev.ptr, bp->dlaf.ptr, nY, table.view(), *structure, R, d_out.ptr, nS);
        else GPP_LAUNCH(staticcorr_all_kernel<0>, grid, 128, 0, 0, bp->dx.ptr, bp->dy.ptr, bp->dz.ptr, bp->delev.ptr, bp->dlaf.ptr, nY, table.view(), *structure, R, d_out.ptr, nS);
    }
    else {
        const int k = std::min(max_points, nS);
        if(mode == 1) GPP_LAUNCH(staticcorr_kernel<1>, grid, 128, 0, 0, bp->dx.ptr, bp->dy.ptr, bp->dz.ptr, bp->delev.ptr, bp->dlaf.ptr, nY, table.view(), *structure, R, k, d_out.ptr, nS);
        else GPP_LAUNCH(staticcorr_kernel<0>, grid, 128, 0, 0, bp->dx.ptr, bp->dy.ptr, bp->dz.ptr, bp->delev.ptr, bp->dlaf.ptr, nY, table.view(), *structure, R, k, d_out.ptr, nS);
    }
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

}  // extern "C"
