// Ensemble optimal interpolation (EnSI, Lussana et al. 2019) on the device ("K7"/"K8" in SURVEY.md).
// Replaces gridpp::optimal_interpolation_ensi, src/api/oi_ensi.cpp:33-568.
//
// One warp per background point, lane e <-> valid ensemble member e (E <= 32), at most 64 observations per point:
//   1. gather / top-k exactly as the deterministic OI (oi_ensi.cpp:207-269; only pobs validity is tested, :232);
//   2. Pinv = Y' Rinv Y + (E-1) I, E x E symmetric (oi_ensi.cpp:296-385), one row per lane, fp64;
//   3. ONE symmetric eigen-decomposition Pinv = V L V' (parallel-order cyclic Jacobi in shared memory) replaces the
//      reference's inv() + eig_sym((E-1) P) pair (oi_ensi.cpp:398-401): P = V L^-1 V', sqrt((E-1) P) = V sqrt((E-1)/L) V';
//   4. w = P C (obs - yhat), W = sqrtm + w 1' (oi_ensi.cpp:419-444), analysis_e = mean + sum_k X_k W_ke with the
//      reference's FLOAT accumulation over k (oi_ensi.cpp:506-512), optional clamp (:517-551, including the linear
//      index lY[e] of :523-524), written to member validEns[e].
// The reference runs this loop serially (its omp pragma is disabled, oi_ensi.cpp:203-206).
#include "oi.cuh"

#include <omp.h>

#include <algorithm>
#include <cstring>

using namespace gpp;

namespace {

constexpr int ENSI_WARPS = 2;
constexpr int ENSI_KMAX = 64;     // observations per point
constexpr int ENSI_EMAX = 32;     // valid ensemble members
constexpr int ENSI_GRAB = 32;     // consecutive points a warp takes per grab of the work counter
constexpr int ENSI_NSLOT = 3;     // candidate buffer = 96 entries
constexpr int ENSI_CHUNKS = 8;    // blocks of the field returned to the host while the next ones are analysed
// Jacobi: converged when off(A)^2 <= ENSI_CONV * diag(A)^2; a rotation is skipped when apq^2 <= ENSI_SKIP * |app aqq|.
// 1e-18: off-diagonal rms below 1e-9 of the diagonal, four orders under the 1e-5 parity bar even for a condition number of
// 1e3 (the smallest eigenvalue of Pinv is E - 1). Measured on C5 (profiles/r2_ensi_conv.log): 1e-26 496 ms, 1e-22 472 ms,
// 1e-18 450 ms, 1e-16 428 ms; the fields differ from the 1e-26 build by at most 2.4e-7 / 4.8e-7 / 7.2e-7 relative (the member sum
// is accumulated in float, oi_ensi.cpp:506-512, so a last-bit change of a term shows as an ulp or two of the result).
#ifndef ENSI_CONV
#define ENSI_CONV 1e-18
#endif
#ifndef ENSI_SKIP
#define ENSI_SKIP 1e-22
#endif

struct EnsiParams {
    const float *gx, *gy, *gz, *gelev, *glaf;
    const float* background;      // nB x nE, member fastest
    float* analysis;              // nB x nE, pre-filled with the background
    int first, count, nE, E;
    int valid_ens[ENSI_EMAX];
    ObsView obs;                  // ratio := psigma, innov := (double) pobs - (double) yhat
    const float* gY;              // [table slot][E]: pbackground minus its ensemble mean, valid members only
                                  // (utem: the standardised perturbations of pbackground_corr, oi_ensi_multi.cpp:978-993)
    const float* gY_raw;          // utem: [table slot][E] pbackground minus its mean (:968-976), for the clamp's lY[e]
    const float* background_corr; // utem: nB x nE
    const float* bratios;         // utem: nB
    gpp_structure s;
    float R;
    int k;
    int allow_extrapolation;
    int* num_skipped;
    int* work_counter;            // [0] next block of ENSI_GRAB points to hand out, [1] warps that ran out of work; both zero
                                  // between launches (the last warp of a launch resets them)
    int ld;                       // leading dimension of the shared matrices: E rounded up to odd
    int smem_per_warp;            // bytes, see EnsiSmem::layout
    int off[20];                  // byte offsets of the per-warp arrays (EnsiSmem::layout), read from constant memory
};

// Per-warp working set, carved from dynamic shared memory for the actual ensemble size and observation cap (a
// fixed 32 x 64 layout costs 36 KB per warp = 4 warps per SM; E = 20, k = 50 needs 19 KB). The host lays the arrays
// out once (byte offsets in EnsiParams::off); the kernel only adds them to its warp's base.
struct EnsiSmem {
    unsigned long long* key;      // [32 * ENSI_NSLOT] candidate keys
    float* Y;                     // lY, k x E (leading dimension ld), the float values the reference holds (gY is a vec2)
    double* A;                    // Pinv, then its diagonalisation, E x E (aliases Y)
    double* T;                    // Pinv V of the warm start, E x E (aliases Y, behind A)
    double* V;                    // eigenvectors in columns; kept from one point to the next (warm start)
    double *rinv, *dd;            // [k]
    double *b, *t, *lam, *w, *sc, *X;   // [E]
    double2* cs;                  // [E / 2 + 1] (cos, sin) of the rotations of a round
    int* pos;                     // [32 * ENSI_NSLOT]
    int* spos;                    // [k]
    int* pp;                      // [E / 2 + 1]
    float* sval;                  // [E]

    enum { O_KEY, O_Y, O_V, O_RINV, O_DD, O_B, O_T, O_LAM, O_W, O_SC, O_X, O_CS, O_POS, O_SPOS, O_PP, O_SVAL, O_COUNT };

    // fills off[O_COUNT] and returns the bytes per warp
    static size_t layout(int* off, int E, int kcap, int ld) {
        size_t at = 0;
        auto take = [&](int which, size_t bytes) { off[which] = (int) at; at += (bytes + 15) / 16 * 16; };
        const int Ee = E + (E & 1), h = Ee / 2 + 1;
        take(O_KEY, sizeof(unsigned long long) * 32 * ENSI_NSLOT);
        // A and T reuse lY's storage: lY is consumed (Pinv, C d, the clamp's lY[e]) before they are written
        take(O_Y, std::max(sizeof(float) * (size_t) kcap * ld, sizeof(double) * (size_t) 2 * Ee * ld));
        take(O_V, sizeof(double) * Ee * ld);
        take(O_RINV, sizeof(double) * kcap);
        take(O_DD, sizeof(double) * kcap);
        const int per_member[6] = {O_B, O_T, O_LAM, O_W, O_SC, O_X};
        for(int i = 0; i < 6; i++) take(per_member[i], sizeof(double) * Ee);
        take(O_CS, sizeof(double2) * h);
        take(O_POS, sizeof(int) * 32 * ENSI_NSLOT);
        take(O_SPOS, sizeof(int) * kcap);
        take(O_PP, sizeof(int) * h);
        take(O_SVAL, sizeof(float) * Ee);
        return at;
    }
    __device__ __forceinline__ void bind(unsigned char* base, const int* off, int E, int ld) {
        const int Ee = E + (E & 1);
        key = reinterpret_cast<unsigned long long*>(base + off[O_KEY]);
        Y = reinterpret_cast<float*>(base + off[O_Y]);
        A = reinterpret_cast<double*>(base + off[O_Y]);
        T = A + Ee * ld;
        V = reinterpret_cast<double*>(base + off[O_V]);
        rinv = reinterpret_cast<double*>(base + off[O_RINV]);
        dd = reinterpret_cast<double*>(base + off[O_DD]);
        b = reinterpret_cast<double*>(base + off[O_B]);
        t = reinterpret_cast<double*>(base + off[O_T]);
        lam = reinterpret_cast<double*>(base + off[O_LAM]);
        w = reinterpret_cast<double*>(base + off[O_W]);
        sc = reinterpret_cast<double*>(base + off[O_SC]);
        X = reinterpret_cast<double*>(base + off[O_X]);
        cs = reinterpret_cast<double2*>(base + off[O_CS]);
        pos = reinterpret_cast<int*>(base + off[O_POS]);
        spos = reinterpret_cast<int*>(base + off[O_SPOS]);
        pp = reinterpret_cast<int*>(base + off[O_PP]);
        sval = reinterpret_cast<float*>(base + off[O_SVAL]);
    }
};
static_assert(EnsiSmem::O_COUNT <= 20, "EnsiParams::off is too small");

#ifdef ENSI_STATS
// debug build only (profiles/variants.sh -DENSI_STATS): [0] points, [1] sweeps, [2] rotations, [3] warm starts
__device__ unsigned long long g_ensi_stats[4];
#define ENSI_COUNT(i, n) do { if(lane_id() == 0) atomicAdd(&g_ensi_stats[i], (unsigned long long) (n)); } while(0)
#else
#define ENSI_COUNT(i, n) do {} while(0)
#endif

// members with an invalid value anywhere in the background are left untouched (oi_ensi.cpp:187-201)
__global__ void ensi_invalid_members_kernel(const float* __restrict__ background, size_t n, int nE, int* __restrict__ invalid) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(!is_valid(background[i])) atomicOr(&invalid[i % nE], 1);
}

#ifndef ENSI_MINB
#define ENSI_MINB 10
#endif
#ifdef ENSI_REGISTER_JACOBI
// EXPERIMENTAL, NOT COMPILED BY DEFAULT (profiles/variants.sh regj "-DENSI_REGISTER_JACOBI"; unverified on a GPU so
// far). The Jacobi iteration with row `lane` of A and of V in registers, for an even compile-time E <= 20: the M = E - 1
// rounds of the round-robin schedule are unrolled, so the columns a round rotates are static register indices; the two
// lanes of a pair exchange their rows with shuffles. Index formulas, phase order and signs are those checked lane by lane
// against numpy in profiles/jacobi_systolic_sim.py:
//   partner(r, i) = r if i == M; M if i == r; else (2 r - i) mod M       slot = 0 | min(l, M - l), l = (i - r) mod M
//   pair of slot t: t == 0 -> (r, M), else ((r + t) mod M, (r - t) mod M), smaller index first
// Shared memory only carries what needs a lane-dependent index: the off-diagonal element of each pair (written by the
// pair's first lane with a static predicated store) and the (cos, sin) of the round's rotations.
template <int E>
struct RegisterJacobi {
    static constexpr int M = E - 1, H = E / 2;
    // one round; R is the compile-time round number
    template <int R>
    static __device__ __forceinline__ void round(double (&a)[E], double (&v)[E], double& dgl, double2* cs, double* apq_s, int lane) {
        // ---- this lane's pair
        int partner = lane, slot = H;
        if(lane < E) {
            if(lane == M) { partner = R; slot = 0; }
            else if(lane == R) { partner = M; slot = 0; }
            else {
                int l = lane - R;
                if(l < 0) l += M;
                partner = (2 * R - lane + 2 * M) % M;
                slot = l < M - l ? l : M - l;
            }
        }
        const bool first = lane < partner;   // the pair's smaller index p; the other lane is q
        // ---- the pairs' off-diagonal elements: lane p holds a[p][q] at the static index q of slot t
        #pragma unroll
        for(int t = 0; t < H; t++) {
            constexpr int dummy = 0; (void) dummy;
            const int p0 = t == 0 ? R : (R + t) % M, q0 = t == 0 ? M : (R - t + M) % M;
            const int pt = p0 < q0 ? p0 : q0, qt = p0 < q0 ? q0 : p0;
            if(lane == pt) apq_s[t] = a[qt];
        }
        __syncwarp();
        const double other = shfl_double(dgl, partner);
        double c = 1.0, s = 0.0, tq = 0.0;
        if(lane < E) {
            const double apq = apq_s[slot];
            const double app = first ? dgl : other, aqq = first ? other : dgl;
            if(apq * apq > ENSI_SKIP * fabs(app * aqq)) {
                const double theta = (aqq - app) * (0.5 * fast_rcp(apq));
                const double at = fabs(theta);
                double tt;
                if(at > 1e100) tt = 0.5 * fast_rcp(theta);
                else tt = copysign(fast_rcp(at + (at * at + 1.0) * fast_rsqrt(at * at + 1.0)), theta);
                c = fast_rsqrt(tt * tt + 1.0);
                s = tt * c;
                tq = tt * apq;
            }
            if(first) cs[slot] = make_double2(c, s);
        }
        __syncwarp();
        // ---- columns p_t, q_t of A and V (static indices), every lane its own two elements
        #pragma unroll
        for(int t = 0; t < H; t++) {
            const int p0 = t == 0 ? R : (R + t) % M, q0 = t == 0 ? M : (R - t + M) % M;
            const int pt = p0 < q0 ? p0 : q0, qt = p0 < q0 ? q0 : p0;
            const double2 r2 = cs[t];
            const double ap = a[pt], aq = a[qt], vp = v[pt], vq = v[qt];
            a[pt] = r2.x * ap - r2.y * aq;
            a[qt] = r2.y * ap + r2.x * aq;
            v[pt] = r2.x * vp - r2.y * vq;
            v[qt] = r2.y * vp + r2.x * vq;
        }
        // ---- rows p, q of A: new_p = c row_p - s row_q, new_q = s row_p + c row_q (the partner's row by shuffle)
        const double ss = first ? -s : s;
        #pragma unroll
        for(int j = 0; j < E; j++) {
            const double o = shfl_double(a[j], partner);
            a[j] = c * a[j] + ss * o;
        }
        dgl = first ? dgl - tq : dgl + tq;   // a_pp - t a_pq, a_qq + t a_pq
        __syncwarp();                         // cs / apq_s are rewritten by the next round
    }
    template <int R>
    static __device__ __forceinline__ void rounds(double (&a)[E], double (&v)[E], double& dgl, double2* cs, double* apq_s, int lane) {
        if constexpr(R < M) {
            round<R>(a, v, dgl, cs, apq_s, lane);
            rounds<R + 1>(a, v, dgl, cs, apq_s, lane);
        }
    }
    // A (rows in S_A) is diagonalised, V (rows in S_V) accumulates the rotations; returns false when A is not finite
    static __device__ __forceinline__ bool run(double* S_A, double* S_V, int LD, double2* cs, double* apq_s, int lane) {
        double a[E], v[E];
        #pragma unroll
        for(int f = 0; f < E; f++) {
            a[f] = lane < E ? S_A[lane * LD + f] : 0.0;
            v[f] = lane < E ? S_V[lane * LD + f] : 0.0;
        }
        bool ok = true;
        for(int sweep = 0; sweep < 30; sweep++) {
            double off = 0.0, dg = 0.0, dgl = 0.0;
            #pragma unroll
            for(int f = 0; f < E; f++) {
                const double sq = a[f] * a[f];
                if(f == lane) { dg += sq; dgl = a[f]; }
                else off += sq;
            }
            #pragma unroll
            for(int o = 16; o > 0; o >>= 1) { off += shfl_double(off, lane ^ o); dg += shfl_double(dg, lane ^ o); }
            if(!(off == off) || !(dg == dg) || isinf(off) || isinf(dg)) { ok = false; break; }
            if(off <= ENSI_CONV * dg) break;
            rounds<0>(a, v, dgl, cs, apq_s, lane);
        }
        if(lane < E) {
            #pragma unroll
            for(int f = 0; f < E; f++) {
                S_V[lane * LD + f] = v[f];
                if(f == lane) S_A[lane * LD + f] = a[f];
            }
        }
        __syncwarp();
        return ok;
    }
};
#endif

// EC > 0: the number of valid members is the compile-time constant EC (loops over the members unroll without guards, the
// strides of the shared matrices are immediates); EC == 0: any number up to ENSI_EMAX.
#ifdef ENSI_REGISTER_JACOBI
#define ENSI_MINB_FOR(EC) (((EC) > 0 && (EC) <= 20 && (EC) % 2 == 0) ? 7 : ENSI_MINB)
#else
#define ENSI_MINB_FOR(EC) ENSI_MINB
#endif
// UTEM: gridpp::optimal_interpolation_ensi_multi_utem (oi_ensi_multi.cpp:862-1311) -- the same transform computed from the
// standardised *_corr ensembles (Pinv = C Y_corr + I, Rinv = rho / pratios) and applied to the standardised perturbations
// of background_corr, scaled by the spread of background and by bratios.
// PAD: the number of valid members is P.E <= EC -- the member loops unroll to EC with guards, which keeps the per-lane arrays at
// EC entries instead of ENSI_EMAX (the fully generic EC = 0 form needs 64 + 64 registers for them and spills: 2.8x slower at E = 20)
template <int SMODE, int EC, bool UTEM = false, bool PAD = false>
__global__ void __launch_bounds__(ENSI_WARPS * 32, ENSI_MINB_FOR(EC)) ensi_kernel(const __grid_constant__ EnsiParams P) {
    constexpr bool EXACT = EC > 0 && !PAD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EnsiSmem S;
    S.bind(smem_raw + (size_t) (threadIdx.x >> 5) * P.smem_per_warp, P.off, P.E, P.ld);
    const int LD = EXACT ? (EC | 1) : P.ld;
    const int lane = (int) lane_id();
    const CandBuf cb = {S.key, S.pos};
    const int E = EXACT ? EC : P.E;
    constexpr int EU = EC > 0 ? EC : ENSI_EMAX;   // trip count of the unrolled member loops
    const int Eeven = E + (E & 1), m = Eeven - 1;   // round-robin schedule over an even number of indices
    const int H = Eeven / 2;                        // rotations per round
    __shared__ unsigned char tri[ENSI_EMAX / 2 * (ENSI_EMAX / 2 + 1) / 2];   // pair index -> (t, u >= t) of two rotations of a round
    for(int u = 0, mm = 0; u < ENSI_EMAX / 2; u++)
        for(int t = 0; t <= u; t++, mm++)
            if((int) threadIdx.x == (mm & 63)) tri[mm] = (unsigned char) (t | (u << 4));
    __syncthreads();

    // Blocks of ENSI_GRAB consecutive points are handed out dynamically (the cost per point varies with k and the
    // sweep count). Inside a block every point starts its Jacobi iteration from the eigenvectors of the point before
    // it (neighbouring points have nearly the same Pinv). The blocks are aligned to ABSOLUTE point indices, and a
    // range that begins inside a block replays the block's earlier points without writing them, so that a point's
    // result does not depend on how the field was cut into ranges.
    const int blk0 = P.first / ENSI_GRAB;
    const int n_blk = (P.first + P.count + ENSI_GRAB - 1) / ENSI_GRAB - blk0;
    for(;;) {
    int blk = 0;
    if(lane == 0) blk = atomicAdd(P.work_counter, 1);
    blk = __shfl_sync(0xffffffffu, blk, 0);
    if(blk >= n_blk) break;
    const int g_begin = (blk0 + blk) * ENSI_GRAB, g_end = min(g_begin + ENSI_GRAB, P.first + P.count);
    bool warm = false;   // S.V holds the eigenvectors of an earlier point of this block
    for(int g = g_begin; g < g_end; g++) {
        const bool emit = g >= P.first;
        const Pt p1 = {P.gx[g], P.gy[g], P.gz[g], P.gelev[g], P.glaf[g]};
        bool cut = false;
        const int k = gather_candidates<SMODE, ENSI_NSLOT>(P.obs, P.s, p1, P.R, P.k, cb, &cut);
        if(k == 0) continue;   // oi_ensi.cpp:210-214,265-269: too few observations, keep the background
        // ---- order of the local observations: best-first when the selection was cut to max_points
        // (oi_ensi.cpp:244-255; the prune leaves exactly that order), otherwise the order of the radius query =
        // ascending index (oi_ensi.cpp:256-263). The order matters for the linear index lY[e] of the clamp.
        if(!cut) {
            const unsigned long long k0 = lane < k ? S.key[lane] : 0ull, k1 = lane + 32 < k ? S.key[lane + 32] : 0ull;
            const int p0 = lane < k ? S.pos[lane] : 0, p1s = lane + 32 < k ? S.pos[lane + 32] : 0;
            // ascending original index = descending low word of the key
            const unsigned* lo_words = reinterpret_cast<const unsigned*>(S.key);
            int r0 = 0, r1 = 0;
            for(int j = 0; j < k; j++) {
                const unsigned lw = lo_words[2 * j];
                r0 += lw > (unsigned) k0;
                r1 += lw > (unsigned) k1;
            }
            __syncwarp();
            if(lane < k) { S.key[r0] = k0; S.pos[r0] = p0; }
            if(lane + 32 < k) { S.key[r1] = k1; S.pos[r1] = p1s; }
            __syncwarp();
        }
        // ---- stage lY (oi_ensi.cpp:282-294), Rinv (:296-302), obs - yhat (:434-437)
        for(int i = lane; i < k; i += 32) {
            const int pos = S.pos[i];
            S.spos[i] = pos;
            const float sigma = P.obs.ratio[pos];
            S.rinv[i] = (double) cand_key_rho(S.key[i]) / (double) (UTEM ? sigma : __fmul_rn(sigma, sigma));   // oi_ensi_multi.cpp:1094
            S.dd[i] = P.obs.innov[pos];
        }
        __syncwarp();
        for(int idx = lane; idx < k * E; idx += 32) {
            const int i = idx / E, e = idx - i * E;
            S.Y[i * LD + e] = P.gY[(size_t) S.spos[i] * E + e];
        }
        __syncwarp();
        // ---- Pinv = C * lY + diag * I, C = lY' Rinv (oi_ensi.cpp:379-385), row `lane`; b = C (obs - yhat) (:428-437)
        double lYe = 0.0;   // the clamp's lY[e]: a LINEAR index into the column-major k x E matrix (oi_ensi.cpp:523-524)
        {
            double acc[EU];
            #pragma unroll
            for(int f = 0; f < EU; f++) acc[f] = 0.0;
            double b = 0.0;
            if(lane < E) {
                for(int i = 0; i < k; i++) {
                    const double c = (double) S.Y[i * LD + lane] * S.rinv[i];
                    b = fma(c, S.dd[i], b);
                    #pragma unroll
                    for(int f = 0; f < EU; f++)
                        if(EXACT || f < E) acc[f] = fma(c, (double) S.Y[i * LD + f], acc[f]);
                }
                S.b[lane] = b;
                lYe = UTEM ? (double) P.gY_raw[(size_t) S.spos[lane % k] * E + lane / k] : (double) S.Y[(lane % k) * LD + (lane / k)];
            }
            __syncwarp();   // lY is dead from here on: A and T take its place
            const double diag = UTEM ? 1.0 : (double) (float) (E - 1);   // oi_ensi.cpp:383 with delta = 1; oi_ensi_multi.cpp:1105
            #pragma unroll
            for(int f = 0; f < EU; f++)
                if(f == lane) acc[f] += diag;
            ENSI_COUNT(0, 1);
            ENSI_COUNT(3, warm ? 1 : 0);
#ifdef ENSI_COLD
            warm = false;
#endif
            if(!warm) {
                #pragma unroll
                for(int f = 0; f < EU; f++)
                    if((EXACT || f < E) && lane < E) {
                        S.A[lane * LD + f] = acc[f];
                        S.V[lane * LD + f] = f == lane ? 1.0 : 0.0;
                    }
            }
            else {
                // warm start: A = V' Pinv V with the previous point's V, which leaves only small off-diagonal elements
                if(lane < E)
                    for(int j = 0; j < E; j++) {   // T = Pinv V, row `lane`
                        double t = 0.0;
                        #pragma unroll
                        for(int f = 0; f < EU; f++)
                            if(EXACT || f < E) t = fma(acc[f], S.V[f * LD + j], t);
                        S.T[lane * LD + j] = t;
                    }
                __syncwarp();
                if(lane < E) {
                    #pragma unroll
                    for(int f = 0; f < EU; f++)
                        if(EXACT || f < E) acc[f] = S.V[f * LD + lane];   // column `lane` of V
                    for(int j = 0; j < E; j++) {   // A = V' T, row `lane`
                        double t = 0.0;
                        #pragma unroll
                        for(int f = 0; f < EU; f++)
                            if(EXACT || f < E) t = fma(acc[f], S.T[f * LD + j], t);
                        S.A[lane * LD + j] = t;
                    }
                }
            }
        }
        __syncwarp();
        // ---- cyclic Jacobi, parallel (round-robin) ordering: E/2 disjoint rotations per round
        bool bad = false;
#ifdef ENSI_REGISTER_JACOBI
        constexpr bool REGJ = EXACT && EC <= 20 && EC % 2 == 0;
#else
        constexpr bool REGJ = false;
#endif
        if constexpr(REGJ) {
#ifdef ENSI_REGISTER_JACOBI
            bad = !RegisterJacobi<REGJ ? EC : 2>::run(S.A, S.V, LD, S.cs, S.t, lane);   // S.t: [E] doubles, free until after the eigenvalues
#endif
        }
        else
        for(int sweep = 0; sweep < 30; sweep++) {
            double off = 0.0, dg = 0.0;
            if(lane < E)
                for(int f = 0; f < E; f++) {
                    const double v = S.A[lane * LD + f];
                    if(f == lane) dg += v * v; else off += v * v;
                }
            #pragma unroll
            for(int o = 16; o > 0; o >>= 1) { off += shfl_double(off, lane ^ o); dg += shfl_double(dg, lane ^ o); }
            if(!(off == off) || !(dg == dg) || isinf(off) || isinf(dg)) { bad = true; break; }
            if(off <= ENSI_CONV * dg) break;
            ENSI_COUNT(1, 1);
            // round-robin schedule: in round r lane l > 0 pairs (r + l) % m with (r - l) % m, lane 0 pairs m with r
            int pr = lane % m, qr = (m - lane % m) % m;
            for(int r = 0; r < m; r++) {
                bool rot = false;
                int p = 0, q = 0;
                double c = 1.0, s = 0.0;
                if(lane < H) {
                    if(lane == 0) { p = m; q = r; }
                    else { p = pr; q = qr; }
                    if(p > q) { int tmp = p; p = q; q = tmp; }
                    if(q < E) {   // (a pair with the padding index of an odd E is the identity)
                        const double apq = S.A[p * LD + q], app = S.A[p * LD + p], aqq = S.A[q * LD + q];
                     
                        if(apq * apq > ENSI_SKIP * fabs(app * aqq)) {
                            // t = sign(theta) / (|theta| + sqrt(theta^2 + 1)), theta = (aqq - app) / (2 apq); c = 1 / sqrt(t^2 + 1)
                            const double theta = (aqq - app) * (0.5 * fast_rcp(apq));
                            const double at = fabs(theta);
                            double tt;
                            if(at > 1e100) tt = 0.5 * fast_rcp(theta);                  // theta^2 would overflow
                            else tt = copysign(fast_rcp(at + (at * at + 1.0) * fast_rsqrt(at * at + 1.0)), theta);
                            c = fast_rsqrt(tt * tt + 1.0);
                            s = tt * c;
                            rot = true;
                        }
                    }
                }
                pr = pr + 1 == m ? 0 : pr + 1;
                qr = qr + 1 == m ? 0 : qr + 1;
                // the rotations of this round, compacted
                const unsigned rmask = __ballot_sync(0xffffffffu, rot);
                const int nrot = __popc(rmask);
                if(nrot == 0) continue;
                ENSI_COUNT(2, nrot);
#ifndef ENSI_NO_BLOCKS
                const bool BLOCKS = (E & 1) == 0;   // (an odd E pairs one index with a padding index that has no row)
#else
                constexpr bool BLOCKS = false;
#endif
#ifndef ENSI_BLOCK_FRAC
#define ENSI_BLOCK_FRAC 2
#endif
                const bool by_blocks = BLOCKS && ENSI_BLOCK_FRAC * nrot > H;   // skipped rotations take part as identities (c = 1, s = 0)
                if(by_blocks) {
                    if(lane < H) { S.pp[lane] = p | (q << 8); S.cs[lane] = make_double2(c, s); }
                }
                else if(rot) {
                    const int slot = __popc(rmask & ((1u << lane) - 1u));
                    S.pp[slot] = p | (q << 8); S.cs[slot] = make_double2(c, s);
                }
                __syncwarp();
#ifndef ENSI_NO_BLOCKS
                if(by_blocks) {
                    // A round in which most index pairs are rotated: A <- J' A J one 2 x 2 block at a time. Block (t, u) of the rotations'
                    // index pairs gets the left transform of rotation t and the right transform of rotation u in registers -- the
                    // same operations in the same order as the row pass followed by the column pass below -- and is written once,
                    // together with its mirror image (A stays symmetric): 4 loads and 8 stores per block for the H (H + 1) / 2
                    // blocks with t <= u, against two loads and two stores per element in each of the two passes.
                    for(int mm = lane; mm < H * (H + 1) / 2; mm += 32) {
                        const int tu = tri[mm], t = tu & 15, u = tu >> 4;
                        const int pq = S.pp[t], PQ = S.pp[u];
                        const int p0 = pq & 255, q0 = pq >> 8, p1 = PQ & 255, q1 = PQ >> 8;
                        const double2 rt = S.cs[t], ru = S.cs[u];
                        const double a00 = S.A[p0 * LD + p1], a01 = S.A[p0 * LD + q1], a10 = S.A[q0 * LD + p1], a11 = S.A[q0 * LD + q1];
                        const double b00 = rt.x * a00 - rt.y * a10, b10 = rt.y * a00 + rt.x * a10;
                        const double b01 = rt.x * a01 - rt.y * a11, b11 = rt.y * a01 + rt.x * a11;
                        const double r00 = ru.x * b00 - ru.y * b01, r01 = ru.y * b00 + ru.x * b01;
                        const double r10 = ru.x * b10 - ru.y * b11, r11 = ru.y * b10 + ru.x * b11;
                        S.A[p0 * LD + p1] = r00; S.A[p0 * LD + q1] = r01; S.A[q0 * LD + p1] = r10; S.A[q0 * LD + q1] = r11;
                        if(t != u) { S.A[p1 * LD + p0] = r00; S.A[q1 * LD + p0] = r01; S.A[p1 * LD + q0] = r10; S.A[q1 * LD + q0] = r11; }
                    }
                    if(lane < E) {   // V <- V J: columns p, q, row `lane`
                        double* vrow = S.V + lane * LD;
                        for(int t = 0; t < H; t++) {
                            const int pq = S.pp[t], p = pq & 255, q = pq >> 8;
                            const double2 rot2 = S.cs[t];
                            const double vp = vrow[p], vq = vrow[q];
                            vrow[p] = rot2.x * vp - rot2.y * vq;
                            vrow[q] = rot2.y * vp + rot2.x * vq;
                        }
                    }
                    __syncwarp();
                    continue;
                }
#endif
                if(lane < E)   // A <- J' A: rows p, q of A, column `lane`
                    for(int t = 0; t < nrot; t++) {
                        const int pq = S.pp[t];
                        double* rp = S.A + (pq & 255) * LD + lane;
                        double* rq = S.A + (pq >> 8) * LD + lane;
                        const double2 rot2 = S.cs[t];
                        const double c = rot2.x, s = rot2.y;
                        const double ap = *rp, aq = *rq;
                        *rp = c * ap - s * aq;
                        *rq = s * ap + c * aq;
                    }
                __syncwarp();
                if(lane < E) {   // A <- A J, V <- V J: columns p, q, row `lane`
                    double* arow = S.A + lane * LD;
                    double* vrow = S.V + lane * LD;
                    for(int t = 0; t < nrot; t++) {
                        const int pq = S.pp[t], p = pq & 255, q = pq >> 8;
                        const double2 rot2 = S.cs[t];
                        const double c = rot2.x, s = rot2.y;
                        const double ap = arow[p], aq = arow[q];
                        arow[p] = c * ap - s * aq;
                        arow[q] = s * ap + c * aq;
                        const double vp = vrow[p], vq = vrow[q];
                        vrow[p] = c * vp - s * vq;
                        vrow[q] = s * vp + c * vq;
                    }
                }
                __syncwarp();
            }
        }
        // eigenvalues of Pinv; rcond(Pinv) <= 0 in the reference (oi_ensi.cpp:386-390) <=> not positive definite / not finite
        double lam = lane < E ? S.A[lane * LD + lane] : 1.0;
        bad = bad || __any_sync(0xffffffffu, !(lam > 0.0) || isinf(lam));
        warm = !bad;
        if(bad) {
            if(lane == 0 && emit && P.num_skipped) atomicAdd(P.num_skipped, 1);
            continue;
        }
        if(!emit) continue;   // replayed for the warm-start chain only
        if(lane < E) {
            S.lam[lane] = lam;
            S.sc[lane] = sqrt((double) (E - 1) / lam);   // sqrt of the eigenvalues of (E-1) P, oi_ensi.cpp:401,419
            // X: background perturbations about the ensemble mean (oi_ensi.cpp:447-462), float mean
            S.sval[lane] = P.background[(size_t) g * P.nE + P.valid_ens[lane]];
        }
        __syncwarp();
        float total = 0.f;
        for(int e = 0; e < E; e++) total = __fadd_rn(total, S.sval[e]);
        const float ensMean = __fdiv_rn(total, (float) E);
        if(lane < E) {
            S.X[lane] = (double) S.sval[lane] - (double) ensMean;
            double t = 0.0;   // t = V' b
            for(int e = 0; e < E; e++) t = fma(S.V[e * LD + lane], S.b[e], t);
            S.t[lane] = t / lam;
        }
        __syncwarp();
        if(lane < E) {
            double w = 0.0;   // w = P C (obs - yhat) = V L^-1 V' b  (oi_ensi.cpp:428-437)
            for(int f = 0; f < E; f++) w = fma(S.V[lane * LD + f], S.t[f], w);
            S.w[lane] = w;
        }
        __syncwarp();
        float ensStd = 1.f, bratio = 1.f;
        if(UTEM) {
            // oi_ensi_multi.cpp:1160-1198: X_corr, the standardised perturbations of background_corr (into S.t, which is
            // dead now, as is S.b), and the spread of background
            float* cval = reinterpret_cast<float*>(S.b);
            if(lane < E) cval[lane] = P.background_corr[(size_t) g * P.nE + P.valid_ens[lane]];
            __syncwarp();
            float mean_c, std_c, unused;
            seq_mean_std(cval, E, &mean_c, &std_c);
            seq_mean_std(S.sval, E, &unused, &ensStd);
            const float const_fact = (float) (1.0 / sqrt((double) (E - 1)));
            if(lane < E)
                S.t[lane] = std_c <= 0.0013f ? 0.0 : (double) __fdiv_rn(__fmul_rn(const_fact, __fsub_rn(cval[lane], mean_c)), std_c);
            bratio = P.bratios[g];
            __syncwarp();
        }
        if(lane < E) {
            // analysis for member `lane`: total += X(k) * W(k, e) accumulated in FLOAT (oi_ensi.cpp:506-512),
            // W(k, e) = sum_f V(k,f) sqrt((E-1)/lam_f) V(e,f) + w(k)  (oi_ensi.cpp:419-444)
            float tot = 0.f;
            double vs[EU];   // row `lane` of V scaled by sqrt((E-1) / lambda)
            #pragma unroll
            for(int f = 0; f < EU; f++) vs[f] = (EXACT || f < E) ? S.V[lane * LD + f] * S.sc[f] : 0.0;
            for(int kk = 0; kk < E; kk++) {
                double wke = 0.0;
                #pragma unroll
                for(int f = 0; f < EU; f++)
                    if(EXACT || f < E) wke = fma(S.V[kk * LD + f], vs[f], wke);
                if(UTEM) {   // W(e, e2) = ensStd * W(e, e2) + std_ratios_lr * w(e), applied to X_corr (oi_ensi_multi.cpp:1201-1254)
                    wke = __dadd_rn(__dmul_rn((double) ensStd, wke), __dmul_rn((double) bratio, S.w[kk]));
                    tot = (float) __dadd_rn((double) tot, __dmul_rn(S.t[kk], wke));
                }
                else {
                    wke += S.w[kk];
                    tot = (float) ((double) tot + S.X[kk] * wke);
                }
            }
            float currIncrement = tot;
            if(!P.allow_extrapolation) {   // oi_ensi.cpp:517-551
                // lY[e] is a LINEAR index into the column-major k x E matrix: row e % k, column e / k (:523-524)
                double mx = -INFINITY, mn = INFINITY;
                for(int i = 0; i < k; i++) {
                    const double v = S.dd[i] - lYe;
                    mx = fmax(mx, v);
                    mn = fmin(mn, v);
                }
                const float maxInc = (float) mx, minInc = (float) mn;
                const float memberIncrement = (float) ((double) currIncrement - S.X[lane]);
                if(maxInc > 0 && memberIncrement > maxInc) currIncrement = (float) ((double) maxInc + S.X[lane]);
                else if(maxInc < 0 && memberIncrement > 0) currIncrement = (float) (0.0 + S.X[lane]);
                else if(minInc < 0 && memberIncrement < minInc) currIncrement = (float) ((double) minInc + S.X[lane]);
                else if(minInc > 0 && memberIncrement < 0) currIncrement = (float) (0.0 + S.X[lane]);
            }
            P.analysis[(size_t) g * P.nE + P.valid_ens[lane]] = __fadd_rn(ensMean, currIncrement);
        }
        __syncwarp();
    }
    }
    if(lane == 0) {   // the last warp to run out of work leaves the counters ready for the next launch
        __threadfence();
        if(atomicAdd(P.work_counter + 1, 1) == (int) (gridDim.x * ENSI_WARPS) - 1) {
            P.work_counter[0] = 0;
            P.work_counter[1] = 0;
            __threadfence();
        }
    }
}

// util.cpp:19-43, Mean branch of calc_statistic: float accumulation over the valid values
float mean_valid(const float* a, int n) {
    float total = 0;
    int count = 0;
    for(int i = 0; i < n; i++)
        if(is_valid(a[i])) { total += a[i]; count++; }
    return count > 0 ? total / count : NAN;
}

constexpr unsigned ENSI_COUNTER_SLOTS = 16;
}  // namespace

// Observation side of an EnSI call, resident on the device: the bucket grid over the observations with a valid value
// (oi_ensi.cpp:232), sigma and obs - yhat per observation, and the perturbations of the background at the observation
// points about their ensemble mean (oi_ensi.cpp:163-178) for the valid members.
struct gpp_ensi_obs {
    gpp_oi_obs table;
    gpp::DeviceBuffer<float> gY;          // [table slot][E]
    gpp::DeviceBuffer<float> gY_raw;      // utem only: [table slot][E]
    gpp::DeviceBuffer<int> counters;      // ENSI_COUNTER_SLOTS x {next block, warps done}, zero between launches
    mutable std::atomic<unsigned> next_slot{0};
    int nE = 0, E = 0;
    int valid_ens[ENSI_EMAX];
};

namespace {
int build_ensi_obs(gpp_ensi_obs& st, const gpp_points* opoints, const float* pobs, const float* psigmas, const float* pbackground, int nE,
                   const int* member_valid, const gpp_structure* structure) {
    const int nS = opoints->n;
    st.nE = nE;
    st.E = 0;
    for(int e = 0; e < nE; e++)
        if(!member_valid || member_valid[e]) {
            if(st.E >= ENSI_EMAX)
                return fail(GPP_ERR_NOT_IMPLEMENTED, "optimal_interpolation_ensi supports at most %d valid ensemble members on the device", ENSI_EMAX);
            st.valid_ens[st.E++] = e;
        }
    // ---- oi_ensi.cpp:163-178: remove the ensemble mean at the observation points
    std::vector<float> gY(pbackground, pbackground + (size_t) nS * nE), gYhat(nS);
    for(int i = 0; i < nS; i++) {
        float mean = mean_valid(&gY[(size_t) i * nE], nE);
        for(int e = 0; e < nE; e++) {
            float value = gY[(size_t) i * nE + e];
            if(is_valid(value) && is_valid(mean)) gY[(size_t) i * nE + e] -= mean;
        }
        gYhat[i] = mean;
    }
    // ---- observation table: only pobs validity is required here (oi_ensi.cpp:232)
    std::vector<char> valid(nS);
    std::vector<double> innov(nS);
    std::vector<float> sig(psigmas, psigmas + nS);
    for(int i = 0; i < nS; i++) {
        valid[i] = is_valid(pobs[i]);
        innov[i] = (double) pobs[i] - (double) gYhat[i];   // lObs - lYhat, oi_ensi.cpp:437
    }
    std::vector<int> order;
    GPP_TRY(build_obs_table(opoints, valid, innov, sig, structure->term[0].loc_dist, &st.table, &order));
    const int E = st.E;
    std::vector<float> gYs(std::max<size_t>(1, order.size() * (size_t) std::max(E, 1)));
    for(size_t slot = 0; slot < order.size(); slot++)
        for(int e = 0; e < E; e++) gYs[slot * E + e] = gY[(size_t) order[slot] * nE + st.valid_ens[e]];
    GPP_TRY(st.gY.upload(gYs.data(), gYs.size()));
    GPP_TRY(st.counters.alloc(2 * ENSI_COUNTER_SLOTS));
    GPP_CUDA(cudaMemsetAsync(st.counters.ptr, 0, sizeof(int) * 2 * ENSI_COUNTER_SLOTS, 0));
    GPP_CUDA(cudaStreamSynchronize(0));   // the staging vectors go out of scope
    return GPP_OK;
}

// Largest number of observations per point the kernel must hold (0 = nothing to do); max_points == 0 needs a counting pass
int ensi_kcap(const gpp_ensi_obs& st, gpp_points* bp, int first, int count, const gpp_structure* structure, int max_points, cudaStream_t stream, int* out) {
    int kcap = max_points > 0 ? std::min(max_points, st.table.n_valid) : st.table.n_valid;
    if(max_points == 0 && kcap > ENSI_KMAX) {
        int hmax = 0;
        GPP_TRY(count_max_candidates(bp, first, count, nullptr, st.table.view(), structure->term[0].loc_dist, stream, &hmax));
        kcap = std::min(kcap, std::max(hmax, 1));
    }
    if(kcap > ENSI_KMAX)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "optimal_interpolation_ensi supports at most %d observations per point on the device (got %d)", ENSI_KMAX, kcap);
    *out = kcap;
    return GPP_OK;
}

// Analyses points [first, first + count): d_analysis must already hold the background there. One kernel launch.
int ensi_launch(const gpp_ensi_obs& st, gpp_points* bp, int first, int count, const float* d_background, float* d_analysis,
                const gpp_structure* structure, int kcap, int allow_extrapolation, int* d_num_skipped, int* counter_pair, cudaStream_t stream,
                const float* d_background_corr = nullptr, const float* d_bratios = nullptr /* both given: the utem variant */) {
    if(st.E == 0 || st.table.n_valid == 0 || count == 0) return GPP_OK;
    EnsiParams P;
    std::memset(&P, 0, sizeof(P));
    for(int e = 0; e < st.E; e++) P.valid_ens[e] = st.valid_ens[e];
    P.gx = bp->dx.ptr; P.gy = bp->dy.ptr; P.gz = bp->dz.ptr; P.gelev = bp->delev.ptr; P.glaf = bp->dlaf.ptr;
    P.background = d_background;
    P.analysis = d_analysis;
    P.first = first; P.count = count; P.nE = st.nE; P.E = st.E;
    P.obs = st.table.view();
    P.gY = st.gY.ptr;
    const bool utem = d_background_corr && d_bratios;
    P.gY_raw = st.gY_raw.ptr; P.background_corr = d_background_corr; P.bratios = d_bratios;
    P.s = *structure;
    P.R = structure->term[0].loc_dist;
    P.allow_extrapolation = allow_extrapolation;
    P.num_skipped = d_num_skipped;
    P.work_counter = counter_pair;
    P.k = kcap;
    const int E = st.E;
    P.ld = E | 1;
    P.smem_per_warp = (int) EnsiSmem::layout(P.off, E, kcap, P.ld);
    const size_t smem = (size_t) P.smem_per_warp * ENSI_WARPS;
    const int mode = structure_mode(*structure);
    void (*kernel)(EnsiParams) = nullptr;
    // the common ensemble sizes have their own instantiation; the others take the next padded size (8 / 16 / 24, else 32)
#define ENSI_PICK(EC, PAD) (utem ? (mode == 1 ? ensi_kernel<1, EC, true, PAD> : ensi_kernel<0, EC, true, PAD>) \
                                 : (mode == 1 ? ensi_kernel<1, EC, false, PAD> : ensi_kernel<0, EC, false, PAD>))
    switch(E) {
        case 10: kernel = ENSI_PICK(10, false); break;
        case 20: kernel = ENSI_PICK(20, false); break;
        case 30: kernel = ENSI_PICK(30, false); break;
        default:
            if(E <= 8) kernel = ENSI_PICK(8, true);
            else if(E <= 16) kernel = ENSI_PICK(16, true);
            else if(E <= 24) kernel = ENSI_PICK(24, true);
            else kernel = ENSI_PICK(0, false);
            break;
    }
#undef ENSI_PICK
    GPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const long long want = ((long long) count + ENSI_WARPS - 1) / ENSI_WARPS;
    int per_sm = 1;   // what actually fits (registers and shared memory): one wave of resident CTAs
    GPP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ENSI_WARPS * 32, smem));
    const unsigned grid = (unsigned) std::max<long long>(1, std::min<long long>(want, (long long) sm_count() * std::max(per_sm, 1)));
    GPP_LAUNCH(kernel, grid, ENSI_WARPS * 32, smem, stream, P);
    return GPP_OK;
}
}  // namespace

extern "C" int gpp_ensi_obs_create(const gpp_points* opoints, const float* pobs, const float* psigmas, const float* pbackground, int nE,
                                   const int* member_valid, const gpp_structure* structure, gpp_ensi_obs** out) {
    if(!out) return fail(GPP_ERR_INVALID_ARGUMENT, "out must not be NULL");
    *out = nullptr;
    if(!opoints || !structure || !pobs || !psigmas || !pbackground) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(nE < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative ensemble size");
    GPP_TRY(reject_unset_scales(structure));
    GPP_TRY(ensure_device());
    gpp_ensi_obs* st = new(std::nothrow) gpp_ensi_obs();
    if(!st) return fail(GPP_ERR_RUNTIME, "out of memory");
    const int rc = build_ensi_obs(*st, opoints, pobs, psigmas, pbackground, nE, member_valid, structure);
    if(rc != GPP_OK) { delete st; return rc; }
    *out = st;
    return GPP_OK;
}
extern "C" void gpp_ensi_obs_destroy(gpp_ensi_obs* obs) { delete obs; }

extern "C" int gpp_ensi_valid_members_device(const float* d_background, long long n_points, int nE, int* member_valid, void* stream_) {
    cudaStream_t stream = (cudaStream_t) stream_;
    if(nE < 0 || n_points < 0 || !member_valid) return fail(GPP_ERR_INVALID_ARGUMENT, "bad argument");
    GPP_TRY(ensure_device());
    if(nE == 0) return GPP_OK;
    DeviceBuffer<int> d_flags;
    GPP_TRY(d_flags.alloc(nE));
    GPP_CUDA(cudaMemsetAsync(d_flags.ptr, 0, sizeof(int) * nE, stream));
    const size_t n = (size_t) n_points * nE;
    if(n) GPP_LAUNCH(ensi_invalid_members_kernel, (unsigned) ((n + 255) / 256), 256, 0, stream, d_background, n, nE, d_flags.ptr);
    std::vector<int> flags(nE);
    GPP_TRY(d_flags.download(flags.data(), nE, stream));
    GPP_CUDA(cudaStreamSynchronize(stream));
    for(int e = 0; e < nE; e++) member_valid[e] = !flags[e];
    return GPP_OK;
}

extern "C" int gpp_optimal_interpolation_ensi_device(const gpp_points* cbp, int first, int count, const float* d_background, int nE,
                                                     const gpp_ensi_obs* obs, const gpp_structure* structure, int max_points,
                                                     int allow_extrapolation, float* d_analysis, int* d_num_skipped, void* stream_) {
    cudaStream_t stream = (cudaStream_t) stream_;
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");   // oi_ensi.cpp:124-125
    if(!cbp || !obs || !structure) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(nE != obs->nE) return fail(GPP_ERR_INVALID_ARGUMENT, "the observation state was built for %d members, not %d", obs->nE, nE);
    gpp_points* bp = const_cast<gpp_points*>(cbp);
    if(first < 0 || count < 0 || first + count > bp->n) return fail(GPP_ERR_INVALID_ARGUMENT, "background range out of bounds");
    GPP_TRY(reject_unset_scales(structure));
    if(count == 0 || nE == 0) return GPP_OK;
    GPP_TRY(bp->ensure_on_device());
    if(d_analysis != d_background)   // oi_ensi.cpp:148: the analysis starts as the background
        GPP_CUDA(cudaMemcpyAsync(d_analysis + (size_t) first * nE, d_background + (size_t) first * nE, sizeof(float) * (size_t) count * nE,
                                 cudaMemcpyDeviceToDevice, stream));
    if(obs->E == 0 || obs->table.n_valid == 0) return GPP_OK;
    int kcap = 0;
    GPP_TRY(ensi_kcap(*obs, bp, first, count, structure, max_points, stream, &kcap));
    int* pair = obs->counters.ptr + 2 * (obs->next_slot.fetch_add(1, std::memory_order_relaxed) % ENSI_COUNTER_SLOTS);
    return ensi_launch(*obs, bp, first, count, d_background, d_analysis, structure, kcap, allow_extrapolation, d_num_skipped, pair, stream);
}

extern "C" int gpp_optimal_interpolation_ensi_host(const gpp_points* cbp, const float* background, int nE, const gpp_points* opoints,
                                                   const float* pobs, const float* psigmas, const float* pbackground,
                                                   const gpp_structure* structure, int max_points, int allow_extrapolation,
                                                   float* analysis, int* num_skipped) {
    if(num_skipped) *num_skipped = 0;
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");   // oi_ensi.cpp:124-125
    if(!cbp || !opoints || !structure) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(nE < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative ensemble size");
    gpp_points* bp = const_cast<gpp_points*>(cbp);
    const int nB = bp->n, nS = opoints->n;
    const size_t nBE = (size_t) nB * nE;
    if(nS == 0) {   // oi_ensi.cpp:137-139
        if(nBE) std::memcpy(analysis, background, sizeof(float) * nBE);
        return GPP_OK;
    }
    if(bp->type != opoints->type)
        return fail(GPP_ERR_INVALID_ARGUMENT, "Both background and observations points must be of same coorindate type (lat/lon or x/y)");
    GPP_TRY(reject_unset_scales(structure));
    GPP_TRY(ensure_device());
    if(nBE == 0) return GPP_OK;
    Trace trace("optimal_interpolation_ensi_host");

    // ---- valid members (oi_ensi.cpp:187-201): a member with an invalid value anywhere in the background is left alone.
    // Large fields are scanned on the host (threads) so that their upload can be pipelined with the analysis, block by
    // block; small ones are uploaded at once and scanned on the device.
    DeviceBuffer<float> d_bg, d_out;
    DeviceBuffer<int> d_skipped;
    const int n_chunks = nB >= (1 << 18) ? ENSI_CHUNKS : 1;
    std::vector<int> member_valid(std::max(nE, 1), 1);
    GPP_TRY(d_skipped.alloc(1));
    GPP_CUDA(cudaMemsetAsync(d_skipped.ptr, 0, sizeof(int), 0));
    if(n_chunks > 1) {
        GPP_TRY(d_bg.alloc(nBE));
        std::vector<unsigned char> bad((size_t) nE, 0);
        // (an explicit team size: launchers such as torchrun export OMP_NUM_THREADS=1)
        const int scan_threads = std::max(1, std::min(8, omp_get_num_procs()));
        #pragma omp parallel num_threads(scan_threads)
        {
            std::vector<unsigned char> mine((size_t) nE, 0);
            #pragma omp for schedule(static) nowait
            for(long long p = 0; p < (long long) nB; p++) {
                const float* row = background + (size_t) p * nE;
                for(int e = 0; e < nE; e++) mine[e] |= (unsigned char) !is_valid(row[e]);
            }
            #pragma omp critical
            for(int e = 0; e < nE; e++) bad[e] |= mine[e];
        }
        for(int e = 0; e < nE; e++) member_valid[e] = !bad[e];
        trace.lap("valid-member scan (host)");
    }
    else {
        GPP_TRY(d_bg.upload(background, nBE));
        GPP_TRY(gpp_ensi_valid_members_device(d_bg.ptr, nB, nE, member_valid.data(), nullptr));
        trace.lap("H2D + valid-member scan");
    }
    gpp_ensi_obs st;
    GPP_TRY(build_ensi_obs(st, opoints, pobs, psigmas, pbackground, nE, member_valid.data(), structure));
    GPP_TRY(d_out.alloc(nBE));
    if(n_chunks == 1) GPP_CUDA(cudaMemcpyAsync(d_out.ptr, d_bg.ptr, sizeof(float) * nBE, cudaMemcpyDeviceToDevice, 0));   // oi_ensi.cpp:148
    bool downloaded = false;
    if(st.E > 0 && st.table.n_valid > 0) {
        GPP_TRY(bp->ensure_on_device());
        int kcap = 0;
        GPP_TRY(ensi_kcap(st, bp, 0, nB, structure, max_points, 0, &kcap));   // once for the whole field: every block takes the same kernel
        // blocks of points, each returned to the host (through pinned staging) while the next ones are analysed; every
        // block has its own work counter because consecutive blocks overlap on the device
        std::vector<size_t> bounds(n_chunks + 1);
        for(int c = 0; c <= n_chunks; c++) bounds[c] = (size_t) ((long long) nB * c / n_chunks) * nE;
        static_assert(ENSI_CHUNKS <= (int) ENSI_COUNTER_SLOTS, "one counter pair per block in flight");
        auto launch = [&](int c, cudaStream_t stream) {
            const int first = (int) (bounds[c] / nE), count = (int) ((bounds[c + 1] - bounds[c]) / nE);
            if(n_chunks > 1) {   // this block's slice of the background comes in on the stream that analyses it
                const size_t n = bounds[c + 1] - bounds[c];
                GPP_CUDA(cudaMemcpyAsync(d_bg.ptr + bounds[c], background + bounds[c], sizeof(float) * n, cudaMemcpyHostToDevice, stream));
                GPP_CUDA(cudaMemcpyAsync(d_out.ptr + bounds[c], d_bg.ptr + bounds[c], sizeof(float) * n, cudaMemcpyDeviceToDevice, stream));
            }
            return ensi_launch(st, bp, first, count, d_bg.ptr, d_out.ptr, structure, kcap, allow_extrapolation, d_skipped.ptr,
                               st.counters.ptr + 2 * c, stream);
        };
        if(n_chunks > 1) {
            GPP_TRY(pipelined_download(bounds, launch, d_out.ptr, analysis, true));
            downloaded = true;
        }
        else GPP_TRY(launch(0, 0));
        if(trace.on) { cudaStreamSynchronize(0); trace.lap(downloaded ? "H2D + kernel + D2H (pipelined)" : "kernel"); }
    }
    if(!downloaded) {
        // one block, or nothing was analysed (no valid member / observation): the analysis is d_out or the background (oi_ensi.cpp:148)
        if(n_chunks > 1) std::memcpy(analysis, background, sizeof(float) * nBE);
        else GPP_TRY(d_out.download(analysis, nBE));
    }
    if(num_skipped) GPP_CUDA(cudaMemcpyAsync(num_skipped, d_skipped.ptr, sizeof(int), cudaMemcpyDeviceToHost, 0));
    GPP_CUDA(cudaStreamSynchronize(0));   // `st` and the staging vectors go out of scope after this
    trace.lap("D2H");
    return GPP_OK;
}

// gridpp::optimal_interpolation_ensi_multi_utem, Points overload (oi_ensi_multi.cpp:862-1311)
extern "C" int gpp_optimal_interpolation_ensi_multi_utem_host(const gpp_points* cbp, const float* bratios, const float* background,
                                                              const float* background_corr, int nE, const gpp_points* opoints,
                                                              const float* pobs, const float* pratios, const float* pbackground,
                                                              const float* pbackground_corr, const gpp_structure* structure, int max_points,
                                                              int allow_extrapolation, float* analysis, int* num_skipped) {
    if(num_skipped) *num_skipped = 0;
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");   // :874-875
    if(!cbp || !opoints || !structure) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(nE < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative ensemble size");
    gpp_points* bp = const_cast<gpp_points*>(cbp);
    const int nB = bp->n, nS = opoints->n;
    const size_t nBE = (size_t) nB * nE;
    if(bp->type != opoints->type)
        return fail(GPP_ERR_INVALID_ARGUMENT, "Both background and observations points must be of same coorindate type (lat/lon or x/y)");
    if(nS == 0 || nBE == 0) {   // :895-897
        if(nBE) std::memcpy(analysis, background, sizeof(float) * nBE);
        return GPP_OK;
    }
    if(!bratios || !background || !background_corr || !pobs || !pratios || !pbackground || !pbackground_corr)
        return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    GPP_TRY(reject_unset_scales(structure));
    GPP_TRY(ensure_device());
    Trace trace("optimal_interpolation_ensi_multi_utem_host");
    // ---- members valid in all four ensembles (:929-953)
    DeviceBuffer<float> d_bg, d_bgc, d_br, d_out;
    DeviceBuffer<int> d_skipped;
    // large fields: scanned by host threads so that the uploads can be pipelined with the analysis, block by block
    const int n_chunks = nB >= (1 << 18) ? ENSI_CHUNKS : 1;
    std::vector<int> ok(nE, 1), ok2(nE, 1);
    if(n_chunks > 1) {
        GPP_TRY(d_bg.alloc(nBE));
        GPP_TRY(d_bgc.alloc(nBE));
        std::vector<unsigned char> bad((size_t) nE, 0), bad2((size_t) nE, 0);
        const int scan_threads = std::max(1, std::min(8, omp_get_num_procs()));
        #pragma omp parallel num_threads(scan_threads)
        {
            std::vector<unsigned char> mine((size_t) nE, 0), mine2((size_t) nE, 0);
            #pragma omp for schedule(static) nowait
            for(long long p = 0; p < (long long) nB; p++) {
                const float* row = background + (size_t) p * nE;
                const float* row2 = background_corr + (size_t) p * nE;
                for(int e = 0; e < nE; e++) { mine[e] |= (unsigned char) !is_valid(row[e]); mine2[e] |= (unsigned char) !is_valid(row2[e]); }
            }
            #pragma omp critical
            for(int e = 0; e < nE; e++) { bad[e] |= mine[e]; bad2[e] |= mine2[e]; }
        }
        for(int e = 0; e < nE; e++) { ok[e] = !bad[e]; ok2[e] = !bad2[e]; }
    }
    else {
        GPP_TRY(d_bg.upload(background, nBE));
        GPP_TRY(d_bgc.upload(background_corr, nBE));
        GPP_TRY(gpp_ensi_valid_members_device(d_bg.ptr, nB, nE, ok.data(), nullptr));
        GPP_TRY(gpp_ensi_valid_members_device(d_bgc.ptr, nB, nE, ok2.data(), nullptr));
    }
    gpp_ensi_obs st;
    st.nE = nE;
    st.E = 0;
    for(int e = 0; e < nE; e++) {
        bool good = ok[e] && ok2[e];
        for(int i = 0; i < nS && good; i++) good = is_valid(pbackground[(size_t) i * nE + e]) && is_valid(pbackground_corr[(size_t) i * nE + e]);
        if(!good) continue;
        if(st.E >= ENSI_EMAX)
            return fail(GPP_ERR_NOT_IMPLEMENTED, "optimal_interpolation_ensi_multi_utem supports at most %d valid ensemble members on the device", ENSI_EMAX);
        st.valid_ens[st.E++] = e;
    }
    if(trace.on) { cudaStreamSynchronize(0); trace.lap("H2D + valid-member scan"); }
    const int E = st.E;
    if(E == 0) {
        std::memcpy(analysis, background, sizeof(float) * nBE);
        return GPP_OK;
    }
    // ---- the observation side (:955-994): perturbations about the mean, standardised *_corr perturbations, obs - yhat
    std::vector<char> valid(nS);
    std::vector<double> innov(nS);
    std::vector<float> ratio(pratios, pratios + nS), gY((size_t) nS * E), gYc((size_t) nS * E), row(E);
    const float const_fact = (float) (1 / std::sqrt((double) (E - 1)));
    auto mean_of = [&](const float* v) {   // calc_statistic(Mean), util.cpp:22-37
        float total = 0;
        for(int e = 0; e < E; e++) total += v[e];
        return total / E;
    };
    for(int i = 0; i < nS; i++) {
        for(int e = 0; e < E; e++) row[e] = pbackground[(size_t) i * nE + st.valid_ens[e]];
        const float mean = mean_of(row.data());
        for(int e = 0; e < E; e++) gY[(size_t) i * E + e] = row[e] - mean;
        valid[i] = is_valid(pobs[i]);
        innov[i] = (double) pobs[i] - (double) mean;   // lObs - lYhat, :1157
        for(int e = 0; e < E; e++) row[e] = pbackground_corr[(size_t) i * nE + st.valid_ens[e]];
        const float mean_c = mean_of(row.data());
        float t1 = 0, t2 = 0;   // calc_statistic(Std), util.cpp:40-73
        for(int e = 0; e < E; e++) { const float d = row[e] - row[0]; t1 += d; t2 += d * d; }
        const float m1 = t1 / E, m2 = t2 / E;
        float var = m2 - m1 * m1;
        if(var < 0) var = 0;
        const float std_c = std::sqrt(var);
        for(int e = 0; e < E; e++) gYc[(size_t) i * E + e] = std_c <= 0.0013f ? 0.f : const_fact * (row[e] - mean_c) / std_c;
    }
    std::vector<int> order;
    GPP_TRY(build_obs_table(opoints, valid, innov, ratio, structure->term[0].loc_dist, &st.table, &order));
    if(st.table.n_valid == 0) {
        std::memcpy(analysis, background, sizeof(float) * nBE);
        return GPP_OK;
    }
    std::vector<float> a(order.size() * (size_t) E), b(order.size() * (size_t) E);
    for(size_t slot = 0; slot < order.size(); slot++)
        for(int e = 0; e < E; e++) {
            a[slot * E + e] = gYc[(size_t) order[slot] * E + e];
            b[slot * E + e] = gY[(size_t) order[slot] * E + e];
        }
    GPP_TRY(st.gY.upload(a.data(), a.size()));
    GPP_TRY(st.gY_raw.upload(b.data(), b.size()));
    GPP_TRY(st.counters.alloc(2 * ENSI_COUNTER_SLOTS));
    GPP_CUDA(cudaMemsetAsync(st.counters.ptr, 0, sizeof(int) * 2 * ENSI_COUNTER_SLOTS, 0));
    GPP_TRY(d_skipped.alloc(1));
    GPP_CUDA(cudaMemsetAsync(d_skipped.ptr, 0, sizeof(int), 0));
    GPP_TRY(d_br.upload(bratios, (size_t) nB));
    GPP_TRY(d_out.alloc(nBE));
    if(n_chunks == 1) GPP_CUDA(cudaMemcpyAsync(d_out.ptr, d_bg.ptr, sizeof(float) * nBE, cudaMemcpyDeviceToDevice, 0));
    GPP_TRY(bp->ensure_on_device());
    int kcap = 0;
    GPP_TRY(ensi_kcap(st, bp, 0, nB, structure, max_points, 0, &kcap));
    trace.lap("observation tables");
    std::vector<size_t> bounds(n_chunks + 1);
    for(int c = 0; c <= n_chunks; c++) bounds[c] = (size_t) ((long long) nB * c / n_chunks) * nE;
    static_assert(ENSI_CHUNKS <= (int) ENSI_COUNTER_SLOTS, "one counter pair per block in flight");
    auto launch = [&](int c, cudaStream_t stream) {
        const int first = (int) (bounds[c] / nE), count = (int) ((bounds[c + 1] - bounds[c]) / nE);
        if(n_chunks > 1) {   // this block's slices of the two ensembles come in on the stream that analyses it
            const size_t n = bounds[c + 1] - bounds[c];
            GPP_CUDA(cudaMemcpyAsync(d_bg.ptr + bounds[c], background + bounds[c], sizeof(float) * n, cudaMemcpyHostToDevice, stream));
            GPP_CUDA(cudaMemcpyAsync(d_bgc.ptr + bounds[c], background_corr + bounds[c], sizeof(float) * n, cudaMemcpyHostToDevice, stream));
            GPP_CUDA(cudaMemcpyAsync(d_out.ptr + bounds[c], d_bg.ptr + bounds[c], sizeof(float) * n, cudaMemcpyDeviceToDevice, stream));
        }
        return ensi_launch(st, bp, first, count, d_bg.ptr, d_out.ptr, structure, kcap, allow_extrapolation, d_skipped.ptr, st.counters.ptr + 2 * c, stream,
                           d_bgc.ptr, d_br.ptr);
    };
    if(n_chunks > 1) GPP_TRY(pipelined_download(bounds, launch, d_out.ptr, analysis, true));
    else {
        GPP_TRY(launch(0, 0));
        if(trace.on) { cudaStreamSynchronize(0); trace.lap("kernel"); }
        GPP_TRY(d_out.download(analysis, nBE));
    }
    if(num_skipped) GPP_CUDA(cudaMemcpyAsync(num_skipped, d_skipped.ptr, sizeof(int), cudaMemcpyDeviceToHost, 0));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

#ifdef ENSI_STATS
extern "C" int gpp_debug_ensi_stats(unsigned long long* out, int reset) {
    GPP_CUDA(cudaDeviceSynchronize());
    GPP_CUDA(cudaMemcpyFromSymbol(out, g_ensi_stats, sizeof(unsigned long long) * 4));
    if(reset) {
        unsigned long long zero[4] = {0, 0, 0, 0};
        GPP_CUDA(cudaMemcpyToSymbol(g_ensi_stats, zero, sizeof(zero)));
    }
    return GPP_OK;
}
#endif
