// Deterministic optimal interpolation on the device ("K6" in SURVEY.md).
// Replaces gridpp::optimal_interpolation / optimal_interpolation_full, src/api/oi.cpp:26-412.
//
// One warp analyses one background point:
//   1. gather: scan the bucket-grid cells overlapping the localization box, keep observations strictly inside
//      the box with straight distance <= R and rho > 0 (oi.cpp:229-258, kdtree.cpp:39-62), streaming top-k by
//      (rho, index) in shared memory (oi.cpp:262-273); put the selection in canonical (index) order;
//   2. if the selection differs from the one the warp solved last: assemble P + R from pairwise structure-function
//      evaluations (oi.cpp:298-314) packed over the 32 lanes, then
//   3. eliminate: lane j holds row j of the symmetric augmented matrix [[P+R, rho, d], [rho', 0, 0], [d', 0, 0]]
//      in registers (fp64). k Gauss-Jordan steps give z = (P+R)^-1 d (one component per lane) and leave
//      -rho'(P+R)^-1 rho in the trailing block (the analysis-variance factor, oi.cpp:336-337);
//   4. increment = rho . z (oi.cpp:315-317) -- no inverse is ever formed.
// The register path covers symmetric structure functions with k <= 30. Everything else (k > 30, unlimited
// max_points, non-symmetric structure functions) takes the general kernel: Gauss-Jordan with partial pivoting
// on a per-warp global-memory scratch matrix.
#include "oi.cuh"
#include "structure_field.cuh"

#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>

using namespace gpp;

namespace {

constexpr int FAST_K = 30;          // max observations per point on the register path
constexpr int WARPS_PER_CTA = 8;

#ifndef OI_LRU
#define OI_LRU 16
#endif
#ifndef OI_TOP_SHIFT
#define OI_TOP_SHIFT 2
#endif
constexpr int LRU_ENTRIES = OI_LRU;           // solved systems a warp remembers
constexpr int TOP_SHIFT = OI_TOP_SHIFT;       // the largest work unit is (1 << TOP_SHIFT)^2 tiles of 4 x 4 points
constexpr int N_LEVELS = TOP_SHIFT + 1;       // unit sizes (1 << TOP_SHIFT)^2, ..., 4, 1 tiles
static_assert(TOP_SHIFT >= 0 && TOP_SHIFT <= 3, "unit sizes up to 8 x 8 tiles");

// Work units of the register path, largest first: level l holds units of (1 << (TOP_SHIFT - l))^2 tiles (grids) or that
// many 16-point runs (point sets). end[l] = number of units of levels 0..l, base[l] = first tile row (run) of level l,
// cols[l] = units per row of units.
struct UnitPlan {
    int n_units;
    int end[4], base[4], cols[4];
};
// Head of the launch workspace; both counters are zero between launches (the kernel resets them as it ends).
struct OiWorkHeader {
    int next_unit;
    int warps_done;
};
constexpr size_t OI_WORK_HEADER_BYTES = 256;
// Progress words of the units that can be shared between warps (the last OI_PROG_WORDS units of more than one tile): tiles
// taken from the front by the owner (low half) and from the back by warps that ran out of units (high half). Zero between
// launches.
#ifndef OI_PROG_WORDS
#define OI_PROG_WORDS 65536
#endif
constexpr int PROG_WORDS = OI_PROG_WORDS;
constexpr size_t OI_WORK_LRU_OFFSET = OI_WORK_HEADER_BYTES + sizeof(unsigned) * (size_t) PROG_WORDS;

struct OiParams {
    // background points; gz / gelev / glaf may be NULL: z = 0 (Cartesian), no elevations / land fractions (NaN)
    const float *gx, *gy, *gz, *gelev, *glaf;
    const float* background;
    const float* bvariance;          // may be NULL (= 1)
    float* analysis;
    float* analysis_variance;        // may be NULL
    int first, count;
    ObsView obs;
    gpp_structure s;
    float R;
    int k;                           // max observations per point (> 0)
    int allow_extrapolation;
    int tile_nx;                     // > 0: the range is whole rows of a grid with this row length -> 4 x 4 tiles
    unsigned char* workspace;        // register path: OiWorkHeader + one LruBlock per warp (gpp_oi_workspace_bytes)
    UnitPlan plan;                   // register path: how the range is cut into work units (oi_plan_units)
    int* work_counter;               // Cholesky path: next block of points to hand out (zeroed before the launch)
    // spatially varying structure function (general kernel only): scales at every background point, and the constants of
    // localization_distance(h) = loc_c * h (Toar: (float) (loc_d * h), structure.cpp:603-609)
    const float *sbh, *sbv, *sbw;
    float loc_c;
    double loc_d;
};

// oi.cpp:318-337: optional clamp, then output and analysis variance
__device__ __forceinline__ void write_result(const OiParams& P, int g, float bg, double dx, double a, double dmax, double dmin) {
    float increment = (float) dx;                      // oi.cpp:317
    if(!P.allow_extrapolation) {                       // oi.cpp:318-334
        float maxInc = (float) dmax, minInc = (float) dmin;
        if(maxInc > 0 && increment > maxInc) increment = maxInc;
        else if(maxInc < 0 && increment > 0) increment = maxInc;
        else if(minInc < 0 && increment < minInc) increment = minInc;
        else if(minInc > 0 && increment < 0) increment = minInc;
    }
    P.analysis[g] = __fadd_rn(bg, increment);          // oi.cpp:335
    if(P.analysis_variance) {
        float bv = P.bvariance ? P.bvariance[g] : 1.f;
        P.analysis_variance[g] = (float) __dmul_rn((double) bv, __dsub_rn(1.0, a));   // oi.cpp:336-337
    }
}

// background point g; planes the point set does not have are not read (20 -> 8 bytes per point for a Cartesian grid
// without elevations and land fractions)
__device__ __forceinline__ Pt bg_point(const OiParams& P, int g) {
    Pt p;
    p.x = P.gx[g]; p.y = P.gy[g];
    p.z = P.gz ? P.gz[g] : 0.f;
    p.elev = P.gelev ? P.gelev[g] : NAN;
    p.laf = P.glaf ? P.glaf[g] : NAN;
    return p;
}

constexpr int RUN = 16;             // consecutive background points analysed by one warp
constexpr int NPAIR_LUT = 496;      // 31 * 32 / 2 >= FAST_K * (FAST_K + 1) / 2
constexpr int NCAND = 64;           // capacity of a run's candidate list (2 slots per lane)
constexpr unsigned FULL = 0xffffffffu;

struct FastSmem {
    unsigned long long key[64];     // candidate keys (oi.cuh)
    double M[32 * 33];              // augmented symmetric matrix, row-major with stride 33
    __align__(16) double colbuf[2][64];   // pivot column broadcast: [0] as is, [1] shifted by one entry (entries 32.. are padding for the shifted reads)
    double sd[32];                  // innovations of the selection
    int pos[64];                    // candidate slots in the observation table
    int c_pos[32];                  // the selection in canonical order (ascending original index)
    int c_orig[32];
    float c_rho[32];
    float sx[32], sy[32], sz[32], selev[32], slaf[32];
    float sratio[32];
    // run path: the run's candidate observations (ascending original index) and its points
    float cand_x[NCAND], cand_y[NCAND], cand_z[NCAND], cand_elev[NCAND], cand_laf[NCAND];
    int cand_pos[NCAND], cand_orig[NCAND];
    float px[RUN], py[RUN], pz[RUN], pelev[RUN], plaf[RUN], pbg[RUN];
    int pit[RUN];                   // offset of each point of the run inside the range
};

// State a warp carries from point to point: the last system it solved.
//   prev_orig (per lane r: original index of row r of the solved set, -1 beyond k), prev_k, z (per lane r: component r
//   of (P+R)^-1 d), dmax / dmin (extreme innovations of the set), dirty (non-zero extent of S.M).
struct WarpState {
    int prev_orig, prev_k, dirty;
    int inv_off;                    // analysis variance: row j of (P+R)^-1 of the solved set is S.M[33 j + inv_off .. + k)
    unsigned clock;
    double z, dmax, dmin, avar;
};

#ifdef OI_STATS
// debug build only (profiles/variants.sh -DOI_STATS): [0] points on the run path, [1] selection changes, [2] systems solved,
// [3] selections that needed the full ranking
__device__ unsigned long long g_oi_stats[4];
#define OI_COUNT(i) do { if(lane_id() == 0) atomicAdd(&g_oi_stats[i], 1ull); } while(0)
#else
#define OI_COUNT(i) do {} while(0)
#endif

// A warp's cache of solved systems, in the launch workspace (global memory; 2368 warps x 3.4 KB stay L2-resident): the set
// (canonical original indices), z = (P+R)^-1 d and the innovation extremes. `sig` is an order-independent hash of the set,
// so a lookup compares one word per entry and reads an entry's indices only when the hash matches.
struct LruBlock {
    double z[LRU_ENTRIES][32];
    double dmax[LRU_ENTRIES], dmin[LRU_ENTRIES];
    int orig[LRU_ENTRIES][32];
    unsigned sig[LRU_ENTRIES], stamp[LRU_ENTRIES];   // stamp: last use (0 = empty)
    int k[LRU_ENTRIES];
};
static_assert(LRU_ENTRIES <= 32, "one lane per cache entry");

// Look the canonical set S.c_orig[0..k) up in the warp's cache; on a hit load its solution into W.
__device__ __forceinline__ bool lru_lookup(LruBlock* C, const FastSmem& S, int k, unsigned now, WarpState& W, unsigned& sig) {
    const int lane = (int) lane_id();
    const int mine = lane < k ? S.c_orig[lane] : -1;
    sig = __reduce_add_sync(FULL, (unsigned) (mine + 1) * 0x9E3779B1u);
    unsigned m = __ballot_sync(FULL, lane < LRU_ENTRIES && C->stamp[lane] != 0 && C->sig[lane] == sig && C->k[lane] == k);
    while(m) {
        const int e = __ffs(m) - 1;
        m &= m - 1;
        if(__all_sync(FULL, C->orig[e][lane] == mine)) {
            W.z = C->z[e][lane];
            W.dmax = C->dmax[e];
            W.dmin = C->dmin[e];
            W.prev_orig = mine;
            W.prev_k = k;
            if(lane == 0) C->stamp[e] = now;
            __syncwarp();
            return true;
        }
    }
    return false;
}
__device__ __forceinline__ void lru_store(LruBlock* C, int k, unsigned now, const WarpState& W, unsigned sig) {
    const int lane = (int) lane_id();
    // victim: the least recently used entry (empty ones first)
    const unsigned age = lane < LRU_ENTRIES ? (C->stamp[lane] << 5) | (unsigned) lane : 0xffffffffu;
    const int victim = (int) (__reduce_min_sync(FULL, age) & 31u);
    C->z[victim][lane] = W.z;
    C->orig[victim][lane] = W.prev_orig;
    if(lane == 0) { C->dmax[victim] = W.dmax; C->dmin[victim] = W.dmin; C->k[victim] = k; C->sig[victim] = sig; C->stamp[victim] = now; }
    __syncwarp();
}

// max / min of a 64-bit key over the warp with the single-instruction 32-bit reductions (REDUX): the high words first,
// then the low words of the lanes that hold the extreme high word. Keys are unique, so the result identifies one slot.
__device__ __forceinline__ unsigned long long warp_max64(unsigned long long x) {
    const unsigned hi = (unsigned) (x >> 32), lo = (unsigned) x;
    const unsigned H = __reduce_max_sync(FULL, hi);
    const unsigned L = __reduce_max_sync(FULL, hi == H ? lo : 0u);
    return ((unsigned long long) H << 32) | L;
}
__device__ __forceinline__ unsigned long long warp_min64(unsigned long long x) {
    const unsigned hi = (unsigned) (x >> 32), lo = (unsigned) x;
    const unsigned H = __reduce_min_sync(FULL, hi);
    const unsigned L = __reduce_min_sync(FULL, hi == H ? lo : 0xffffffffu);
    return ((unsigned long long) H << 32) | L;
}

// Assemble P + R for the k observations staged in canonical order in S.c_pos / S.c_rho and solve for
// z = (P+R)^-1 d by Gauss-Jordan in registers (lane j = row j of the symmetric augmented matrix
// [[P+R, rho, d], [rho', 0, 0], [d', 0, 0]]); also leaves rho'(P+R)^-1 rho in W.avar. oi.cpp:298-317,336.
// INV (the analysis variance is wanted): the same elimination also leaves (P+R)^-1 behind. The column of the identity
// that a Gauss-Jordan step would touch in [A | I] takes the place of the eliminated pivot column (it enters the register
// window at the far end as the window shifts), the pivot row is normalised by the same FMA pass (its factor is
// 1 - 1/pivot), and because the partly swept matrix stays symmetric up to sign -- the (swept, unswept) block of
// [[A11^-1, A11^-1 A12], [-A21 A11^-1, S]] is minus the transpose of the (unswept, swept) one -- the pivot ROW is still
// the broadcast pivot COLUMN, negated for the swept part: the line buffer carries my(lane) at [lane] and -my(lane) at
// [30 + lane]. After k steps the window holds row `lane` of the inverse in its last k slots. With it the variance
// factor rho' (P+R)^-1 rho of every further point that selects the same set is a k x k matrix-vector product
// (variance_factor) instead of a new elimination.
template <int SMODE, bool INV>
__device__ __noinline__ void solve_selected(const OiParams& P, FastSmem& S, const unsigned short* lut, int k, WarpState& W) {
    const int lane = (int) lane_id();
    // ---- stage the selected observations
    if(lane < k) {
        const int pos = S.c_pos[lane];
        S.sx[lane] = P.obs.x[pos]; S.sy[lane] = P.obs.y[pos]; S.sz[lane] = P.obs.z[pos];
        S.selev[lane] = P.obs.elev[pos]; S.slaf[lane] = P.obs.laf[pos];
        S.sratio[lane] = P.obs.ratio[pos];
        S.sd[lane] = P.obs.innov[pos];
    }
    if(k < W.dirty) {   // keep everything outside [0,k) U {30,31} at zero
        for(int i = k; i < W.dirty; i++) S.M[lane * 33 + i] = 0.0;
        if(lane >= k && lane < W.dirty)
            for(int i = 0; i < 32; i++) S.M[lane * 33 + i] = 0.0;
    }
    W.dirty = k;
    __syncwarp();
    // ---- pairwise correlations (oi.cpp:298-314), lower triangle incl. diagonal packed over the lanes
    const int npairs = k * (k + 1) / 2;
    for(int p = lane; p < npairs; p += 32) {
        const int code = lut[p];
        const int j = code >> 8, i = code & 255;
        const Pt a = {S.sx[j], S.sy[j], S.sz[j], S.selev[j], S.slaf[j]};
        const Pt b = {S.sx[i], S.sy[i], S.sz[i], S.selev[i], S.slaf[i]};
        const float hdist = straight_distance(a.x, a.y, a.z, b.x, b.y, b.z);
        double v = (double) corr_call<SMODE>(P.s, a, b, hdist);
        if(i == j) v = __dadd_rn(v, (double) S.sratio[j]);   // lP + lR, oi.cpp:315
        S.M[j * 33 + i] = v;
        S.M[i * 33 + j] = v;
    }
    if(lane < k) {
        const double r = (double) S.c_rho[lane], d = S.sd[lane];
        S.M[30 * 33 + lane] = r; S.M[lane * 33 + 30] = r;
        S.M[31 * 33 + lane] = d; S.M[lane * 33 + 31] = d;
    }
    // max / min innovation for the optional clamp (oi.cpp:319-320)
    double dmax = lane < k ? S.sd[lane] : -INFINITY, dmin = lane < k ? S.sd[lane] : INFINITY;
    #pragma unroll
    for(int off = 16; off > 0; off >>= 1) {
        dmax = fmax(dmax, shfl_double(dmax, lane ^ off));
        dmin = fmin(dmin, shfl_double(dmin, lane ^ off));
    }
    W.dmax = dmax;
    W.dmin = dmin;
    __syncwarp();
    // ---- Gauss-Jordan on the k observation rows, pivots on the diagonal (SPD: no pivoting needed). By symmetry
    // of the not-yet-eliminated block the pivot ROW equals the pivot COLUMN, which is spread over the lanes: one
    // shared store per lane broadcasts it. The loop is ROLLED (a fully unrolled triangular version is 25 KB of
    // code and stalls on instruction fetch): the lane's row lives in a register window that shifts left by one
    // column per step, so that the pivot column is always a[0] and every step runs the same 30-FMA body.
    double a[FAST_K];
    #pragma unroll
    for(int i = 0; i < FAST_K; i++) a[i] = S.M[lane * 33 + i];
    double rr = S.M[lane * 33 + 30], rd = S.M[lane * 33 + 31];   // the rho and d columns of this lane's row
    double my_inv = 0.0;
    for(int c = 0; c < k; c++) {
        const double my = a[0];
        const double inv = fast_rcp(shfl_double(my, c));
        // the pivot column is published twice, the second copy shifted by one entry, so that the entries of columns
        // c+1 .. c+29 can always be read as aligned 16-byte pairs
        if(INV) {
            if(lane < FAST_K) {
                S.colbuf[0][lane] = my;
                S.colbuf[0][FAST_K + lane] = -my;
                if(lane > 0) S.colbuf[1][lane - 1] = my;
                S.colbuf[1][FAST_K - 1 + lane] = -my;
            }
            else S.colbuf[0][32 + lane] = my;    // rows rho and d of the augmented matrix: entries 62, 63
        }
        else {
            S.colbuf[0][lane] = my;
            if(lane > 0) S.colbuf[1][lane - 1] = my;
        }
        const double f = lane == c ? (INV ? 1.0 - inv : 0.0) : my * inv;
        if(lane == c) my_inv = inv;
        __syncwarp();
        const int base = c + 1;       // pivot-row entry of column base + t (columns past 29 feed slots that are never read)
        const double2* src = reinterpret_cast<const double2*>((base & 1) ? &S.colbuf[1][base - 1] : &S.colbuf[0][base]);
        #pragma unroll
        for(int q = 0; q < FAST_K / 2; q++) {
            const double2 t = src[q];
            a[2 * q] = fma(-f, t.x, a[2 * q + 1]);
            if(2 * q + 2 < FAST_K) a[2 * q + 1] = fma(-f, t.y, a[2 * q + 2]);
        }
        if(INV) a[FAST_K - 1] = (lane == c ? 1.0 : 0.0) - f;     // column c of the identity after this step
        rr = fma(-f, S.colbuf[0][INV ? 62 : 30], rr);
        rd = fma(-f, S.colbuf[0][INV ? 63 : 31], rd);
        __syncwarp();                 // the next step overwrites the buffers
    }
    W.z = lane < k ? (INV ? rd : rd * my_inv) : 0.0;   // z = (P+R)^-1 d, one component per lane (INV: the pivot rows are normalised)
    W.avar = -shfl_double(rr, 30);            // rho'(P+R)^-1 rho (oi.cpp:336)
    W.prev_orig = lane < k ? S.c_orig[lane] : -1;
    W.prev_k = k;
    if(INV) {
        // row `lane` of the inverse: window slots FAST_K - k .. FAST_K - 1 (static register indices; the offset is applied on
        // the shared-memory side). S.M is scratch until the next assembly, which re-zeroes what it needs (dirty = 32).
        #pragma unroll
        for(int q = 0; q < FAST_K; q++) S.M[lane * 33 + q] = a[q];
        W.inv_off = FAST_K - k;
        W.dirty = 32;
        __syncwarp();
    }
}

// rho' (P+R)^-1 rho (oi.cpp:336) for the point whose correlations are in S.c_rho (canonical order) from the inverse the
// last elimination left in S.M: lane j forms row j of the product, the warp adds up rho_j times it
__device__ __forceinline__ double variance_factor(const FastSmem& S, int k, const WarpState& W) {
    const int lane = (int) lane_id();
    double u = 0.0;
    if(lane < k) {
        const double* row = S.M + lane * 33 + W.inv_off;
        for(int i = 0; i < k; i++) u = fma(row[i], (double) S.c_rho[i], u);
        u *= (double) S.c_rho[lane];
    }
    #pragma unroll
    for(int off = 16; off > 0; off >>= 1) u += shfl_double(u, lane ^ off);
    return u;
}

// increment = rho . z (oi.cpp:315-317) with rho in canonical order in S.c_rho; lane 0 writes the result
__device__ __forceinline__ void finish_point(const OiParams& P, const FastSmem& S, int g, float bg, int k, const WarpState& W) {
    const int lane = (int) lane_id();
    double dx = lane < k ? (double) S.c_rho[lane] * W.z : 0.0;
    #pragma unroll
    for(int off = 16; off > 0; off >>= 1) dx += shfl_double(dx, lane ^ off);
    if(lane == 0) write_result(P, g, bg, dx, W.avar, W.dmax, W.dmin);
}
__device__ __forceinline__ void keep_background(const OiParams& P, int g, float bg) {
    if(lane_id() == 0) {   // oi.cpp:223,234-237,284-287: the analysis stays at the background
        P.analysis[g] = bg;
        if(P.analysis_variance) P.analysis_variance[g] = P.bvariance ? P.bvariance[g] : 1.f;
    }
}

// Per-point path: gather from the bucket grid, select, canonicalise, reuse or solve. Used when a run's candidate
// list does not fit NCAND slots (dense observations, scattered points).
template <int SMODE>
__device__ __noinline__ void analyse_point(const OiParams& P, FastSmem& S, const unsigned short* lut, LruBlock* cache, int g, WarpState& W) {
    const int lane = (int) lane_id();
    const CandBuf cb = {S.key, S.pos};
    const bool need_var = P.analysis_variance != nullptr;
    const float bg = P.background[g];
    int k = 0;
    if(is_valid(bg)) {   // oi.cpp:223
        const Pt p1 = bg_point(P, g);
        k = gather_candidates<SMODE, 2>(P.obs, P.s, p1, P.R, P.k, cb);
    }
    if(k == 0) { keep_background(P, g, bg); return; }
    // ---- canonical order of the selection: ascending original index
    {
        const unsigned long long my_key = lane < k ? S.key[lane] : 0ull;
        const int my_pos = lane < k ? S.pos[lane] : 0;
        const unsigned my_inv = (unsigned) my_key;    // 0x7fffffff - original index
        int crank = 0;
        const unsigned* lo_words = reinterpret_cast<const unsigned*>(S.key);
        #pragma unroll 4
        for(int m = 0; m < k; m++) crank += lo_words[2 * m] > my_inv;
        __syncwarp();
        if(lane < k) {
            S.c_orig[crank] = cand_key_orig(my_key);
            S.c_rho[crank] = cand_key_rho(my_key);
            S.c_pos[crank] = my_pos;
        }
        __syncwarp();
    }
    const int c_orig = lane < k ? S.c_orig[lane] : -1;
    const bool same = __all_sync(FULL, c_orig == W.prev_orig) && k == W.prev_k;
    if(need_var) {
        if(!same) solve_selected<SMODE, true>(P, S, lut, k, W);
        else W.avar = variance_factor(S, k, W);
    }
    else if(!same) {
        const unsigned now = ++W.clock;
        unsigned sig;
        if(!lru_lookup(cache, S, k, now, W, sig)) {
            solve_selected<SMODE, false>(P, S, lut, k, W);
            lru_store(cache, k, now, W, sig);
        }
    }
    finish_point(P, S, g, bg, k, W);
    __syncwarp();
}

// The reference's predicate and correlation (kdtree.cpp:46-53,247-260; oi.cpp:250-258) for the two candidates a lane holds,
// as selection keys (0 = not a candidate of this point). SMODE 1 (one Barnes term with an active horizontal scale, no
// cross-validation) is evaluated in straight-line code for both slots at once: no call, no divergent branch, two
// independent exp() chains in flight.
template <int SMODE>
__device__ __forceinline__ void slot_keys(const OiParams& P, const Pt& p1, const Pt& q0, const Pt& q1, bool h0, bool h1, int orig0,
                                          int orig1, unsigned long long& key0, unsigned long long& key1) {
    const float lo0 = __fsub_rn(p1.x, P.R), lo1 = __fsub_rn(p1.y, P.R), lo2 = __fsub_rn(p1.z, P.R);
    const float hi0 = __fadd_rn(p1.x, P.R), hi1 = __fadd_rn(p1.y, P.R), hi2 = __fadd_rn(p1.z, P.R);
    const bool b0 = h0 && q0.x > lo0 && q0.x < hi0 && q0.y > lo1 && q0.y < hi1 && q0.z > lo2 && q0.z < hi2;
    const bool b1 = h1 && q1.x > lo0 && q1.x < hi0 && q1.y > lo1 && q1.y < hi1 && q1.z > lo2 && q1.z < hi2;
    const float d0 = straight_distance(q0.x, q0.y, q0.z, p1.x, p1.y, p1.z);
    const float d1 = straight_distance(q1.x, q1.y, q1.z, p1.x, p1.y, p1.z);
    key0 = 0ull;
    key1 = 0ull;
    if(SMODE == 1) {
        // BarnesStructure::corr, structure.cpp:214-228 with barnes_rho :26-34. P.R is the term's localization distance, so the
        // `hdist > localization_distance` test is the d <= R below; the horizontal scale is valid and positive
        // (structure_mode), the distance of an accepted candidate is finite: barnes_rho's early returns cannot trigger.
        const gpp_structure_term& t = P.s.term[0];
        const double v0 = (double) __fdiv_rn(d0, t.h), v1 = (double) __fdiv_rn(d1, t.h);
        float rho0 = (float) exp_nonpos(__dmul_rn(__dmul_rn(-0.5, v0), v0));
        float rho1 = (float) exp_nonpos(__dmul_rn(__dmul_rn(-0.5, v1), v1));
        if(is_valid(t.v) && t.v != 0.f && is_valid(p1.elev)) {   // vertical term (uniform over the warp)
            if(is_valid(q0.elev)) rho0 = __fmul_rn(rho0, term_rho(GPP_STRUCT_BARNES, __fsub_rn(p1.elev, q0.elev), t.v));
            if(is_valid(q1.elev)) rho1 = __fmul_rn(rho1, term_rho(GPP_STRUCT_BARNES, __fsub_rn(p1.elev, q1.elev), t.v));
        }
        if(is_valid(t.w) && t.w != 0.f && is_valid(p1.laf)) {     // land / sea term
            if(is_valid(q0.laf)) rho0 = __fmul_rn(rho0, term_rho(GPP_STRUCT_BARNES, __fsub_rn(p1.laf, q0.laf), t.w));
            if(is_valid(q1.laf)) rho1 = __fmul_rn(rho1, term_rho(GPP_STRUCT_BARNES, __fsub_rn(p1.laf, q1.laf), t.w));
        }
        if(b0 && d0 <= P.R && rho0 > 0.f) key0 = cand_key(rho0, orig0);
        if(b1 && d1 <= P.R && rho1 > 0.f) key1 = cand_key(rho1, orig1);
    }
    else {
        if(b0 && d0 <= P.R) {
            const float rho = corr_background_call<SMODE>(P.s, p1, q0, d0);
            if(rho > 0.f) key0 = cand_key(rho, orig0);
        }
        if(b1 && d1 <= P.R) {
            const float rho = corr_background_call<SMODE>(P.s, p1, q1, d1);
            if(rho > 0.f) key1 = cand_key(rho, orig1);
        }
    }
}

// One warp walks RUN consecutive background points.
//
// Solution reuse: neighbouring points usually select the SAME set of observations (the set only changes when the
// point crosses a boundary of the order-k Voronoi diagram of the observations). With the selection in canonical
// order (ascending original index), z = (P+R)^-1 d depends on the set alone and the increment is the k-term dot
// product rho . z (oi.cpp:315-316: lG * inv(lP+lR) * (lObs - lY)). The warp keeps (set, z) of the last system it
// solved. Every point computes its increment with the same dot product in the same order, so results do not depend
// on where a run starts or on which path found the set. The analysis variance needs rho'(P+R)^-1 rho, which depends
// on the point: when it is requested the elimination also leaves (P+R)^-1 behind (solve_selected<.., true>) and every
// further point with the same set evaluates the quadratic form with it.
//
// Run path: the observations that can be within R of ANY point of the run (distance to the run's centre
// <= R + extent) are gathered ONCE, sorted by original index, and kept lane-resident (2 slots per lane). Per point
// each lane evaluates the reference's exact predicate (strict box, distance <= R, rho > 0, oi.cpp:233-258) for its
// slots. The selection (oi.cpp:262-273) is MAINTAINED rather than recomputed: it starts from the set solved last and is
// repaired by exchanges -- while the worst member's key is below the best outsider's, the two trade places; each test is
// four single-instruction warp reductions -- and the full ranking only runs when the set has changed by more than a few
// members.
//
// Work distribution: units of (1 << s)^2 tiles, s = TOP_SHIFT .. 0, handed out through an atomic counter, the largest
// units first (oi_plan_units): the last rows of the range are cut into ever smaller units so that the warps run out of
// work at about the same time, while most of the grid is walked in large units, which is what the reuse wants.
template <int SMODE>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 2) oi_fast_kernel(const __grid_constant__ OiParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem& S = reinterpret_cast<FastSmem*>(smem_raw)[threadIdx.x >> 5];
    unsigned short* lut = reinterpret_cast<unsigned short*>(smem_raw + sizeof(FastSmem) * WARPS_PER_CTA);
    const int lane = (int) lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    const int warp_global = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const bool need_var = P.analysis_variance != nullptr;

    // pair index p -> (row j, column i <= j) of the lower triangle, packed over the lanes during assembly
    for(int p = threadIdx.x; p < NPAIR_LUT; p += blockDim.x) {
        int j = (int) ((sqrtf(8.f * (float) p + 1.f) - 1.f) * 0.5f);
        while(j * (j + 1) / 2 > p) j--;
        while((j + 1) * (j + 2) / 2 <= p) j++;
        lut[p] = (unsigned short) ((j << 8) | (p - j * (j + 1) / 2));
    }
    for(int e = lane; e < 32 * 33; e += 32) S.M[e] = 0.0;
    __syncthreads();
    WarpState W;
    W.prev_orig = -2; W.prev_k = -1; W.dirty = 0; W.clock = 0; W.inv_off = 0;
    W.z = 0.0; W.dmax = 0.0; W.dmin = 0.0; W.avar = 0.0;

    OiWorkHeader* hdr = reinterpret_cast<OiWorkHeader*>(P.workspace);
    unsigned* prog = reinterpret_cast<unsigned*>(P.workspace + OI_WORK_HEADER_BYTES);
    LruBlock* cache = reinterpret_cast<LruBlock*>(P.workspace + OI_WORK_LRU_OFFSET) + warp_global;
    // units [split_from, n_multi) hold more than one tile and have a progress word
    const int n_multi = N_LEVELS > 1 ? P.plan.end[N_LEVELS - 2] : 0;
    const int split_from = max(0, n_multi - PROG_WORDS);
    int scan_from = split_from;   // units before this one are known to be exhausted (only moves forward)
    if(lane < LRU_ENTRIES) cache->stamp[lane] = 0;
    __syncwarp();
    const int tiles_x = P.tile_nx > 0 ? (P.tile_nx + 3) / 4 : 0;
    const int rows = P.tile_nx > 0 ? P.count / P.tile_nx : 0;
    const int tiles_y = (rows + 3) / 4;
    const int n_runs = P.tile_nx > 0 ? tiles_x * tiles_y : (P.count + RUN - 1) / RUN;
    for(;;) {
    int unit = 0;
    if(lane == 0) unit = atomicAdd(&hdr->next_unit, 1);
    unit = __shfl_sync(FULL, unit, 0);
    bool helper = false;
    if(unit >= P.plan.n_units) {
        // Out of units: help with the unit that has the most tiles left, taking them from its far end. A unit's cost varies
        // several-fold with the local churn of the selection, and without this the launch ends when the slowest large
        // unit does -- which is what a row block of a multi-GPU run, with only a few large units per warp, is timed by.
        int my_left = 0, my_u = 0, lowest = n_multi;
        #pragma unroll 4
        for(int u0 = scan_from; u0 < n_multi; u0 += 32) {
            const int u = u0 + lane;
            if(u < n_multi) {
                const unsigned w = __ldcg(&prog[u - split_from]);
                int level = 0;
                #pragma unroll
                for(int l = 0; l < N_LEVELS - 1; l++) level += u >= P.plan.end[l];
                const int left = (1 << (2 * (TOP_SHIFT - level))) - (int) (w & 0xffffu) - (int) (w >> 16);
                if(left > 0) lowest = min(lowest, u);
                if(left > my_left) { my_left = left; my_u = u; }
            }
        }
        scan_from = max(scan_from, __reduce_min_sync(FULL, lowest) & ~31);   // everything before it has been taken for good
        const int best_left = __reduce_max_sync(FULL, my_left);
        if(best_left == 0) break;
        // several warps usually arrive here together: spread them over the units that tie
        const unsigned ties = __ballot_sync(FULL, my_left == best_left);
        unsigned pick = ties;
        for(int skip = warp_global % __popc(ties); skip > 0; skip--) pick &= pick - 1;
        unit = __shfl_sync(FULL, my_u, __ffs(pick) - 1);
        helper = true;
    }
    // ---- the unit's level (size), and where it lies
    int level = 0;
    #pragma unroll
    for(int l = 0; l < N_LEVELS - 1; l++) level += unit >= P.plan.end[l];
    const int local = unit - (level > 0 ? P.plan.end[level - 1] : 0);
    const int sh = TOP_SHIFT - level;                     // the unit is (1 << sh) x (1 << sh) tiles, or 1 << 2 sh runs
    const int n_sub = 1 << (2 * sh);
    const int ucy = P.tile_nx > 0 ? local / P.plan.cols[level] : 0, ucx = local - ucy * P.plan.cols[level];
    const bool shared_unit = unit >= split_from && unit < n_multi;
    for(int j_seq = 0; j_seq < n_sub; j_seq++) {
        int j_run = j_seq;
        if(shared_unit) {
            unsigned w = 0;
            if(lane == 0) w = atomicAdd(&prog[unit - split_from], helper ? 0x10000u : 1u);
            w = __shfl_sync(FULL, w, 0);
            const int front = (int) (w & 0xffffu), back = (int) (w >> 16);
            if(front + back >= n_sub) break;   // every tile of the unit has been taken
            j_run = helper ? n_sub - 1 - back : front;
        }
        // ---- the run's points (offsets into the range), and a bounding sphere
        int npts, my_it = 0;
        if(P.tile_nx > 0) {
            // a unit is walked in serpentine order. The region over which one observation set is selected is a few points
            // across (C3: ~30 points), so the squarer the unit, the fewer regions are cut by its edge and solved again by
            // another warp.
            const int tyy = j_run >> sh, txx = j_run & ((1 << sh) - 1);
            const int ty = P.plan.base[level] + (ucy << sh) + tyy, tx = (ucx << sh) + ((tyy & 1) ? (1 << sh) - 1 - txx : txx);
            if(ty >= tiles_y || tx >= tiles_x) continue;
            const int h = min(4, rows - 4 * ty), wdt = min(4, P.tile_nx - 4 * tx);
            npts = h * wdt;
            if(lane < npts) {
                const int i = lane / wdt, j = lane - i * wdt;
                my_it = (4 * ty + i) * P.tile_nx + 4 * tx + ((i & 1) ? wdt - 1 - j : j);
            }
        }
        else {
            const int run = P.plan.base[level] + (local << (2 * sh)) + j_run;
            if(run >= n_runs) continue;
            npts = min(RUN, P.count - run * RUN);
            my_it = run * RUN + lane;
        }
        float ext = 0.f;
        {
            float x = 0.f, y = 0.f, z = 0.f;
            if(lane < npts) {
                const int g = P.first + my_it;
                const Pt p = bg_point(P, g);
                x = p.x; y = p.y; z = p.z;
                S.px[lane] = x; S.py[lane] = y; S.pz[lane] = z;
                S.pelev[lane] = p.elev; S.plaf[lane] = p.laf; S.pbg[lane] = P.background[g];
                S.pit[lane] = my_it;
            }
            __syncwarp();
            // centre of the run: mean of its points (any centre is correct; a central one keeps the list short)
            float sx_ = x, sy_ = y, sz_ = z;
            #pragma unroll
            for(int off = 16; off > 0; off >>= 1) {
                sx_ += __shfl_xor_sync(FULL, sx_, off);
                sy_ += __shfl_xor_sync(FULL, sy_, off);
                sz_ += __shfl_xor_sync(FULL, sz_, off);
            }
            const float inv_n = 1.f / (float) npts;
            const float cx = sx_ * inv_n, cy = sy_ * inv_n, cz = sz_ * inv_n;
            if(lane < npts) ext = sqrtf((x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz));
            #pragma unroll
            for(int off = 16; off > 0; off >>= 1) ext = fmaxf(ext, __shfl_xor_sync(FULL, ext, off));
            // ---- candidate list: every table observation within Rs = R + ext (+ rounding slack) of the centre
            const float Rs = (P.R + ext) * 1.0001f + 1e-3f;
            int nL = 0;
            bool overflow = !(Rs < INFINITY);
            if(!overflow) {
                const ObsView& obs = P.obs;
                const int cx0 = cell_coord(obs.geom, 0, cx - Rs), cx1 = cell_coord(obs.geom, 0, cx + Rs);
                const int cy0 = cell_coord(obs.geom, 1, cy - Rs), cy1 = cell_coord(obs.geom, 1, cy + Rs);
                const int cz0 = cell_coord(obs.geom, 2, cz - Rs), cz1 = cell_coord(obs.geom, 2, cz + Rs);
                for(int czi = cz0; czi <= cz1 && !overflow; czi++)
                    for(int cyi = cy0; cyi <= cy1 && !overflow; cyi++) {
                        const int base = (czi * obs.geom.n[1] + cyi) * obs.geom.n[0];
                        const int s0 = obs.cell_start[base + cx0], s1 = obs.cell_start[base + cx1 + 1];
                        for(int chunk = s0; chunk < s1; chunk += 32) {
                            const int i = chunk + lane;
                            bool ok = i < s1;
                            if(ok) {
                                const float dx = obs.x[i] - cx, dy = obs.y[i] - cy, dz = obs.z[i] - cz;
                                ok = sqrtf(dx * dx + dy * dy + dz * dz) <= Rs;
                            }
                            const unsigned mask = __ballot_sync(FULL, ok);
                            const int add = __popc(mask);
                            if(nL + add > NCAND) { overflow = true; break; }
                            if(ok) {
                                const int slot = nL + __popc(mask & lt_mask);
                                S.key[slot] = (unsigned long long) (unsigned) obs.orig[i];
                                S.pos[slot] = i;
                            }
                            nL += add;
                        }
                    }
            }
            __syncwarp();
            if(overflow) {
                // too many candidates for the lane-resident list: analyse the run point by point
                for(int i = 0; i < npts; i++) analyse_point<SMODE>(P, S, lut, cache, P.first + S.pit[i], W);
                continue;
            }
            // ---- sort the candidates by original index (rank by counting) and load their coordinates
            {
                const unsigned o0 = lane < nL ? (unsigned) S.key[lane] : 0xffffffffu;
                const unsigned o1 = lane + 32 < nL ? (unsigned) S.key[lane + 32] : 0xffffffffu;
                const int p0 = lane < nL ? S.pos[lane] : 0, p1 = lane + 32 < nL ? S.pos[lane + 32] : 0;
                int r0 = 0, r1 = 0;
                const unsigned* lo_words = reinterpret_cast<const unsigned*>(S.key);
                #pragma unroll 4
                for(int m = 0; m < nL; m++) {
                    const unsigned om = lo_words[2 * m];
                    r0 += om < o0;
                    r1 += om < o1;
                }
                if(lane < nL) {
                    S.cand_orig[r0] = (int) o0; S.cand_pos[r0] = p0;
                    S.cand_x[r0] = P.obs.x[p0]; S.cand_y[r0] = P.obs.y[p0]; S.cand_z[r0] = P.obs.z[p0];
                    S.cand_elev[r0] = P.obs.elev[p0]; S.cand_laf[r0] = P.obs.laf[p0];
                }
                if(lane + 32 < nL) {
                    S.cand_orig[r1] = (int) o1; S.cand_pos[r1] = p1;
                    S.cand_x[r1] = P.obs.x[p1]; S.cand_y[r1] = P.obs.y[p1]; S.cand_z[r1] = P.obs.z[p1];
                    S.cand_elev[r1] = P.obs.elev[p1]; S.cand_laf[r1] = P.obs.laf[p1];
                }
                __syncwarp();
            }
            // ---- this lane's two candidates, and where the previously solved set sits in the new list
            const bool h0 = lane < nL, h1 = lane + 32 < nL;
            const Pt q0 = {h0 ? S.cand_x[lane] : 0.f, h0 ? S.cand_y[lane] : 0.f, h0 ? S.cand_z[lane] : 0.f,
                           h0 ? S.cand_elev[lane] : 0.f, h0 ? S.cand_laf[lane] : 0.f};
            const Pt q1 = {h1 ? S.cand_x[lane + 32] : 0.f, h1 ? S.cand_y[lane + 32] : 0.f, h1 ? S.cand_z[lane + 32] : 0.f,
                           h1 ? S.cand_elev[lane + 32] : 0.f, h1 ? S.cand_laf[lane + 32] : 0.f};
            const int orig0 = h0 ? S.cand_orig[lane] : -1, orig1 = h1 ? S.cand_orig[lane + 32] : -1;
            // members of the solved set among the slots (T0/T1), valid only if every member is in the list
            unsigned T0 = 0, T1 = 0;
            {
                bool m0 = false, m1 = false;
                for(int r = 0; r < 32; r++) {
                    const int po = __shfl_sync(FULL, W.prev_orig, r);
                    if(po < 0) break;
                    m0 = m0 || po == orig0;
                    m1 = m1 || po == orig1;
                }
                T0 = __ballot_sync(FULL, m0);
                T1 = __ballot_sync(FULL, m1);
                if(__popc(T0) + __popc(T1) != W.prev_k) { T0 = 0; T1 = 0; }
            }
            // ---- the points of the run
            for(int i = 0; i < npts; i++) {
                const int g = P.first + S.pit[i];
                const float bg = S.pbg[i];
                if(!is_valid(bg)) { keep_background(P, g, bg); continue; }   // oi.cpp:223
                const Pt p1 = {S.px[i], S.py[i], S.pz[i], S.pelev[i], S.plaf[i]};
                unsigned long long key0, key1;
                slot_keys<SMODE>(P, p1, q0, q1, h0, h1, orig0, orig1, key0, key1);
                const unsigned v0 = __ballot_sync(FULL, key0 != 0ull), v1 = __ballot_sync(FULL, key1 != 0ull);
                const int nv = __popc(v0) + __popc(v1);
                if(nv == 0) { keep_background(P, g, bg); continue; }   // oi.cpp:234-237,284-287
                // ---- selection (oi.cpp:262-273) as slot masks
                unsigned sel0 = v0, sel1 = v1;
                int k = nv;
                if(nv > P.k) {
                    k = P.k;
                    // start from the members of the solved set that are still candidates and repair by exchanges
                    sel0 = T0 & v0;
                    sel1 = T1 & v1;
                    int cnt = __popc(sel0) + __popc(sel1);
                    bool settled = false;
                    if(cnt + 3 >= k) {
                        for(int it = 0; it < 8; it++) {
                            const bool in0 = (sel0 >> lane) & 1u, in1 = (sel1 >> lane) & 1u;
                            const unsigned long long out_best = max(in0 ? 0ull : key0, in1 ? 0ull : key1);
                            const unsigned long long BO = warp_max64(out_best);   // never 0: more candidates than members
                            if(cnt < k) {   // a member dropped out of reach: the best outsider takes its place
                                sel0 |= __ballot_sync(FULL, !in0 && key0 == BO);
                                sel1 |= __ballot_sync(FULL, !in1 && key1 == BO);
                                cnt++;
                                continue;
                            }
                            const unsigned long long in_worst = min(in0 ? key0 : ~0ull, in1 ? key1 : ~0ull);
                            const unsigned long long WI = warp_min64(in_worst);
                            if(WI > BO) { settled = true; break; }
                            sel0 = (sel0 & ~__ballot_sync(FULL, in0 && key0 == WI)) | __ballot_sync(FULL, !in0 && key0 == BO);
                            sel1 = (sel1 & ~__ballot_sync(FULL, in1 && key1 == WI)) | __ballot_sync(FULL, !in1 && key1 == BO);
                        }
                    }
                    if(!settled) {
                        // rank the valid keys; the k largest are selected
                        OI_COUNT(3);
                        S.key[lane] = key0;
                        S.key[lane + 32] = key1;
                        __syncwarp();
                        int r0 = 0, r1 = 0;
                        #pragma unroll 4
                        for(int m = 0; m < nL; m++) {
                            const unsigned long long km = S.key[m];
                            r0 += km > key0;
                            r1 += km > key1;
                        }
                        sel0 = __ballot_sync(FULL, key0 != 0ull && r0 < k);
                        sel1 = __ballot_sync(FULL, key1 != 0ull && r1 < k);
                        __syncwarp();
                    }
                }
                // ---- rho of the selection in canonical (slot = original index) order
                const int n_sel0 = __popc(sel0);
                const int row0 = __popc(sel0 & lt_mask), row1 = n_sel0 + __popc(sel1 & lt_mask);
                const bool s0 = (sel0 >> lane) & 1u, s1 = (sel1 >> lane) & 1u;
                if(s0) S.c_rho[row0] = cand_key_rho(key0);
                if(s1) S.c_rho[row1] = cand_key_rho(key1);
                const bool same = sel0 == T0 && sel1 == T1;
                OI_COUNT(0);
                if(!same) {
                    OI_COUNT(1);
                    if(s0) { S.c_pos[row0] = S.cand_pos[lane]; S.c_orig[row0] = orig0; }
                    if(s1) { S.c_pos[row1] = S.cand_pos[lane + 32]; S.c_orig[row1] = orig1; }
                    __syncwarp();
                    if(need_var) solve_selected<SMODE, true>(P, S, lut, k, W);   // (the cache holds no inverses)
                    else {
                        const unsigned now = ++W.clock;
                        unsigned sig;
                        if(!lru_lookup(cache, S, k, now, W, sig)) {
                            OI_COUNT(2);
                            solve_selected<SMODE, false>(P, S, lut, k, W);
                            lru_store(cache, k, now, W, sig);
                        }
                    }
                    T0 = sel0;
                    T1 = sel1;
                }
                else if(need_var) {
                    __syncwarp();                                             // c_rho of this point
                    W.avar = variance_factor(S, k, W);
                }
                __syncwarp();
                finish_point(P, S, g, bg, k, W);
                __syncwarp();
            }
        }
    }
    }
    // the last warp to run out of work leaves the workspace ready for the next launch (no memset between launches)
    int last = 0;
    if(lane == 0) {
        __threadfence();
        last = atomicAdd(&hdr->warps_done, 1) == (int) (gridDim.x * WARPS_PER_CTA) - 1;
    }
    last = __shfl_sync(FULL, last, 0);
    if(last) {
        for(int u = lane; u < n_multi - split_from; u += 32) prog[u] = 0u;
        __syncwarp();
        if(lane == 0) {
            hdr->next_unit = 0;
            hdr->warps_done = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------ Cholesky path: 30 < k <= 64 --------
// Symmetric structure functions with more than FAST_K and at most CHOL_K observations per point: one warp per grid
// point, the selection from gather_candidates (oi.cpp:229-273), P + R assembled as a packed lower triangle in shared
// memory (fp64, k (k + 1) / 2 entries, oi.cpp:298-314) and factorised in place, A = L L'. With y = L^-1 d and
// u = L^-1 rho the increment is u . y and rho' A^-1 rho is u . u (oi.cpp:315-317,336): no inverse, no back-substitution.
// (The general kernel keeps its k x (k + 2) matrix in global memory and eliminates with partial pivoting; it is the
// fallback for non-symmetric structure functions and larger k, two orders of magnitude slower.)
constexpr int CHOL_K = 128;            // largest k of the Cholesky path (instantiated for 64 and 128)
constexpr int CHOL_WARPS = 2;
// Per-warp working set, carved from dynamic shared memory for the ACTUAL bound kcap on the observations per point (a fixed
// KMAX = 64 layout is 23.7 KB per warp = 9 warps per SM; max_points 50 needs 14 KB = 16 warps, 40 needs 10 KB).
template <int KMAX>
struct CholSmem {
    static constexpr int NSLOT = KMAX / 32 + 1;
    double* A;                             // packed lower triangle, row-major: (i, m <= i) at i (i + 1) / 2 + m
    double *d, *r;                         // innovations -> y -> z = A^-1 d (kept for reuse); rho -> u
    unsigned long long* key;               // [32 NSLOT]
    int* pos;                              // [32 NSLOT]
    int *c_pos, *c_orig;                   // the selection in canonical order (ascending original index)
    int* prev_orig;                        // the set whose z is in d
    float *sx, *sy, *sz, *selev, *slaf, *sratio;
    __host__ __device__ static size_t bytes(int kcap) {
        size_t b = sizeof(double) * ((size_t) kcap * (kcap + 1) / 2 + 2 * (size_t) kcap);
        b += sizeof(unsigned long long) * 32 * NSLOT;
        b += sizeof(int) * (32 * NSLOT + 3 * (size_t) kcap);
        b += sizeof(float) * 6 * (size_t) kcap;
        return (b + 15) / 16 * 16;
    }
    __device__ __forceinline__ void bind(unsigned char* base, int kcap) {
        A = reinterpret_cast<double*>(base);
        d = A + kcap * (kcap + 1) / 2;
        r = d + kcap;
        key = reinterpret_cast<unsigned long long*>(r + kcap);
        pos = reinterpret_cast<int*>(key + 32 * NSLOT);
        c_pos = pos + 32 * NSLOT;
        c_orig = c_pos + kcap;
        prev_orig = c_orig + kcap;
        sx = reinterpret_cast<float*>(prev_orig + kcap);
        sy = sx + kcap; sz = sy + kcap; selev = sz + kcap; slaf = selev + kcap; sratio = slaf + kcap;
    }
};

// A warp's cache of solved systems for the Cholesky path (global memory, L2-resident): the set in canonical order, z = (P+R)^-1 d
// and the innovation extremes -- the register path's LruBlock for up to KMAX observations.
constexpr int CHOL_LRU = 8;
template <int KMAX>
struct CholLru {
    double z[CHOL_LRU][KMAX];
    int orig[CHOL_LRU][KMAX];
    double dmax[CHOL_LRU], dmin[CHOL_LRU];
    unsigned sig[CHOL_LRU], stamp[CHOL_LRU];   // stamp: last use (0 = empty)
    int k[CHOL_LRU];
};

// Points are taken 32 at a time: on whole rows of a known grid as a tile of 4 rows x 8 columns walked in serpentine order (the
// region over which one set is selected is ~30 points across, so a compact block cuts fewer of them than a strip of 32, and
// the cache above then sees a set again on the next row), otherwise 32 consecutive points.
template <int SMODE, int KMAX>
__global__ void __launch_bounds__(CHOL_WARPS * 32) oi_chol_kernel(const __grid_constant__ OiParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef CholSmem<KMAX> Smem;
    constexpr int NSLOT = Smem::NSLOT;
    const int kcap = P.k, CHOL_PAIRS = kcap * (kcap + 1) / 2;
    Smem S;
    S.bind(smem_raw + (size_t) (threadIdx.x >> 5) * Smem::bytes(kcap), kcap);
    unsigned short* lut = reinterpret_cast<unsigned short*>(smem_raw + Smem::bytes(kcap) * CHOL_WARPS);   // pair -> (i << 8) | m
    const int lane = (int) lane_id();
    const bool need_var = P.analysis_variance != nullptr;
    for(int p = threadIdx.x; p < CHOL_PAIRS; p += blockDim.x) {
        int i = (int) ((sqrtf(8.f * (float) p + 1.f) - 1.f) * 0.5f);
        while(i * (i + 1) / 2 > p) i--;
        while((i + 1) * (i + 2) / 2 <= p) i++;
        lut[p] = (unsigned short) ((i << 8) | (p - i * (i + 1) / 2));
    }
    __syncthreads();
    const CandBuf cb = {S.key, S.pos};
    // the system solved last: neighbouring points usually select the same observations, and with the set in canonical
    // order z = (P + R)^-1 d depends on the set alone (see oi_fast_kernel); the increment is then rho . z
    int prev_k = -1;
    double prev_dmax = 0.0, prev_dmin = 0.0;
    CholLru<KMAX>* cache = reinterpret_cast<CholLru<KMAX>*>(P.workspace) + (blockIdx.x * CHOL_WARPS + (threadIdx.x >> 5));
    if(lane < CHOL_LRU) cache->stamp[lane] = 0;
    __syncwarp();
    unsigned now = 0;
    const int tiles_x = P.tile_nx > 0 ? (P.tile_nx + 7) / 8 : 0;
    const int n_rows = P.tile_nx > 0 ? P.count / P.tile_nx : 0;
    const int n_blocks = P.tile_nx > 0 ? tiles_x * ((n_rows + 3) / 4) : (P.count + 31) / 32;
    for(;;) {
        int blk = 0;
        if(lane == 0) blk = atomicAdd(P.work_counter, 1);
        blk = __shfl_sync(0xffffffffu, blk, 0);
        if(blk >= n_blocks) break;
        for(int j = 0; j < 32; j++) {
            int it;
            if(P.tile_nx > 0) {
                const int ty = blk / tiles_x, tx = blk - ty * tiles_x, r = j >> 3, c = (r & 1) ? 7 - (j & 7) : (j & 7);
                const int row = 4 * ty + r, col = 8 * tx + c;
                if(row >= n_rows || col >= P.tile_nx) continue;
                it = row * P.tile_nx + col;
            }
            else {
                it = blk * 32 + j;
                if(it >= P.count) break;
            }
            const int g = P.first + it;
            const float bg = P.background[g];
            int k = 0;
            if(is_valid(bg)) {   // oi.cpp:223
                const Pt p1 = bg_point(P, g);
                k = gather_candidates<SMODE, NSLOT>(P.obs, P.s, p1, P.R, P.k, cb);
            }
            if(k == 0) { keep_background(P, g, bg); continue; }
            // ---- canonical order: ascending original index = descending low word of the key
            {
                unsigned long long kk[KMAX / 32];
                int pp[KMAX / 32], rk[KMAX / 32];
                #pragma unroll
                for(int t = 0; t < KMAX / 32; t++) {
                    const bool has = lane + 32 * t < k;
                    kk[t] = has ? S.key[lane + 32 * t] : 0ull;
                    pp[t] = has ? S.pos[lane + 32 * t] : 0;
                    rk[t] = 0;
                }
                const unsigned* lo_words = reinterpret_cast<const unsigned*>(S.key);
                for(int j = 0; j < k; j++) {
                    const unsigned lw = lo_words[2 * j];
                    #pragma unroll
                    for(int t = 0; t < KMAX / 32; t++) rk[t] += lw > (unsigned) kk[t];
                }
                #pragma unroll
                for(int t = 0; t < KMAX / 32; t++)
                    if(lane + 32 * t < k) { S.c_pos[rk[t]] = pp[t]; S.c_orig[rk[t]] = cand_key_orig(kk[t]); S.r[rk[t]] = (double) cand_key_rho(kk[t]); }
                __syncwarp();
            }
            bool same = k == prev_k && !need_var;
            if(same) {
                bool eq = true;
                for(int i = lane; i < k; i += 32) eq = eq && S.c_orig[i] == S.prev_orig[i];
                same = __all_sync(0xffffffffu, eq);
            }
            double avar = 0.0;
            unsigned sig = 0;
            if(!same && !need_var) {
                // ---- the warp's cache of solved systems: one word per entry first, the member list only when that matches
                now++;
                unsigned h = 0;
                for(int i = lane; i < k; i += 32) h += (unsigned) (S.c_orig[i] + 1) * 0x9E3779B1u;
                sig = __reduce_add_sync(0xffffffffu, h);
                unsigned m = __ballot_sync(0xffffffffu, lane < CHOL_LRU && cache->stamp[lane] != 0 && cache->sig[lane] == sig && cache->k[lane] == k);
                while(m && !same) {
                    const int e = __ffs(m) - 1;
                    m &= m - 1;
                    bool eq = true;
                    for(int i = lane; i < k; i += 32) eq = eq && cache->orig[e][i] == S.c_orig[i];
                    if(__all_sync(0xffffffffu, eq)) {
                        for(int i = lane; i < k; i += 32) { S.d[i] = cache->z[e][i]; S.prev_orig[i] = S.c_orig[i]; }
                        prev_dmax = cache->dmax[e]; prev_dmin = cache->dmin[e]; prev_k = k;
                        if(lane == 0) cache->stamp[e] = now;
                        __syncwarp();
                        same = true;
                    }
                }
            }
            if(!same) {
                // ---- stage the selection
                double dmax = -INFINITY, dmin = INFINITY;
                for(int i = lane; i < k; i += 32) {
                    const int pos = S.c_pos[i];
                    S.sx[i] = P.obs.x[pos]; S.sy[i] = P.obs.y[pos]; S.sz[i] = P.obs.z[pos];
                    S.selev[i] = P.obs.elev[pos]; S.slaf[i] = P.obs.laf[pos]; S.sratio[i] = P.obs.ratio[pos];
                    const double dd = P.obs.innov[pos];
                    S.d[i] = dd;
                    S.prev_orig[i] = S.c_orig[i];
                    dmax = fmax(dmax, dd);
                    dmin = fmin(dmin, dd);
                }
                #pragma unroll
                for(int off = 16; off > 0; off >>= 1) {
                    dmax = fmax(dmax, shfl_double(dmax, lane ^ off));
                    dmin = fmin(dmin, shfl_double(dmin, lane ^ off));
                }
                prev_dmax = dmax; prev_dmin = dmin; prev_k = k;
                __syncwarp();
                // ---- P + R, lower triangle (oi.cpp:298-315)
                const int npairs = k * (k + 1) / 2;
                for(int p = lane; p < npairs; p += 32) {
                    const int code = lut[p], i = code >> 8, m = code & 255;
                    const Pt a = {S.sx[i], S.sy[i], S.sz[i], S.selev[i], S.slaf[i]};
                    const Pt b = {S.sx[m], S.sy[m], S.sz[m], S.selev[m], S.slaf[m]};
                    const float hdist = straight_distance(a.x, a.y, a.z, b.x, b.y, b.z);
                    double v = (double) corr_call<SMODE>(P.s, a, b, hdist);
                    if(i == m) v = __dadd_rn(v, (double) S.sratio[i]);
                    S.A[p] = v;
                }
                __syncwarp();
                // ---- left-looking Cholesky: column j of L from the columns before it; lanes own rows
                bool ok = true;
                for(int j = 0; j < k && ok; j++) {
                    const double* rowj = S.A + j * (j + 1) / 2;
                    double inv = 0.0;
                    for(int i0 = j; i0 < k; i0 += 32) {
                        const int i = i0 + lane;
                        double acc = 0.0;
                        double* rowi = S.A + (i < k ? i * (i + 1) / 2 : 0);
                        if(i < k) {
                            acc = rowi[j];
                            for(int p = 0; p < j; p++) acc = fma(-rowi[p], rowj[p], acc);
                        }
                        if(i0 == j) {   // the diagonal entry sits in lane 0 of the first chunk: L_jj = a_jj / sqrt(a_jj)
                            const double ajj = shfl_double(acc, 0);
                            ok = ajj > 0.0 && ajj < INFINITY;
                            inv = ok ? fast_rsqrt(ajj) : 0.0;
                        }
                        __syncwarp();
                        if(i < k) rowi[j] = acc * inv;
                    }
                    __syncwarp();
                }
                if(!ok) { prev_k = -1; keep_background(P, g, bg); continue; }   // not positive definite (arma::inv would throw)
                // ---- y = L^-1 d (and u = L^-1 rho when the variance is wanted): forward substitution
                for(int j = 0; j < k; j++) {
                    const double inv = fast_rcp(S.A[j * (j + 1) / 2 + j]);
                    const double yj = S.d[j] * inv, uj = need_var ? S.r[j] * inv : 0.0;
                    __syncwarp();
                    if(lane == 0) { S.d[j] = yj; if(need_var) S.r[j] = uj; }
                    for(int i = j + 1 + lane; i < k; i += 32) {
                        const double lij = S.A[i * (i + 1) / 2 + j];
                        S.d[i] = fma(-lij, yj, S.d[i]);
                        if(need_var) S.r[i] = fma(-lij, uj, S.r[i]);
                    }
                    __syncwarp();
                }
                if(need_var) {
                    // rho' A^-1 rho = u . u and the increment u . y (every point is solved when the variance is wanted)
                    double dx = 0.0, aa = 0.0;
                    for(int i = lane; i < k; i += 32) {
                        dx = fma(S.r[i], S.d[i], dx);
                        aa = fma(S.r[i], S.r[i], aa);
                    }
                    #pragma unroll
                    for(int off = 16; off > 0; off >>= 1) {
                        dx += shfl_double(dx, lane ^ off);
                        aa += shfl_double(aa, lane ^ off);
                    }
                    prev_k = -1;   // d holds y, not z
                    if(lane == 0) write_result(P, g, bg, dx, aa, dmax, dmin);
                    __syncwarp();
                    continue;
                }
                // ---- z = L'^-1 y: back substitution, in place
                for(int j = k - 1; j >= 0; j--) {
                    const double* rowj = S.A + j * (j + 1) / 2;
                    const double zj = S.d[j] * fast_rcp(rowj[j]);
                    __syncwarp();
                    if(lane == 0) S.d[j] = zj;
                    for(int i = lane; i < j; i += 32) S.d[i] = fma(-rowj[i], zj, S.d[i]);
                    __syncwarp();
                }
                {   // keep the solution: the least recently used entry goes (empty ones first)
                    const unsigned age = lane < CHOL_LRU ? (cache->stamp[lane] << 3) | (unsigned) lane : 0xffffffffu;
                    const int victim = (int) (__reduce_min_sync(0xffffffffu, age) & 7u);
                    for(int i = lane; i < k; i += 32) { cache->z[victim][i] = S.d[i]; cache->orig[victim][i] = S.c_orig[i]; }
                    if(lane == 0) {
                        cache->dmax[victim] = prev_dmax; cache->dmin[victim] = prev_dmin; cache->k[victim] = k;
                        cache->sig[victim] = sig; cache->stamp[victim] = now;
                    }
                    __syncwarp();
                }
            }
            // ---- increment = rho . z, the same dot product whether the system was solved here or reused
            double dx = 0.0;
            for(int i = lane; i < k; i += 32) dx = fma(S.r[i], S.d[i], dx);
            #pragma unroll
            for(int off = 16; off > 0; off >>= 1) dx += shfl_double(dx, lane ^ off);
            if(lane == 0) write_result(P, g, bg, dx, avar, prev_dmax, prev_dmin);
            __syncwarp();
        }
    }
}

// <Family>Structure::localization_distance(h) for a spatially varying scale (structure.cpp:280-282, 454-459, 603-609,
// 755-757, 902-904): a constant of min_rho times h, evaluated with the reference's float / double mix
__device__ __forceinline__ float spatial_loc_dist(const OiParams& P, float h) {
    return P.s.term[0].type == GPP_STRUCT_TOAR ? (float) __dmul_rn(P.loc_d, (double) h) : __fmul_rn(P.loc_c, h);
}

// ------------------------------------------------------------------ general path ----------------------
// Per-warp scratch in global memory: two candidate buffers of (kcap + 32) entries and a kcap x (kcap + 2)
// row-major fp64 matrix [P+R | d | rho].
struct GeneralScratch {
    float* rho[2];
    int* pos[2];
    int* orig[2];
    double* M;
};
__device__ __forceinline__ GeneralScratch scratch_for(unsigned char* base, size_t bytes_per_warp, int warp, int kcap) {
    unsigned char* p = base + (size_t) warp * bytes_per_warp;
    GeneralScratch s;
    s.M = reinterpret_cast<double*>(p);
    p += sizeof(double) * (size_t) kcap * (kcap + 3);
    int cap = kcap + 32;
    for(int b = 0; b < 2; b++) {
        s.rho[b] = reinterpret_cast<float*>(p); p += sizeof(float) * cap;
        s.pos[b] = reinterpret_cast<int*>(p); p += sizeof(int) * cap;
        s.orig[b] = reinterpret_cast<int*>(p); p += sizeof(int) * cap;
    }
    return s;
}
size_t scratch_bytes(int kcap) {
    size_t b = sizeof(double) * (size_t) kcap * (kcap + 3) + 2 * 3 * sizeof(float) * (size_t) (kcap + 32);
    return (b + 255) / 256 * 256;
}

// rank-select the best min(n, k) of src into dst (best-first); any n
__device__ int general_prune(const GeneralScratch& s, int src, int n, int k) {
    unsigned lane = lane_id();
    int dst = src ^ 1;
    for(int e = (int) lane; e < n; e += 32) {
        float r = s.rho[src][e];
        int o = s.orig[src][e];
        int rank = 0;
        for(int j = 0; j < n; j++) rank += cand_better(s.rho[src][j], s.orig[src][j], r, o) ? 1 : 0;
        if(rank < k) { s.rho[dst][rank] = r; s.pos[dst][rank] = s.pos[src][e]; s.orig[dst][rank] = o; }
    }
    __syncwarp();
    return min(n, k);
}

// smem_matrix_bytes > 0: the k x (k + 2) matrix of each warp lives in dynamic shared memory (k <= 64) instead of its
// global scratch slab; the elimination is a chain of dependent row operations, so the memory latency is what it costs.
template <bool IN_SMEM>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) oi_general_kernel(const __grid_constant__ OiParams P, unsigned char* scratch,
                                                                       size_t bytes_per_warp, int kcap) {
    extern __shared__ __align__(16) unsigned char general_smem[];
    const unsigned lane = lane_id();
    const int warps_per_cta = blockDim.x >> 5;
    const int warp_global = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * warps_per_cta;
    // IN_SMEM (k <= 64): the warp's whole scratch -- matrix and candidate buffers -- is shared memory, and the compiler sees it
    // (shared-space loads and 32-bit addresses instead of generic ones)
    const GeneralScratch S = IN_SMEM ? scratch_for(general_smem, bytes_per_warp, (int) (threadIdx.x >> 5), kcap)
                                     : scratch_for(scratch, bytes_per_warp, warp_global, kcap);
    const ObsView& obs = P.obs;

    for(int it = warp_global; it < P.count; it += warps_total) {
        const int g = P.first + it;
        const float bg = P.background[g];
        bool done = !is_valid(bg);
        int k = 0, cur = 0;
        Pt p1 = {0, 0, 0, 0, 0};
        if(!done) {
            p1 = bg_point(P, g);
            // the structure function as this point sees it (structure.cpp:189-199: the scales of the nearest node)
            gpp_structure sp = P.s;
            float R = P.R;
            if(P.sbh) {
                R = spatial_loc_dist(P, P.sbh[g]);
                sp.term[0].h = P.sbh[g]; sp.term[0].v = P.sbv[g]; sp.term[0].w = P.sbw[g]; sp.term[0].loc_dist = R;
            }
            float lo0 = __fsub_rn(p1.x, R), lo1 = __fsub_rn(p1.y, R), lo2 = __fsub_rn(p1.z, R);
            float hi0 = __fadd_rn(p1.x, R), hi1 = __fadd_rn(p1.y, R), hi2 = __fadd_rn(p1.z, R);
            int n = 0;
            if(lo0 < hi0 && lo1 < hi1 && lo2 < hi2) {
                int cx0 = cell_coord(obs.geom, 0, lo0), cx1 = cell_coord(obs.geom, 0, hi0);
                int cy0 = cell_coord(obs.geom, 1, lo1), cy1 = cell_coord(obs.geom, 1, hi1);
                int cz0 = cell_coord(obs.geom, 2, lo2), cz1 = cell_coord(obs.geom, 2, hi2);
                for(int cz = cz0; cz <= cz1; cz++)
                    for(int cy = cy0; cy <= cy1; cy++) {
                        int base = (cz * obs.geom.n[1] + cy) * obs.geom.n[0];
                        int s0 = obs.cell_start[base + cx0], s1 = obs.cell_start[base + cx1 + 1];
                        for(int chunk = s0; chunk < s1; chunk += 32) {
                            int i = chunk + (int) lane;
                            float rho = 0.f;
                            bool ok = i < s1;
                            if(ok) {
                                float ox = obs.x[i], oy = obs.y[i], oz = obs.z[i];
                                ok = ox > lo0 && ox < hi0 && oy > lo1 && oy < hi1 && oz > lo2 && oz < hi2;
                                if(ok) {
                                    float dist = straight_distance(ox, oy, oz, p1.x, p1.y, p1.z);
                                    ok = dist <= R;
                                    if(ok) {
                                        Pt p2 = {ox, oy, oz, obs.elev[i], obs.laf[i]};
                                        rho = structure_corr_background(sp, p1, p2, dist);
                                        ok = rho > 0.f;
                                    }
                                }
                            }
                            unsigned mask = __ballot_sync(0xffffffffu, ok);
                            if(ok) {
                                int slot = n + __popc(mask & ((1u << lane) - 1u));
                                S.rho[cur][slot] = rho; S.pos[cur][slot] = i; S.orig[cur][slot] = obs.orig[i];
                            }
                            n += __popc(mask);
                            __syncwarp();
                            if(n > kcap) { n = general_prune(S, cur, n, P.k); cur ^= 1; }
                        }
                    }
            }
            if(n > 0) { n = general_prune(S, cur, n, P.k); cur ^= 1; }
            k = n;
            done = k == 0;
        }
        if(done) {
            if(lane == 0) {
                P.analysis[g] = bg;
                if(P.analysis_variance) P.analysis_variance[g] = P.bvariance ? P.bvariance[g] : 1.f;
            }
            continue;
        }
        // ---- assemble [P+R | d | rho], k x (k+2), row-major; lP(i,j) = corr(p_i, p_j) (oi.cpp:305-313)
        const int ld = (k + 2) | 1;   // odd: rows a lane apart fall into different banks
        for(int e = (int) lane; e < k * k; e += 32) {
            int i = e / k, j = e % k;
            int pi = S.pos[cur][i], pj = S.pos[cur][j];
            Pt a = {obs.x[pi], obs.y[pi], obs.z[pi], obs.elev[pi], obs.laf[pi]};
            Pt b = {obs.x[pj], obs.y[pj], obs.z[pj], obs.elev[pj], obs.laf[pj]};
            double v;
            if(obs.sh) {   // corr(obs_i, obs_j) uses the scales at obs_i (structure.cpp:185-213): P is not symmetric
                gpp_structure si = P.s;
                si.term[0].h = obs.sh[pi]; si.term[0].v = obs.sv[pi]; si.term[0].w = obs.sw[pi];
                si.term[0].loc_dist = spatial_loc_dist(P, obs.sh[pi]);
                v = (double) structure_corr(si, a, b);
            }
            else v = (double) structure_corr(P.s, a, b);
            if(i == j) v = __dadd_rn(v, (double) obs.ratio[pi]);
            S.M[(size_t) i * ld + j] = v;
        }
        double dmax = -INFINITY, dmin = INFINITY;
        for(int i = (int) lane; i < k; i += 32) {
            double d = obs.innov[S.pos[cur][i]];
            S.M[(size_t) i * ld + k] = d;
            S.M[(size_t) i * ld + k + 1] = (double) S.rho[cur][i];
            dmax = fmax(dmax, d);
            dmin = fmin(dmin, d);
        }
        #pragma unroll
        for(int off = 16; off > 0; off >>= 1) {
            dmax = fmax(dmax, shfl_double(dmax, lane ^ off));
            dmin = fmin(dmin, shfl_double(dmin, lane ^ off));
        }
        __syncwarp();
        // ---- [A^-1 d | A^-1 rho]: Gaussian elimination with partial pivoting and two right-hand sides (oi.cuh ge_solve) for up to
        // 128 observations; beyond that (no register-resident pivot row) Gauss-Jordan, one row operation at a time
        bool singular = false;
        if(k + 2 <= 32) singular = !ge_solve<1>(S.M, k, ld, 2, (int) lane);          // (the pivot row's register chunks: 32 columns each)
        else if(k + 2 <= 64) singular = !ge_solve<2>(S.M, k, ld, 2, (int) lane);
        else if(k + 2 <= 96) singular = !ge_solve<3>(S.M, k, ld, 2, (int) lane);
        else if(k <= 128) singular = !ge_solve<(128 + 2 + 31) / 32>(S.M, k, ld, 2, (int) lane);
        else
        for(int c = 0; c < k; c++) {
            double best = -1.0;
            int brow = c;
            for(int r = c + (int) lane; r < k; r += 32) {
                double v = fabs(S.M[(size_t) r * ld + c]);
                if(v > best) { best = v; brow = r; }
            }
            #pragma unroll
            for(int off = 16; off > 0; off >>= 1) {
                double ob = shfl_double(best, lane ^ off);
                int orow = __shfl_xor_sync(0xffffffffu, brow, off);
                if(ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
            }
            if(!(best > 0.0)) { singular = true; break; }
            if(brow != c)
                for(int j = (int) lane; j < ld; j += 32) {
                    double t = S.M[(size_t) c * ld + j];
                    S.M[(size_t) c * ld + j] = S.M[(size_t) brow * ld + j];
                    S.M[(size_t) brow * ld + j] = t;
                }
            __syncwarp();
            const double inv = 1.0 / S.M[(size_t) c * ld + c];
            __syncwarp();
            for(int j = (int) lane; j < ld; j += 32) S.M[(size_t) c * ld + j] *= inv;
            __syncwarp();
            for(int r = 0; r < k; r++) {
                if(r == c) continue;
                const double f = S.M[(size_t) r * ld + c];
                __syncwarp();
                if(f != 0.0)
                    for(int j = (int) lane; j < ld; j += 32) S.M[(size_t) r * ld + j] = fma(-f, S.M[(size_t) c * ld + j], S.M[(size_t) r * ld + j]);
                __syncwarp();
            }
        }
        if(singular) {   // arma::inv would throw; leave the background
            if(lane == 0) {
                P.analysis[g] = bg;
                if(P.analysis_variance) P.analysis_variance[g] = P.bvariance ? P.bvariance[g] : 1.f;
            }
            continue;
        }
        // dx = rho . (A^-1 d), a = rho . (A^-1 rho)   (oi.cpp:315-316,336)
        double dx = 0.0, aa = 0.0;
        for(int i = (int) lane; i < k; i += 32) {
            double rho = (double) S.rho[cur][i];
            dx = fma(rho, S.M[(size_t) i * ld + k], dx);
            aa = fma(rho, S.M[(size_t) i * ld + k + 1], aa);
        }
        #pragma unroll
        for(int off = 16; off > 0; off >>= 1) {
            dx += shfl_double(dx, lane ^ off);
            aa += shfl_double(aa, lane ^ off);
        }
        if(lane == 0) write_result(P, g, bg, dx, aa, dmax, dmin);
        __syncwarp();
    }
}

// Largest number of observations inside the localization radius of any background point; sizes the general
// path when max_points is unlimited. One thread per background point.
__global__ void oi_count_kernel(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gz,
                                const float* __restrict__ background, int first, int count, ObsView obs, float R,
                                int* __restrict__ out_max) {
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    int n = 0;
    if(it < count && (!background || is_valid(background[first + it]))) {
        int g = first + it;
        float x = gx[g], y = gy[g], z = gz[g];
        float lo0 = __fsub_rn(x, R), lo1 = __fsub_rn(y, R), lo2 = __fsub_rn(z, R);
        float hi0 = __fadd_rn(x, R), hi1 = __fadd_rn(y, R), hi2 = __fadd_rn(z, R);
        if(lo0 < hi0 && lo1 < hi1 && lo2 < hi2) {
            int cx0 = cell_coord(obs.geom, 0, lo0), cx1 = cell_coord(obs.geom, 0, hi0);
            int cy0 = cell_coord(obs.geom, 1, lo1), cy1 = cell_coord(obs.geom, 1, hi1);
            int cz0 = cell_coord(obs.geom, 2, lo2), cz1 = cell_coord(obs.geom, 2, hi2);
            for(int cz = cz0; cz <= cz1; cz++)
                for(int cy = cy0; cy <= cy1; cy++) {
                    int base = (cz * obs.geom.n[1] + cy) * obs.geom.n[0];
                    for(int i = obs.cell_start[base + cx0]; i < obs.cell_start[base + cx1 + 1]; i++) {
                        float ox = obs.x[i], oy = obs.y[i], oz = obs.z[i];
                        if(ox > lo0 && ox < hi0 && oy > lo1 && oy < hi1 && oz > lo2 && oz < hi2 &&
                           straight_distance(ox, oy, oz, x, y, z) <= R)
                            n++;
                    }
                }
        }
    }
    #pragma unroll
    for(int off = 16; off > 0; off >>= 1) n = max(n, __shfl_xor_sync(0xffffffffu, n, off));
    if(lane_id() == 0 && n > 0) atomicMax(out_max, n);
}

__global__ void copy_background_kernel(const float* __restrict__ background, const float* __restrict__ bvariance, int first,
                                       int count, float* __restrict__ analysis, float* __restrict__ analysis_variance) {
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    if(it >= count) return;
    analysis[first + it] = background[first + it];
    if(analysis_variance) analysis_variance[first + it] = bvariance ? bvariance[first + it] : 1.f;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
int gpp::build_obs_table(const gpp_points* op, const std::vector<char>& valid, const std::vector<double>& innov,
                         const std::vector<float>& ratio, float loc_dist, gpp_oi_obs* out, std::vector<int>* order,
                         const std::vector<float>* scales) {
    const int nS = op->n;
    out->n_total = nS;
    out->loc_dist = loc_dist;
    // bounding box of the valid observations
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const std::vector<float>* co[3] = {&op->x, &op->y, &op->z};
    int nv = 0;
    for(int i = 0; i < nS; i++) {
        if(!valid[i]) continue;
        nv++;
        for(int d = 0; d < 3; d++) {
            lo[d] = std::min(lo[d], (*co[d])[i]);
            hi[d] = std::max(hi[d], (*co[d])[i]);
        }
    }
    out->n_valid = nv;
    CellGeom& g = out->geom;
    // cell edge = R/2: the box of side 2R then spans at most 5 cells per dimension
    double edge = (is_valid(loc_dist) && loc_dist > 0) ? 0.5 * (double) loc_dist : 0.0;
    long long total = 1;
    double ext[3];
    for(int d = 0; d < 3; d++) {
        ext[d] = nv > 0 ? (double) hi[d] - (double) lo[d] : 0.0;
        int c = 1;
        if(edge > 0 && ext[d] > 0) c = (int) std::min(2048.0, std::max(1.0, std::ceil(ext[d] / edge)));
        g.n[d] = c;
        total *= c;
    }
    while(total > (1LL << 22)) {
        int dmax = 0;
        for(int d = 1; d < 3; d++) if(g.n[d] > g.n[dmax]) dmax = d;
        total /= g.n[dmax];
        g.n[dmax] = (g.n[dmax] + 1) / 2;
        total *= g.n[dmax];
    }
    for(int d = 0; d < 3; d++) {
        g.lo[d] = nv > 0 ? lo[d] : 0.f;
        g.edge[d] = ext[d] > 0 ? (float) (ext[d] / g.n[d]) : 1.f;
        g.inv[d] = ext[d] > 0 ? (float) (g.n[d] / ext[d]) : 0.f;
    }
    out->ncells = (int) total;
    // stable counting sort by cell
    std::vector<int> start(total + 1, 0), cell(nS, -1);
    for(int i = 0; i < nS; i++) {
        if(!valid[i]) continue;
        int c = (cell_coord(g, 2, op->z[i]) * g.n[1] + cell_coord(g, 1, op->y[i])) * g.n[0] + cell_coord(g, 0, op->x[i]);
        cell[i] = c;
        start[c + 1]++;
    }
    for(long long c = 0; c < total; c++) start[c + 1] += start[c];
    std::vector<int> fill(start.begin(), start.end() - 1), orig(std::max(nv, 1));
    std::vector<float> sx(std::max(nv, 1)), sy(std::max(nv, 1)), sz(std::max(nv, 1)), se(std::max(nv, 1)), sl(std::max(nv, 1)),
        sr(std::max(nv, 1));
    std::vector<double> si(std::max(nv, 1));
    for(int i = 0; i < nS; i++) {
        if(!valid[i]) continue;
        int slot = fill[cell[i]]++;
        orig[slot] = i;
        sx[slot] = op->x[i]; sy[slot] = op->y[i]; sz[slot] = op->z[i];
        se[slot] = op->elevs[i]; sl[slot] = op->lafs[i];
        sr[slot] = ratio[i];
        si[slot] = innov[i];
    }
    if(order) order->assign(orig.begin(), orig.begin() + nv);
    if(scales) {   // h, v, w of every observation, in table order
        std::vector<float> t(std::max(nv, 1));
        gpp::DeviceBuffer<float>* dst[3] = {&out->sh, &out->sv, &out->sw};
        for(int c = 0; c < 3; c++) {
            for(int slot = 0; slot < nv; slot++) t[slot] = scales[c][orig[slot]];
            GPP_TRY(dst[c]->upload(t.data(), nv));
            GPP_CUDA(cudaStreamSynchronize(0));
        }
        out->has_scales = true;
    }
    GPP_TRY(out->cell_start.upload(start.data(), start.size()));
    GPP_TRY(out->orig.upload(orig.data(), nv));
    GPP_TRY(out->x.upload(sx.data(), nv));
    GPP_TRY(out->y.upload(sy.data(), nv));
    GPP_TRY(out->z.upload(sz.data(), nv));
    GPP_TRY(out->elev.upload(se.data(), nv));
    GPP_TRY(out->laf.upload(sl.data(), nv));
    GPP_TRY(out->ratio.upload(sr.data(), nv));
    GPP_TRY(out->innov.upload(si.data(), nv));
    GPP_CUDA(cudaStreamSynchronize(0));   // the host staging vectors go out of scope
    return GPP_OK;
}

int gpp::count_max_candidates(gpp_points* bp, int first, int count, const float* d_background, const ObsView& obs, float R,
                              cudaStream_t stream, int* out) {
    DeviceBuffer<int> dmax;
    GPP_TRY(dmax.alloc(1));
    GPP_CUDA(cudaMemsetAsync(dmax.ptr, 0, sizeof(int), stream));
    GPP_LAUNCH(oi_count_kernel, (unsigned) ((count + 255) / 256), 256, 0, stream, bp->dx.ptr, bp->dy.ptr, bp->dz.ptr, d_background, first,
               count, obs, R, dmax.ptr);
    GPP_CUDA(cudaMemcpyAsync(out, dmax.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    GPP_CUDA(cudaStreamSynchronize(stream));
    return GPP_OK;
}

namespace {
// How the range is cut into work units (UnitPlan): most of it in the largest units; the last rows in units a quarter the
// size, then a quarter of that, ... each level holding enough work (half a unit of the level above per resident warp) to
// even out the spread of finishing times the level above leaves behind. The tail of a launch is then one smallest unit
// (16 points) instead of one largest (256): what limited the strong scaling over 8 GPUs, where a rank has only ~3 large
// units per warp.
UnitPlan oi_plan_units(int tile_nx, int count, long long warps) {
    UnitPlan u;
    std::memset(&u, 0, sizeof(u));
    const bool tiles = tile_nx > 0;
    const long long width = tiles ? (tile_nx + 3) / 4 : 1;                                       // tiles per tile row
    long long rem = tiles ? (count / tile_nx + 3) / 4 : ((long long) count + RUN - 1) / RUN;   // tile rows, or runs
    long long rows[4] = {0, 0, 0, 0};
    auto unit_rows = [&](int l) { const int sh = TOP_SHIFT - l; return (long long) (tiles ? (1 << sh) : (1 << (2 * sh))); };
    // The largest unit in use. Level 0 always: large units solve the fewest systems, and the spread of their costs (several-fold,
    // with the local churn of the selection) is evened out by the warps that run out of units taking tiles from the far end of
    // the units still in progress (oi_fast_kernel). Measured on row blocks of C3 (profiles/oi_slices.py): starting from
    // 2 x 2-tile units instead costs 15 % more per point. GPP_OI_FIRST_LEVEL overrides it for experiments.
    int first = 0;
    if(const char* env = std::getenv("GPP_OI_FIRST_LEVEL")) first = std::max(0, std::min(N_LEVELS - 1, std::atoi(env)));
    for(int l = N_LEVELS - 1; l > first; l--) {
        const int sh = TOP_SHIFT - l;
        const long long area = (warps / 2 + 1) << (2 * (sh + 1));   // tiles (runs) wanted at this level
        long long want = (area + width - 1) / width;
        want = (want + unit_rows(l) - 1) / unit_rows(l) * unit_rows(l);
        long long r = std::min(rem, want);
        r -= r % unit_rows(l);
        rows[l] = r;
        rem -= r;
    }
    for(int l = first; l < N_LEVELS; l++) {   // the bulk at the first level; what does not fill a unit goes to the finer ones
        const long long r = rem - rem % unit_rows(l);
        rows[l] += r;
        rem -= r;
    }
    long long base = 0, end = 0;
    for(int l = 0; l < 4; l++) {
        if(l < N_LEVELS) {
            const int sh = TOP_SHIFT - l;
            const long long cols = tiles ? (width + (1 << sh) - 1) >> sh : 1;
            u.base[l] = (int) base;
            u.cols[l] = (int) cols;
            end += rows[l] / unit_rows(l) * cols;
            base += rows[l];
        }
        u.end[l] = (int) end;
    }
    u.n_units = (int) end;
    return u;
}
size_t oi_workspace_bytes() { return OI_WORK_LRU_OFFSET + sizeof(LruBlock) * (size_t) sm_count() * 2 * WARPS_PER_CTA; }

int check_structure(const gpp_structure* s) {
    if(!s) return fail(GPP_ERR_INVALID_ARGUMENT, "structure must not be NULL");
    if(s->n_terms != 1 && s->n_terms != 3) return fail(GPP_ERR_INVALID_ARGUMENT, "structure.n_terms must be 1 or 3");
    for(int t = 0; t < s->n_terms; t++)
        if(s->term[t].type < GPP_STRUCT_BARNES || s->term[t].type > GPP_STRUCT_LINEAR)
            return fail(GPP_ERR_INVALID_ARGUMENT, "unknown structure function type %d", s->term[t].type);
    return reject_unset_scales(s);
}
}  // namespace

namespace gpp {
// Pinned staging for pipelined downloads: two slots, grown on demand, kept for the life of the process.
struct PinnedStage {
    std::mutex lock;
    float* slot[2] = {nullptr, nullptr};
    size_t capacity = 0;   // floats per slot
    int reserve(size_t n) {
        if(n <= capacity) return GPP_OK;
        for(int i = 0; i < 2; i++) {
            if(slot[i]) cudaFreeHost(slot[i]);
            slot[i] = nullptr;
        }
        capacity = 0;
        for(int i = 0; i < 2; i++) GPP_CUDA(cudaHostAlloc((void**) &slot[i], n * sizeof(float), cudaHostAllocDefault));
        capacity = n;
        return GPP_OK;
    }
};
// one staging pair per device: host threads driving different devices (gpp_optimal_interpolation_multi_gpu_host) must not
// serialise on one lock
PinnedStage g_stages[64];
PinnedStage& stage_of_current_device() {
    int dev = 0;
    if(cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return g_stages[dev];
}

// Blocks [bounds[c], bounds[c+1]) (in floats) of d_out are produced back to back on the default stream by launch(c); a
// second stream copies each finished block into a pinned slot, and the host moves it into the caller's array while the
// next block is being computed. A D2H straight into a pageable array runs at ~5 GB/s, first-touch page faults included.
//
// overlap_blocks: launch(c, stream) is given two alternating streams (ordered after the default stream's earlier work), so
// that block c + 1 starts filling the multiprocessors while the last CTAs of block c drain. For kernels whose work items are
// long (EnSI: tens of milliseconds per item) the drain of every block is otherwise idle time.
int pipelined_download(const std::vector<size_t>& bounds, const std::function<int(int, cudaStream_t)>& launch, const float* d_out, float* host_out,
                       bool overlap_blocks) {
    const int n_chunks = (int) bounds.size() - 1;
    size_t largest = 0;
    for(int c = 0; c < n_chunks; c++) largest = std::max(largest, bounds[c + 1] - bounds[c]);
    PinnedStage& g_stage = stage_of_current_device();
    std::lock_guard<std::mutex> guard(g_stage.lock);
    GPP_TRY(g_stage.reserve(largest));
    // streams and events are released on every path out of this function
    struct Stream {
        cudaStream_t s = nullptr;
        ~Stream() { if(s) cudaStreamDestroy(s); }
        int create() { GPP_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return GPP_OK; }
    };
    struct Event {
        cudaEvent_t e = nullptr;
        ~Event() { if(e) cudaEventDestroy(e); }
        int create() { GPP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return GPP_OK; }
    };
    Stream copy, comp[2];
    GPP_TRY(copy.create());
    const cudaStream_t copy_stream = copy.s;
    std::vector<Event> produced(n_chunks), copied(n_chunks);
    for(int c = 0; c < n_chunks; c++) {
        GPP_TRY(produced[c].create());
        GPP_TRY(copied[c].create());
    }
    int rc = GPP_OK;
    cudaStream_t compute[2] = {0, 0};
    if(overlap_blocks) {
        Event ready;
        GPP_TRY(ready.create());
        GPP_CUDA(cudaEventRecord(ready.e, 0));
        for(int i = 0; i < 2; i++) {
            GPP_TRY(comp[i].create());
            compute[i] = comp[i].s;
            GPP_CUDA(cudaStreamWaitEvent(compute[i], ready.e, 0));
        }
    }
    for(int c = 0; c < n_chunks && rc == GPP_OK; c++) {
        rc = launch(c, compute[c & 1]);
        if(rc == GPP_OK && cudaEventRecord(produced[c].e, compute[c & 1]) != cudaSuccess) rc = fail(GPP_ERR_CUDA, "cudaEventRecord failed");
    }
    // The caller's array is usually fresh (untouched pages): fault it in now, while the device is busy with block 0,
    // instead of during the copies at the end (first-touch runs at 2-4 GB/s).
    if(rc == GPP_OK)
        for(size_t i = bounds[0]; i < bounds[n_chunks]; i += 1024) reinterpret_cast<volatile float*>(host_out)[i] = 0.f;
    auto finish = [&](int c) {   // block c: wait for its copy, move it to the caller's array
        if(cudaEventSynchronize(copied[c].e) != cudaSuccess) return fail(GPP_ERR_CUDA, "CUDA error while downloading a result block");
        // a few threads: one core copies ~8 GB/s, and the copy of the last block is not hidden behind any kernel
        const size_t n = bounds[c + 1] - bounds[c];
        const int pieces = n >= (1u << 20) ? 4 : 1;
        #pragma omp parallel for num_threads(pieces) schedule(static)
        for(int t = 0; t < pieces; t++) {
            const size_t a = n * t / pieces, b = n * (t + 1) / pieces;
            std::memcpy(host_out + bounds[c] + a, g_stage.slot[c & 1] + a, sizeof(float) * (b - a));
        }
        return (int) GPP_OK;
    };
    for(int c = 0; c < n_chunks && rc == GPP_OK; c++) {
        if(c >= 2) rc = finish(c - 2);               // frees slot c & 1
        if(rc != GPP_OK) break;
        if(cudaStreamWaitEvent(copy_stream, produced[c].e, 0) != cudaSuccess ||
           cudaMemcpyAsync(g_stage.slot[c & 1], d_out + bounds[c], sizeof(float) * (bounds[c + 1] - bounds[c]), cudaMemcpyDeviceToHost, copy_stream) != cudaSuccess ||
           cudaEventRecord(copied[c].e, copy_stream) != cudaSuccess)
            rc = fail(GPP_ERR_CUDA, "CUDA error while queueing the download of a result block");
    }
    for(int c = std::max(0, n_chunks - 2); c < n_chunks && rc == GPP_OK; c++) rc = finish(c);
    // nothing of this call may still be running when the streams go away
    cudaStreamSynchronize(copy_stream);
    if(overlap_blocks)
        for(int i = 0; i < 2; i++) cudaStreamSynchronize(compute[i]);
    if(rc == GPP_OK) {
        cudaError_t err = cudaGetLastError();
        if(err != cudaSuccess) rc = fail(GPP_ERR_CUDA, "CUDA error %s: %s", cudaGetErrorName(err), cudaGetErrorString(err));
    }
    return rc;
}
}  // namespace gpp

int oi_device_range(const gpp_points* cbp, int first, int count, const float* d_background, const float* d_bvariance, const gpp_oi_obs* obs,
                    const gpp_structure* structure, int max_points, int allow_extrapolation, float* d_analysis, float* d_analysis_variance,
                    void* d_workspace, size_t workspace_bytes, void* stream_, int kcap_given);

namespace {
// Row blocks of the grid go through upload, analysis and download as a pipeline: block c's rows of the background (and
// of its variance) are copied in on the stream that then analyses them, the two streams alternate, and the results
// return through the pinned staging buffer. A point reads no background value but its own, so a block needs no halo.
int analyse_pipelined(const gpp_points* bpoints, int nB, int nx, int n_chunks, const float* background, const float* bvariance, float* d_bg,
                      float* d_bvar, const gpp_oi_obs* obs, const gpp_structure* structure, int max_points, int allow_extrapolation, float* d_out,
                      float* d_var, float* analysis) {
    const int n_rows = nB / nx;
    std::vector<size_t> bounds(n_chunks + 1);
    for(int c = 0; c <= n_chunks; c++) bounds[c] = (size_t) ((long long) n_rows * c / n_chunks) * nx;
    // unlimited max_points: one bound on the observations per point for the whole field, so that every block takes the same kernel
    int kcap = 0;
    if(max_points == 0 && obs->n_valid > FAST_K) {
        gpp_points* bp = const_cast<gpp_points*>(bpoints);
        GPP_TRY(bp->ensure_on_device());
        int hmax = 0;
        GPP_TRY(count_max_candidates(bp, 0, nB, nullptr, obs->view(), structure->term[0].loc_dist, 0, &hmax));
        kcap = std::max(1, std::min(obs->n_valid, hmax));
    }
    return pipelined_download(bounds, [&](int c, cudaStream_t stream) {
        const size_t first = bounds[c], count = bounds[c + 1] - bounds[c];
        GPP_CUDA(cudaMemcpyAsync(d_bg + first, background + first, sizeof(float) * count, cudaMemcpyHostToDevice, stream));
        if(bvariance) GPP_CUDA(cudaMemcpyAsync(d_bvar + first, bvariance + first, sizeof(float) * count, cudaMemcpyHostToDevice, stream));
        return oi_device_range(bpoints, (int) first, (int) count, d_bg, bvariance ? d_bvar : nullptr, obs, structure, max_points,
                               allow_extrapolation, d_out, d_var, nullptr, 0, stream, kcap);
    }, d_out, analysis, true);
}
}  // namespace

namespace {
// The general kernel on the background range in P, with per-warp scratch for up to kcap observations per point.
int launch_general(const OiParams& P, int kcap, int count, cudaStream_t stream) {
    const int sms = sm_count();
    const size_t per_warp = scratch_bytes(kcap);
    // k <= 64: the matrix goes to shared memory, 2 warps per CTA (34 KB each at k = 64)
    const bool in_smem = kcap <= 64;
    const int warps_per_cta = in_smem ? 2 : WARPS_PER_CTA;
    const size_t smem = in_smem ? per_warp * warps_per_cta : 0;
    int per_sm = 2;
    if(in_smem) {
        GPP_CUDA(cudaFuncSetAttribute(oi_general_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        GPP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, oi_general_kernel<true>, warps_per_cta * 32, smem));
        per_sm = std::max(per_sm, 1);
    }
    long long warps = (long long) sms * per_sm * warps_per_cta;
    const size_t budget = (size_t) 4 << 30;
    if((size_t) warps * per_warp > budget) warps = std::max<long long>(warps_per_cta, (long long) (budget / per_warp) / warps_per_cta * warps_per_cta);
    if((size_t) warps * per_warp > ((size_t) 64 << 30))
        return fail(GPP_ERR_RUNTIME, "optimal_interpolation: %d observations per point need %zu bytes of scratch per warp", kcap, per_warp);
    unsigned grid = (unsigned) std::min<long long>(warps / warps_per_cta, ((long long) count + warps_per_cta - 1) / warps_per_cta);
    grid = std::max(grid, 1u);
    unsigned char* scratch = nullptr;
    if(in_smem) oi_general_kernel<true><<<grid, warps_per_cta * 32, smem, stream>>>(P, nullptr, per_warp, kcap);
    else {
        GPP_CUDA(cudaMallocAsync((void**) &scratch, (size_t) grid * warps_per_cta * per_warp, stream));
        oi_general_kernel<false><<<grid, warps_per_cta * 32, 0, stream>>>(P, scratch, per_warp, kcap);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t err = cudaGetLastError();
    if(scratch) cudaFreeAsync(scratch, stream);
    if(err != cudaSuccess) return fail(GPP_ERR_CUDA, "CUDA error %s launching oi_general_kernel: %s", cudaGetErrorName(err), cudaGetErrorString(err));
    return GPP_OK;
}
}  // namespace

extern "C" {

int gpp_oi_obs_create(const gpp_points* opoints, const float* pobs, const float* obs_variance, const float* pbackground,
                      const float* bvariance_at_points, const gpp_structure* structure, gpp_oi_obs** out) {
    if(!out) return fail(GPP_ERR_INVALID_ARGUMENT, "out must not be NULL");
    *out = nullptr;
    if(!opoints) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    GPP_TRY(check_structure(structure));
    GPP_TRY(ensure_device());
    const int nS = opoints->n;
    std::vector<char> valid(nS);
    std::vector<double> innov(nS);
    std::vector<float> ratio(nS);
    for(int i = 0; i < nS; i++) {
        // oi.cpp:252: observations with an invalid value or background never contribute
        valid[i] = is_valid(pobs[i]) && is_valid(pbackground[i]);
        innov[i] = (double) pobs[i] - (double) pbackground[i];                              // lObs - lY, oi.cpp:301-302,316
        ratio[i] = obs_variance[i] / (bvariance_at_points ? bvariance_at_points[i] : 1.f);  // oi.cpp:192-195
    }
    gpp_oi_obs* o = new(std::nothrow) gpp_oi_obs();
    if(!o) return fail(GPP_ERR_RUNTIME, "out of memory");
    int rc = build_obs_table(opoints, valid, innov, ratio, structure->term[0].loc_dist, o);
    // launch workspaces of the register path (work counter + per-warp caches of solved systems): OI_WORK_SLOTS of them, used
    // round-robin, so that a step is one kernel launch with no allocation or memset around it
    if(rc == GPP_OK) {
        o->work_slot_bytes = oi_workspace_bytes();
        rc = o->work.alloc(o->work_slot_bytes * OI_WORK_SLOTS);
        if(rc == GPP_OK && cudaMemsetAsync(o->work.ptr, 0, o->work_slot_bytes * OI_WORK_SLOTS, 0) != cudaSuccess) rc = fail(GPP_ERR_CUDA, "cudaMemsetAsync failed");
        if(rc == GPP_OK && cudaStreamSynchronize(0) != cudaSuccess) rc = fail(GPP_ERR_CUDA, "cudaStreamSynchronize failed");
    }
    if(rc != GPP_OK) { delete o; return rc; }
    *out = o;
    return GPP_OK;
}

void gpp_oi_obs_destroy(gpp_oi_obs* obs) { delete obs; }

size_t gpp_oi_workspace_bytes(void) {
    if(ensure_device() != GPP_OK) return 0;
    return oi_workspace_bytes();
}

int gpp_optimal_interpolation_device(const gpp_points* cbp, int first, int count, const float* d_background,
                                     const float* d_bvariance, const gpp_oi_obs* obs, const gpp_structure* structure,
                                     int max_points, int allow_extrapolation, float* d_analysis, float* d_analysis_variance,
                                     void* stream_) {
    return gpp_optimal_interpolation_device_ws(cbp, first, count, d_background, d_bvariance, obs, structure, max_points, allow_extrapolation,
                                               d_analysis, d_analysis_variance, nullptr, 0, stream_);
}

int gpp_optimal_interpolation_device_ws(const gpp_points* cbp, int first, int count, const float* d_background,
                                        const float* d_bvariance, const gpp_oi_obs* obs, const gpp_structure* structure,
                                        int max_points, int allow_extrapolation, float* d_analysis, float* d_analysis_variance,
                                        void* d_workspace, size_t workspace_bytes, void* stream_) {
    return oi_device_range(cbp, first, count, d_background, d_bvariance, obs, structure, max_points, allow_extrapolation, d_analysis,
                           d_analysis_variance, d_workspace, workspace_bytes, stream_, 0);
}

}  // extern "C"

// kcap_given > 0: the bound on the observations per point has been established for the WHOLE field by the caller (the
// pipelined host path analyses a field block by block: with max_points == 0 every block would otherwise count its own
// largest neighbourhood and could take a different kernel than its neighbours).
int oi_device_range(const gpp_points* cbp, int first, int count, const float* d_background, const float* d_bvariance, const gpp_oi_obs* obs,
                    const gpp_structure* structure, int max_points, int allow_extrapolation, float* d_analysis, float* d_analysis_variance,
                    void* d_workspace, size_t workspace_bytes, void* stream_, int kcap_given) {
    cudaStream_t stream = (cudaStream_t) stream_;
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");   // oi.cpp:152
    if(!cbp || !obs) return fail(GPP_ERR_INVALID_ARGUMENT, "points and observation state must not be NULL");
    GPP_TRY(check_structure(structure));
    gpp_points* bp = const_cast<gpp_points*>(cbp);
    if(first < 0 || count < 0 || first + count > bp->n) return fail(GPP_ERR_INVALID_ARGUMENT, "background range out of bounds");
    if(count == 0) return GPP_OK;
    GPP_TRY(bp->ensure_on_device());
    const unsigned blocks_copy = (unsigned) ((count + 255) / 256);
    if(obs->n_valid == 0) {   // oi.cpp:189-190 and :234-237: nothing to assimilate
        GPP_LAUNCH(copy_background_kernel, blocks_copy, 256, 0, stream, d_background, d_bvariance, first, count, d_analysis,
                   d_analysis_variance);
        return GPP_OK;
    }
    OiParams P;
    P.gx = bp->dx.ptr; P.gy = bp->dy.ptr;
    P.gz = bp->type == GPP_CARTESIAN ? nullptr : bp->dz.ptr;   // util.cpp:583-615: z = 0 for Cartesian points
    P.gelev = bp->has_elevs ? bp->delev.ptr : nullptr;           // points.cpp:23-30: NaN when not given
    P.glaf = bp->has_lafs ? bp->dlaf.ptr : nullptr;
    P.background = d_background;
    P.bvariance = d_bvariance;
    P.analysis = d_analysis;
    P.analysis_variance = d_analysis_variance;
    P.first = first;
    P.count = count;
    P.obs = obs->view();
    P.s = *structure;
    P.R = structure->term[0].loc_dist;
    P.allow_extrapolation = allow_extrapolation;
    P.tile_nx = 0;
    P.workspace = nullptr;
    std::memset(&P.plan, 0, sizeof(P.plan));
    P.work_counter = nullptr;
    P.sbh = P.sbv = P.sbw = nullptr;
    P.loc_c = 0.f;
    P.loc_d = 0.0;

    int kcap = max_points > 0 ? std::min(max_points, obs->n_valid) : obs->n_valid;
    if(kcap_given > 0) kcap = std::min(kcap, kcap_given);
    else if(max_points == 0 && kcap > FAST_K) {
        // unlimited: bound k by the largest neighbourhood actually present
        int hmax = 0;
        GPP_TRY(count_max_candidates(bp, first, count, d_background, P.obs, P.R, stream, &hmax));
        kcap = std::min(kcap, std::max(hmax, 1));
    }
    P.k = kcap;
    const int sms = sm_count();
    if(kcap <= FAST_K && structure_is_symmetric(*structure)) {
        const size_t smem = sizeof(FastSmem) * WARPS_PER_CTA + sizeof(unsigned short) * 512;
        static_assert(sizeof(FastSmem) * WARPS_PER_CTA + 1024 <= 113 * 1024, "two CTAs per SM");
        const int mode = structure_mode(*structure);
        GPP_CUDA(cudaFuncSetAttribute(oi_fast_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        GPP_CUDA(cudaFuncSetAttribute(oi_fast_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        // whole rows of a known grid -> 4 x 4 tiles
        P.tile_nx = (bp->shape_nx > 0 && first % bp->shape_nx == 0 && count % bp->shape_nx == 0) ? bp->shape_nx : 0;
        const long long max_warps = (long long) sms * 2 * WARPS_PER_CTA;   // 2 resident CTAs per SM
        P.plan = oi_plan_units(P.tile_nx, count, max_warps);
        const unsigned grid = (unsigned) std::max<long long>(1, std::min<long long>(((long long) P.plan.n_units + WARPS_PER_CTA - 1) / WARPS_PER_CTA, (long long) sms * 2));
        // the launch workspace: the caller's, else one of the slots the observation state owns, else a stream-ordered
        // allocation (observation states built on the stack by other entry points have no slots)
        const size_t need = OI_WORK_LRU_OFFSET + sizeof(LruBlock) * (size_t) grid * WARPS_PER_CTA;
        unsigned char* temp = nullptr;
        if(d_workspace) {
            if(workspace_bytes < need) return fail(GPP_ERR_INVALID_ARGUMENT, "workspace of %zu bytes given, %zu needed (gpp_oi_workspace_bytes)", workspace_bytes, need);
            P.workspace = static_cast<unsigned char*>(d_workspace);
        }
        else if(obs->work.ptr && obs->work_slot_bytes >= need)
            P.workspace = obs->work.ptr + obs->work_slot_bytes * (obs->work_next.fetch_add(1, std::memory_order_relaxed) % OI_WORK_SLOTS);
        else {
            GPP_CUDA(cudaMallocAsync((void**) &temp, need, stream));
            GPP_CUDA(cudaMemsetAsync(temp, 0, OI_WORK_LRU_OFFSET, stream));
            P.workspace = temp;
        }
        if(mode == 1) oi_fast_kernel<1><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(P);
        else oi_fast_kernel<0><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(P);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t err = cudaGetLastError();
        if(temp) cudaFreeAsync(temp, stream);
        if(err != cudaSuccess) return fail(GPP_ERR_CUDA, "CUDA error %s launching oi_fast_kernel: %s", cudaGetErrorName(err), cudaGetErrorString(err));
        return GPP_OK;
    }
    if(kcap <= CHOL_K && structure_is_symmetric(*structure)) {
        // 30 < k <= 128: packed Cholesky in shared memory, dynamic blocks of 32 points
        const int mode = structure_mode(*structure);
        const bool small = kcap <= 64;
        void (*kernel)(OiParams) = small ? (mode == 1 ? oi_chol_kernel<1, 64> : oi_chol_kernel<0, 64>)
                                         : (mode == 1 ? oi_chol_kernel<1, 128> : oi_chol_kernel<0, 128>);
        const size_t smem = (small ? CholSmem<64>::bytes(kcap) : CholSmem<128>::bytes(kcap)) * CHOL_WARPS + sizeof(unsigned short) * ((size_t) kcap * (kcap + 1) / 2);
        int per_sm = 1;
        GPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        GPP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CHOL_WARPS * 32, smem));
        const long long want = ((long long) count + 32 * CHOL_WARPS - 1) / (32 * CHOL_WARPS);
        const unsigned grid = (unsigned) std::max<long long>(1, std::min<long long>(want, (long long) sms * std::max(per_sm, 1)));
        // whole rows of a known grid -> 4 x 8 tiles
        P.tile_nx = (bp->shape_nx > 0 && first % bp->shape_nx == 0 && count % bp->shape_nx == 0) ? bp->shape_nx : 0;
        // launch scratch: the work counter, then one cache of solved systems per warp (its stamps are reset by the kernel)
        const size_t lru_bytes = (small ? sizeof(CholLru<64>) : sizeof(CholLru<128>)) * (size_t) grid * CHOL_WARPS;
        unsigned char* scratch = nullptr;
        GPP_CUDA(cudaMallocAsync((void**) &scratch, 256 + lru_bytes, stream));
        GPP_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int), stream));
        int* counter = reinterpret_cast<int*>(scratch);
        P.work_counter = counter;
        P.workspace = scratch + 256;
        kernel<<<grid, CHOL_WARPS * 32, smem, stream>>>(P);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t err = cudaGetLastError();
        cudaFreeAsync(scratch, stream);
        if(err != cudaSuccess) return fail(GPP_ERR_CUDA, "CUDA error %s launching oi_chol_kernel: %s", cudaGetErrorName(err), cudaGetErrorString(err));
        return GPP_OK;
    }
    return launch_general(P, kcap, count, stream);
}

extern "C" {

// ---- spatially varying structure functions (structure.cpp:168-184 and the sibling constructors) ------------
int gpp_structure_field_create(const gpp_points* grid, const float* h, const float* v, const float* w, gpp_structure_field** out) {
    if(!out) return fail(GPP_ERR_INVALID_ARGUMENT, "out must not be NULL");
    *out = nullptr;
    if(!grid || !h || !v || !w) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(grid->n <= 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Grid size not the same as scale size");
    gpp_structure_field* f = new(std::nothrow) gpp_structure_field();
    if(!f) return fail(GPP_ERR_RUNTIME, "out of memory");
    f->grid = grid;
    f->h.assign(h, h + grid->n);
    f->v.assign(v, v + grid->n);
    f->w.assign(w, w + grid->n);
    *out = f;
    return GPP_OK;
}
void gpp_structure_field_destroy(gpp_structure_field* f) { delete f; }

int gpp_structure_field_lookup_host(const gpp_structure_field* f, const float* lats, const float* lons, int n, float* h, float* v, float* w) {
    if(!f) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL field");
    if(n <= 0) return GPP_OK;
    std::vector<int> index(n);
    GPP_TRY(gpp_points_nearest_host(f->grid, lats, lons, n, 1, index.data()));    // m_grid.get_nearest_neighbour, structure.cpp:189
    for(int i = 0; i < n; i++) {
        const int k = index[i];
        if(k < 0 || k >= f->grid->n) return fail(GPP_ERR_RUNTIME, "Invalid I[0]");
        h[i] = f->h[k]; v[i] = f->v[k]; w[i] = f->w[k];
    }
    return GPP_OK;
}

int gpp_structure_field_localization_distance(const gpp_structure_field* f, int type, float min_rho, float lat, float lon, float* out) {
    float h = 0.f, v = 0.f, w = 0.f, lc = 0.f;
    double ld = 0.0;
    GPP_TRY(spatial_loc_constants(type, min_rho, &lc, &ld));
    GPP_TRY(gpp_structure_field_lookup_host(f, &lat, &lon, 1, &h, &v, &w));
    *out = spatial_loc_dist_host(type, h, lc, ld);
    return GPP_OK;
}

// gridpp::optimal_interpolation / _full (oi.cpp:26-412) with a structure function whose scales vary in space: every
// point evaluates corr() with the scales of the field node nearest to it (its own position for corr_background, the
// first observation's for the observation-observation matrix, which is therefore not symmetric). General kernel.
int gpp_optimal_interpolation_spatial_host(const gpp_points* bpoints, const float* background, const float* bvariance,
                                           const gpp_points* opoints, const float* pobs, const float* obs_variance,
                                           const float* pbackground, const float* bvariance_at_points, int structure_type,
                                           const gpp_structure_field* field, float min_rho, int max_points,
                                           int allow_extrapolation, float* analysis, float* analysis_variance) {
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");
    if(!bpoints || !opoints || !field) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL argument");
    if(bpoints->type != opoints->type)
        return fail(GPP_ERR_INVALID_ARGUMENT, "Both background points and observations points must be of same coordinate type (lat/lon or x/y)");
    float loc_c;
    double loc_d;
    GPP_TRY(spatial_loc_constants(structure_type, min_rho, &loc_c, &loc_d));
    GPP_TRY(ensure_device());
    const int nB = bpoints->n, nS = opoints->n;
    if(nS == 0 || nB == 0) {
        if(nB > 0) std::memcpy(analysis, background, sizeof(float) * nB);
        if(analysis_variance)
            for(int i = 0; i < nB; i++) analysis_variance[i] = bvariance ? bvariance[i] : 1.f;
        return GPP_OK;
    }
    // the descriptor the kernel specialises per point
    gpp_structure s;
    std::memset(&s, 0, sizeof(s));
    s.n_terms = 1;
    s.has_cv = 0;
    s.cv_dist = NAN;
    s.term[0].type = structure_type;
    s.term[0].min_rho = min_rho;
    // scales at the observations and at the background points
    std::vector<float> oscale[3], bscale[3];
    for(int c = 0; c < 3; c++) { oscale[c].resize(nS); bscale[c].resize(nB); }
    GPP_TRY(gpp_structure_field_lookup_host(field, opoints->lats.data(), opoints->lons.data(), nS, oscale[0].data(), oscale[1].data(), oscale[2].data()));
    GPP_TRY(gpp_structure_field_lookup_host(field, bpoints->lats.data(), bpoints->lons.data(), nB, bscale[0].data(), bscale[1].data(), bscale[2].data()));
    float max_R = 0.f;
    for(int i = 0; i < nB; i++) {
        const float R = spatial_loc_dist_host(structure_type, bscale[0][i], loc_c, loc_d);
        if(is_valid(R)) max_R = std::max(max_R, R);
    }
    s.term[0].loc_dist = max_R;
    std::vector<char> valid(nS);
    std::vector<double> innov(nS);
    std::vector<float> ratio(nS);
    for(int i = 0; i < nS; i++) {
        valid[i] = is_valid(pobs[i]) && is_valid(pbackground[i]);
        innov[i] = (double) pobs[i] - (double) pbackground[i];
        ratio[i] = obs_variance[i] / (bvariance_at_points ? bvariance_at_points[i] : 1.f);
    }
    gpp_oi_obs obs;
    GPP_TRY(build_obs_table(opoints, valid, innov, ratio, max_R, &obs, nullptr, oscale));
    gpp_points* bp = const_cast<gpp_points*>(bpoints);
    GPP_TRY(bp->ensure_on_device());
    DeviceBuffer<float> d_bg, d_bvar, d_out, d_var, d_sc[3];
    GPP_TRY(d_bg.upload(background, nB));
    if(bvariance) GPP_TRY(d_bvar.upload(bvariance, nB));
    GPP_TRY(d_out.alloc(nB));
    if(analysis_variance) GPP_TRY(d_var.alloc(nB));
    for(int c = 0; c < 3; c++) GPP_TRY(d_sc[c].upload(bscale[c].data(), nB));
    if(obs.n_valid == 0) {
        GPP_LAUNCH(copy_background_kernel, (unsigned) ((nB + 255) / 256), 256, 0, 0, d_bg.ptr, bvariance ? d_bvar.ptr : nullptr, 0, nB, d_out.ptr,
                   analysis_variance ? d_var.ptr : nullptr);
    }
    else {
        OiParams P;
        P.gx = bp->dx.ptr; P.gy = bp->dy.ptr; P.gz = bp->dz.ptr; P.gelev = bp->delev.ptr; P.glaf = bp->dlaf.ptr;
        P.background = d_bg.ptr;
        P.bvariance = bvariance ? d_bvar.ptr : nullptr;
        P.analysis = d_out.ptr;
        P.analysis_variance = analysis_variance ? d_var.ptr : nullptr;
        P.first = 0;
        P.count = nB;
        P.obs = obs.view();
        P.s = s;
        P.R = max_R;
        P.allow_extrapolation = allow_extrapolation;
        P.tile_nx = 0;
        P.workspace = nullptr;
        std::memset(&P.plan, 0, sizeof(P.plan));
        P.work_counter = nullptr;
        P.sbh = d_sc[0].ptr; P.sbv = d_sc[1].ptr; P.sbw = d_sc[2].ptr;
        P.loc_c = loc_c;
        P.loc_d = loc_d;
        int kcap = max_points > 0 ? std::min(max_points, obs.n_valid) : obs.n_valid;
        if(max_points == 0) {   // unlimited: bound k by the largest neighbourhood of the largest radius
            int hmax = 0;
            GPP_TRY(count_max_candidates(bp, 0, nB, d_bg.ptr, P.obs, max_R, 0, &hmax));
            kcap = std::min(kcap, std::max(hmax, 1));
        }
        P.k = kcap;
        GPP_TRY(launch_general(P, kcap, nB, 0));
    }
    GPP_TRY(d_out.download(analysis, nB));
    if(analysis_variance) GPP_TRY(d_var.download(analysis_variance, nB));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_optimal_interpolation_host(const gpp_points* bpoints, const float* background, const float* bvariance,
                                   const gpp_points* opoints, const float* pobs, const float* obs_variance,
                                   const float* pbackground, const float* bvariance_at_points, const gpp_structure* structure,
                                   int max_points, int allow_extrapolation, float* analysis, float* analysis_variance) {
    // argument checks in the order of oi.cpp:151-186 (sizes are implied by the flat signature)
    if(max_points < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "max_points must be >= 0");
    if(!bpoints || !opoints) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    if(bpoints->type != opoints->type)
        return fail(GPP_ERR_INVALID_ARGUMENT, "Both background points and observations points must be of same coordinate type (lat/lon or x/y)");
    GPP_TRY(check_structure(structure));
    GPP_TRY(ensure_device());
    const int nB = bpoints->n, nS = opoints->n;
    if(nS == 0 || nB == 0) {   // oi.cpp:189-190: analysis = background (analysis_variance left as bvariance)
        if(nB > 0) std::memcpy(analysis, background, sizeof(float) * nB);
        if(analysis_variance)
            for(int i = 0; i < nB; i++) analysis_variance[i] = bvariance ? bvariance[i] : 1.f;
        return GPP_OK;
    }
    Trace trace("optimal_interpolation_host");
    gpp_oi_obs* obs = nullptr;
    GPP_TRY(gpp_oi_obs_create(opoints, pobs, obs_variance, pbackground, bvariance_at_points, structure, &obs));
    trace.lap("observation table");
    DeviceBuffer<float> d_bg, d_bvar, d_out, d_var;
    // Large grids: row blocks are uploaded, analysed and sent back as a pipeline (the results through a pinned staging
    // buffer: a D2H straight into the caller's pageable array runs at ~5 GB/s and would add ~30 % to the call).
    const int nx = bpoints->shape_nx;
    const int n_rows = nx > 0 ? nB / nx : 0;
    static const int want_chunks = [] { const char* e = getenv("GPP_OI_CHUNKS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 4; }();
    // fields of at least this many points are pipelined by row blocks (GPP_OI_PIPELINE_MIN overrides it for experiments)
    static const int pipeline_min = [] { const char* e = getenv("GPP_OI_PIPELINE_MIN"); const int v = e ? atoi(e) : 0; return v > 0 ? v : (1 << 21); }();
    const int n_chunks = (nx > 0 && nB % nx == 0 && nB >= pipeline_min) ? std::min(want_chunks, n_rows) : 1;
    int rc = n_chunks > 1 ? d_bg.alloc(nB) : d_bg.upload(background, nB);
    if(rc == GPP_OK && bvariance) rc = n_chunks > 1 ? d_bvar.alloc(nB) : d_bvar.upload(bvariance, nB);
    if(rc == GPP_OK) rc = d_out.alloc(nB);
    if(rc == GPP_OK && analysis_variance) rc = d_var.alloc(nB);
    if(trace.on) { cudaStreamSynchronize(0); trace.lap(n_chunks > 1 ? "alloc" : "alloc + H2D"); }
    if(rc == GPP_OK && n_chunks > 1) rc = analyse_pipelined(bpoints, nB, nx, n_chunks, background, bvariance, d_bg.ptr, d_bvar.ptr, obs, structure,
                                                            max_points, allow_extrapolation, d_out.ptr, analysis_variance ? d_var.ptr : nullptr,
                                                            analysis);
    else if(rc == GPP_OK) {
        rc = gpp_optimal_interpolation_device(bpoints, 0, nB, d_bg.ptr, bvariance ? d_bvar.ptr : nullptr, obs, structure, max_points,
                                              allow_extrapolation, d_out.ptr, analysis_variance ? d_var.ptr : nullptr, nullptr);
        if(trace.on) { cudaStreamSynchronize(0); trace.lap("kernels"); }
        if(rc == GPP_OK) rc = d_out.download(analysis, nB);
    }
    if(rc == GPP_OK && analysis_variance) rc = d_var.download(analysis_variance, nB);
    if(rc == GPP_OK) {
        cudaError_t err = cudaStreamSynchronize(0);
        if(err != cudaSuccess) rc = fail(GPP_ERR_CUDA, "CUDA error %s: %s", cudaGetErrorName(err), cudaGetErrorString(err));
    }
    else cudaStreamSynchronize(0);
    trace.lap(n_chunks > 1 ? "H2D + kernels + D2H (pipelined)" : "D2H");
    gpp_oi_obs_destroy(obs);
    d_bg.release(); d_bvar.release(); d_out.release(); d_var.release();
    trace.lap("free");
    return rc;
}

}  // extern "C"

#ifdef OI_STATS
extern "C" int gpp_debug_oi_stats(unsigned long long* out, int reset) {
    GPP_CUDA(cudaDeviceSynchronize());
    GPP_CUDA(cudaMemcpyFromSymbol(out, g_oi_stats, sizeof(unsigned long long) * 4));
    if(reset) {
        unsigned long long zero[4] = {0, 0, 0, 0};
        GPP_CUDA(cudaMemcpyToSymbol(g_oi_stats, zero, sizeof(zero)));
    }
    return GPP_OK;
}
#endif
