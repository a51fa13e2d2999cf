// Neighbourhood statistics on the device ("K1"-"K3" in SURVEY.md).
// Replaces gridpp::neighbourhood (Mean/Sum/Count/Min/Max) and gridpp::neighbourhood_quantile_fast,
// src/api/neighbourhood.cpp:28-242 and :296-409.
//
// All kernels share one tiling: a CTA of 256 threads owns a strip of 256 staged columns (TX = 256 - 2*hw
// output columns plus hw halo columns on each side) and walks down a chunk of output rows. Thread t owns staged
// column t: it streams the rows of that column (coalesced across the CTA), keeps the vertical window state in
// registers / a shared-memory ring, and every RB rows the CTA turns the per-column states into outputs with a
// horizontal pass in which each thread slides along 8 consecutive pixels. The input is read from HBM once (halo
// re-reads hit L2) and the output written once: 8 B/pixel of algorithmic traffic.
// Windows are CLIPPED at the domain edges (neighbourhood.cpp:104-107): out-of-domain cells are staged as NaN,
// which every statistic ignores exactly like a missing value.
#include "common.cuh"

#include <algorithm>
#include <cstring>

using namespace gpp;

namespace gpp {
int nbh_tma_try(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out, int hw, int statistic, float* d_output,
                cudaStream_t stream, int* handled);   // neighbourhood_tma.cu
int qf_tma_try(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out, float quantile, const float* d_quantile_field,
               int hw, const float* thresholds, int T, float* d_output, cudaStream_t stream, int* handled);   // quantile_tma.cu
}

namespace {

constexpr int NT = 256;        // threads per CTA = staged columns per strip
constexpr int RB = 8;          // output rows per horizontal pass
constexpr int SEG = 8;         // consecutive pixels per thread in the horizontal pass
constexpr int CHUNK = 64;      // output rows per CTA
constexpr int HW_FUSED_MAX = 64;
constexpr int PADW = NT + NT / 8 + 8;   // padded row length of the fp64 line buffers (index e -> e + e/8)

__device__ __forceinline__ int pad8(int e) { return e + (e >> 3); }

struct TileArgs {
    const float* in;
    float* out;
    int n_rows_in, nx, row0, n_rows_out, hw;
};

__device__ __forceinline__ float load_cell(const TileArgs& a, int r, int x) {
    return (r >= 0 && r < a.n_rows_in && x >= 0 && x < a.nx) ? __ldg(a.in + (size_t) r * a.nx + x) : NAN;
}

// ------------------------------------------------------------------ K1: mean / sum / count ------------
// neighbourhood.cpp:45-145. The reference builds a double summed-area table and an int count table; here the
// window sum is accumulated directly in fp64 (vertical running sum per column over at most CHUNK + 2*hw rows,
// horizontal sliding sum over at most SEG + 2*hw columns), so no large prefix value ever enters the sum.
struct __align__(16) SumRec {   // one staged column of one output row: vertical window sum and valid count
    double sum;
    int cnt;
    int pad;
};
constexpr int RCP_TABLE = 1024;   // 1 / count for count < RCP_TABLE lives in shared memory (halfwidth <= 15)

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e38f; }   // == is_valid(v)

// Coalesced store of up to RB staged rows (obuf[b * PADW + pad8(column)]) to output rows y0 .. y0+nb-1.
__device__ __forceinline__ void store_rows(const TileArgs& a, const float* obuf, int y0, int nb, int TX) {
    const int tid = threadIdx.x;
    const int x = blockIdx.x * TX + tid;
    if(tid < TX && x < a.nx) {
        float* dst = a.out + (long long) (y0 - a.row0) * a.nx + x;
        const float* ob = obuf + pad8(tid);
        if(nb == RB) {
            #pragma unroll
            for(int b = 0; b < RB; b++) { *dst = ob[b * PADW]; dst += a.nx; }
        }
        else
            for(int b = 0; b < nb; b++) { *dst = ob[b * PADW]; dst += a.nx; }
    }
}

// STAT: 0 = Mean, 1 = Sum, 2 = Count
template <int STAT>
__global__ void __launch_bounds__(NT) nbh_sum_kernel(const TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int w = 2 * a.hw + 1;
    const int TX = NT - 2 * a.hw;
    SumRec* line = reinterpret_cast<SumRec*>(smem);                              // [RB][PADW]
    double* rcp = reinterpret_cast<double*>(line + RB * PADW);                   // [RCP_TABLE]
    float* obuf = reinterpret_cast<float*>(rcp + RCP_TABLE);                     // [RB][PADW]
    float* ring = obuf + RB * PADW;                                              // [w][NT]
    const int tid = threadIdx.x;
    const int x_stage = blockIdx.x * TX - a.hw + tid;
    const bool col_ok = x_stage >= 0 && x_stage < a.nx;
    const int y_begin = a.row0 + blockIdx.y * CHUNK;
    const int y_end = min(y_begin + CHUNK, a.row0 + a.n_rows_out);
    const bool use_table = w * w < RCP_TABLE;

    if(STAT == 0 && use_table)
        for(int c = tid; c < RCP_TABLE; c += NT) rcp[c] = c > 0 ? 1.0 / (double) c : 0.0;
    // records past the last staged column are read (clamped) by the sliding window of pixels that are never
    // stored; keep them at zero so that their counts stay valid table indices
    if(tid < RB * (PADW - 287)) {
        const int tail = PADW - 287;   // pad8(NT - 1) == 286 is the last record the vertical pass writes
        SumRec zero = {0.0, 0, 0};
        line[(tid / tail) * PADW + 287 + tid % tail] = zero;
    }
    float* const ring_col = ring + tid;
    for(int s = 0; s < w; s++) ring_col[s * NT] = NAN;
    const int ring_len = w * NT;
    // prime the vertical window with rows [y_begin - hw, y_begin + hw - 1]
    double csum = 0.0;
    int ccnt = 0;
    int slot = 0;   // ring offset (in floats) of the next row to load
    int r = y_begin - a.hw;
    const float* src = a.in + (long long) r * a.nx + x_stage;   // only dereferenced when the cell is inside the domain
    for(; r < y_begin + a.hw; r++, src += a.nx) {
        const float v = (col_ok && r >= 0 && r < a.n_rows_in) ? __ldg(src) : NAN;
        ring_col[slot] = v;
        slot += NT;
        if(slot == ring_len) slot = 0;
        if(finite_f(v)) { csum += (double) v; ccnt++; }
    }
    __syncthreads();   // the reciprocal table
    SumRec* const line_col = line + pad8(tid);
    for(int y0 = y_begin; y0 < y_end; y0 += RB) {
        const int nb = min(RB, y_end - y0);
        // ---- vertical pass: 8 rows enter the window one after the other (all 8 loads are issued first); row
        // y0+b+hw enters, row y0+b-hw-1 (same ring slot) leaves. Rows past the end of the chunk are processed too;
        // their records are simply not used.
        float vnew[RB];
        #pragma unroll
        for(int b = 0; b < RB; b++) {
            vnew[b] = (col_ok && r + b >= 0 && r + b < a.n_rows_in) ? __ldg(src) : NAN;
            src += a.nx;
        }
        r += RB;
        #pragma unroll
        for(int b = 0; b < RB; b++) {
            const float v_old = ring_col[slot];
            ring_col[slot] = vnew[b];
            slot += NT;
            if(slot == ring_len) slot = 0;
            if(finite_f(vnew[b])) { csum += (double) vnew[b]; ccnt++; }
            if(finite_f(v_old)) { csum -= (double) v_old; ccnt--; }
            SumRec rec = {csum, ccnt, 0};
            line_col[b * PADW] = rec;
        }
        __syncthreads();
        // ---- horizontal pass: 8 rows x 32 segments of 8 pixels; each thread slides along its segment
        {
            const int b = tid >> 5, xo0 = (tid & 31) * SEG;
            if(b < nb && xo0 < TX) {
                const SumRec* ls = line + b * PADW;
                const SumRec* lo = ls + pad8(xo0);           // records xo0 .. xo0+7 are contiguous from here
                double s = 0.0;
                int c = 0;
                int t = xo0;
                #pragma unroll 5
                for(int j = 0; j < w; j++, t++) {
                    const SumRec rec = ls[t + (t >> 3)];
                    s += rec.sum;
                    c += rec.cnt;
                }
                // t == xo0 + w: the next record to enter
                float* ob = obuf + b * PADW + pad8(xo0);   // xo0 is a multiple of 8: the 8 pixels are contiguous
                #pragma unroll
                for(int p = 0; p < SEG; p++) {
                    float o;
                    if(STAT == 2) o = (float) c;
                    else if(STAT == 1) o = c > 0 ? (float) s : NAN;
                    else o = c > 0 ? (use_table ? (float) (s * rcp[c]) : (float) (s / (double) c)) : NAN;   // neighbourhood.cpp:133-142
                    ob[p] = o;
                    if(p + 1 < SEG) {
                        const SumRec in_rec = ls[min(t + (t >> 3), PADW - 1)], out_rec = lo[p];
                        t++;
                        s += in_rec.sum - out_rec.sum;
                        c += in_rec.cnt - out_rec.cnt;
                    }
                }
            }
        }
        __syncthreads();
        store_rows(a, obuf, y0, nb, TX);
        // the next batch overwrites `line` only after its own vertical pass, and obuf only after the next
        // __syncthreads, so no extra barrier is needed here
    }
}

// ------------------------------------------------------------------ K2: min / max ---------------------
// neighbourhood.cpp:146-210: extreme of the valid values in the clipped window (the reference's sliver scheme
// and its border brute force both reduce to that). Invalid cells are mapped to +inf (min) / -inf (max); an
// all-invalid window therefore ends at +-inf, which is reported as NaN (infinite inputs are themselves
// "invalid", util.cpp:16-18, so a genuine result is never infinite).
//
// Eight consecutive windows of width w (w >= 8) over v[0 .. w+6] all contain the core v[7 .. w-1]; window i is
// min(suffix-min of v[i..6], core, prefix-min of v[w .. w+i-1]): (w - 8) + 6 + 6 + 14 operations for eight
// outputs instead of 8 (w - 1). The vertical pass applies this to the 8 rows of a batch (per column, over a ring
// of w + 7 rows), the horizontal pass to the 8 pixels of a thread's segment.
template <bool IS_MAX>
__device__ __forceinline__ float ext(float x, float y) { return IS_MAX ? fmaxf(x, y) : fminf(x, y); }

// v(j) for j in [0, w + 7) -> eight window extremes. `v` is a callable returning the j-th staged value.
template <bool IS_MAX, class F>
__device__ __forceinline__ void eight_windows(F v, int w, float (&out)[RB]) {
    const float ident = IS_MAX ? -INFINITY : INFINITY;
    if(w >= RB) {
        float suf[RB], pre[RB];
        suf[RB - 1] = ident;
        #pragma unroll
        for(int i = RB - 2; i >= 0; i--) suf[i] = ext<IS_MAX>(suf[i + 1], v(i));
        float core = v(RB - 1);
        for(int j = RB; j < w; j++) core = ext<IS_MAX>(core, v(j));
        pre[0] = ident;
        #pragma unroll
        for(int i = 1; i < RB; i++) pre[i] = ext<IS_MAX>(pre[i - 1], v(w + i - 1));
        #pragma unroll
        for(int i = 0; i < RB; i++) out[i] = ext<IS_MAX>(ext<IS_MAX>(suf[i], core), pre[i]);
    }
    else {
        #pragma unroll
        for(int i = 0; i < RB; i++) {
            float m = ident;
            for(int j = 0; j < w; j++) m = ext<IS_MAX>(m, v(i + j));
            out[i] = m;
        }
    }
}

template <bool IS_MAX>
__global__ void __launch_bounds__(NT) nbh_minmax_kernel(const TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const float ident = IS_MAX ? -INFINITY : INFINITY;
    const int w = 2 * a.hw + 1;
    const int RS = w + RB - 1;                         // ring rows: everything the 8 rows of a batch need
    const int TX = NT - 2 * a.hw;
    float* line = reinterpret_cast<float*>(smem);      // [RB][NT] vertical extremes
    float* obuf = line + RB * NT;                      // [RB][PADW]
    float* ring = obuf + RB * PADW;                    // [RS][NT]
    const int tid = threadIdx.x;
    const int x_stage = blockIdx.x * TX - a.hw + tid;
    const bool col_ok = x_stage >= 0 && x_stage < a.nx;
    const int y_begin = a.row0 + blockIdx.y * CHUNK;
    const int y_end = min(y_begin + CHUNK, a.row0 + a.n_rows_out);

    // ring slot of input row r is (r - (y_begin - hw)) mod RS; prime rows [y_begin - hw, y_begin + hw - 1]
    int r = y_begin - a.hw;
    const float* src = a.in + (long long) r * a.nx + x_stage;
    int slot = 0;
    for(; r < y_begin + a.hw; r++, src += a.nx) {
        const float v = (col_ok && r >= 0 && r < a.n_rows_in) ? __ldg(src) : NAN;
        ring[slot * NT + tid] = finite_f(v) ? v : ident;
        slot = slot + 1 == RS ? 0 : slot + 1;
    }
    int start = 0;   // ring slot of row y0 - hw
    for(int y0 = y_begin; y0 < y_end; y0 += RB) {
        const int nb = min(RB, y_end - y0);
        // rows y0+hw .. y0+hw+7 enter (rows beyond the batch are loaded too: they are in the domain or NaN)
        float vnew[RB];
        #pragma unroll
        for(int b = 0; b < RB; b++) {
            vnew[b] = (col_ok && r + b >= 0 && r + b < a.n_rows_in) ? __ldg(src) : NAN;
            src += a.nx;
        }
        r += RB;
        #pragma unroll
        for(int b = 0; b < RB; b++) {
            ring[slot * NT + tid] = finite_f(vnew[b]) ? vnew[b] : ident;
            slot = slot + 1 == RS ? 0 : slot + 1;
        }
        {
            float out[RB];
            const float* base = ring + tid;
            eight_windows<IS_MAX>([&](int j) { int s = start + j; if(s >= RS) s -= RS; return base[s * NT]; }, w, out);
            #pragma unroll
            for(int b = 0; b < RB; b++) line[b * NT + tid] = out[b];
        }
        start += RB;
        if(start >= RS) start -= RS;
        __syncthreads();
        {
            const int b = tid >> 5, xo0 = (tid & 31) * SEG;
            if(b < nb && xo0 < TX) {
                // staged columns xo0 .. xo0 + w + 6 (clamped reads beyond the strip only feed pixels that are not stored)
                const float* l = line + b * NT;
                float out[RB];
                eight_windows<IS_MAX>([&](int j) { return l[min(xo0 + j, NT - 1)]; }, w, out);
                #pragma unroll
                for(int p = 0; p < SEG; p++) obuf[b * PADW + pad8(xo0) + p] = fabsf(out[p]) == INFINITY ? NAN : out[p];
            }
        }
        __syncthreads();
        store_rows(a, obuf, y0, nb, TX);
    }
}

// ------------------------------------------------------------------ large half-widths -----------------
// Separable two-pass fallback through a temporary plane for half-widths beyond the fused kernels' strip. One
// thread per pixel, direct loop over the clipped 1-D window (cost O(w) per pixel and pass).
// Pass 1 (vertical): per pixel the fp64 sum and count (or the extreme) of the column window.
__global__ void nbh_big_vertical_kernel(const TileArgs a, int statistic, double* __restrict__ tsum, int* __restrict__ tcnt,
                                        float* __restrict__ tval, int r_lo, int n_rows_tmp) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= (size_t) n_rows_tmp * a.nx) return;
    int ry = (int) (i / a.nx), x = (int) (i % a.nx);
    int y = r_lo + ry;
    int i0 = max(0, y - a.hw), i1 = min(a.n_rows_in - 1, y + a.hw);
    if(statistic == GPP_MIN || statistic == GPP_MAX) {
        bool is_max = statistic == GPP_MAX;
        float m = is_max ? -INFINITY : INFINITY;
        for(int r = i0; r <= i1; r++) {
            float v = a.in[(size_t) r * a.nx + x];
            if(is_valid(v)) m = is_max ? fmaxf(m, v) : fminf(m, v);
        }
        tval[i] = m;
    }
    else {
        double s = 0.0;
        int c = 0;
        for(int r = i0; r <= i1; r++) {
            float v = a.in[(size_t) r * a.nx + x];
            if(is_valid(v)) { s += (double) v; c++; }
        }
        tsum[i] = s;
        tcnt[i] = c;
    }
}
__global__ void nbh_big_horizontal_kernel(const TileArgs a, int statistic, const double* __restrict__ tsum,
                                          const int* __restrict__ tcnt, const float* __restrict__ tval) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= (size_t) a.n_rows_out * a.nx) return;
    int ry = (int) (i / a.nx), x = (int) (i % a.nx);
    int j0 = max(0, x - a.hw), j1 = min(a.nx - 1, x + a.hw);
    size_t base = (size_t) ry * a.nx;
    float o = NAN;
    if(statistic == GPP_MIN || statistic == GPP_MAX) {
        bool is_max = statistic == GPP_MAX;
        float m = is_max ? -INFINITY : INFINITY;
        for(int j = j0; j <= j1; j++) m = is_max ? fmaxf(m, tval[base + j]) : fminf(m, tval[base + j]);
        o = isinf(m) ? NAN : m;
    }
    else {
        double s = 0.0;
        int c = 0;
        for(int j = j0; j <= j1; j++) { s += tsum[base + j]; c += tcnt[base + j]; }
        if(statistic == GPP_COUNT) o = (float) c;
        else if(c > 0) o = statistic == GPP_MEAN ? (float) (s / (double) c) : (float) s;
    }
    a.out[i] = o;
}

// ------------------------------------------------------------------ K3: quantile_fast -----------------
// neighbourhood.cpp:302-409. For every pixel the reference needs F_t = #(valid v <= thr_t) / #(valid v) over
// the clipped window, for every threshold t; it gets them from one summed-area-table pass per threshold. The
// sums are sums of 0/1 indicators, i.e. exact integers, and F_t = float(double(n_le) / n_valid) == n_le / n_valid
// rounded once to float (both counts are < 2^24). Here each cell is classified once into the bin
// b(v) = #(sorted thresholds < v); a per-column histogram over the vertical window lives in shared memory and
// the horizontal pass slides a (T+1)-bin window histogram along SEG pixels per thread; the CDF at sorted
// position s is the prefix sum over bins <= s. The CDF inversion is gridpp::interpolate (util.cpp:377-414)
// with get_lower_index / get_upper_index (util.cpp:339-376) restated verbatim.
constexpr int QF_MAX_T = 64;

struct QfArgs {
    TileArgs t;
    const float* quantile_field;   // may be NULL
    float quantile;
    int T;
    float thr[QF_MAX_T];           // thresholds as given
    float sorted[QF_MAX_T];        // valid thresholds ascending
    int n_sorted;
    int rank[QF_MAX_T];            // position of threshold t in `sorted` (-1: NaN threshold, F_t = 0)
    int rb;                        // rows per horizontal pass (fits shared memory)
};

__device__ __forceinline__ int qf_bin(const QfArgs& q, float v) {
    // number of sorted thresholds strictly below v; n_sorted + 1 marks an invalid cell
    if(!is_valid(v)) return q.n_sorted + 1;
    int lo = 0, hi = q.n_sorted;
    while(lo < hi) {
        int mid = (lo + hi) >> 1;
        if(q.sorted[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ float qf_interpolate(float x, const float* iX, const float* iY, int n) {
    // gridpp::interpolate, util.cpp:377-414; iX has no missing values here
    if(!is_valid(x)) return NAN;
    if(x > iX[n - 1]) return iY[n - 1];
    if(x < iX[0]) return iY[0];
    int i0 = -1, i1 = -1;
    for(int i = 0; i < n; i++) {           // get_lower_index, util.cpp:339-357
        float c = iX[i];
        if(c < x) i0 = i;
        else if(c == x) { i0 = i; break; }
        else if(c > x) break;
    }
    for(int i = n - 1; i >= 0; i--) {      // get_upper_index, util.cpp:358-376
        float c = iX[i];
        if(c > x) i1 = i;
        else if(c == x) { i1 = i; break; }
        else if(c < x) break;
    }
    if(i0 < 0 || i1 < 0) return NAN;
    float x0 = iX[i0], x1 = iX[i1], y0 = iY[i0], y1 = iY[i1];
    if(x0 == x1) {
        if(i0 == 0 && i1 == n - 1) return __fdiv_rn(__fadd_rn(y0, y1), 2.f);
        if(i0 == 0) return y1;
        if(i1 == n - 1) return y0;
        return __fdiv_rn(__fadd_rn(y0, y1), 2.f);
    }
    // y0 + (y1 - y0) * (x - x0) / (x1 - x0), float, left to right
    return __fadd_rn(y0, __fdiv_rn(__fmul_rn(__fsub_rn(y1, y0), __fsub_rn(x, x0)), __fsub_rn(x1, x0)));
}

__global__ void __launch_bounds__(NT) nbh_quantile_fast_kernel(const __grid_constant__ QfArgs q) {
    extern __shared__ __align__(16) unsigned char smem[];
    const TileArgs& a = q.t;
    const int w = 2 * a.hw + 1;
    const int TX = NT - 2 * a.hw;
    const int NB = q.n_sorted + 2;                       // bins 0..n_sorted, plus the "invalid" bin (never counted)
    const int NBC = q.n_sorted + 1;                      // counted bins
    unsigned short* colhist = reinterpret_cast<unsigned short*>(smem);            // [NBC][NT] current vertical window
    unsigned short* line = colhist + NBC * NT;                                     // [rb][NBC][NT] snapshots
    float* obuf = reinterpret_cast<float*>(line + (size_t) q.rb * NBC * NT);       // [rb][NT]
    unsigned char* ring = reinterpret_cast<unsigned char*>(obuf + q.rb * NT);      // [w][NT] bin of each cell
    const int tid = threadIdx.x;
    const int x_stage = blockIdx.x * TX - a.hw + tid;
    const int y_begin = a.row0 + blockIdx.y * CHUNK;
    const int y_end = min(y_begin + CHUNK, a.row0 + a.n_rows_out);
    (void) NB;

    for(int b = 0; b < NBC; b++) colhist[b * NT + tid] = 0;
    for(int s = 0; s < w; s++) ring[s * NT + tid] = (unsigned char) (q.n_sorted + 1);
    int slot = 0;
    for(int r = y_begin - a.hw; r < y_begin + a.hw; r++) {
        int b = qf_bin(q, load_cell(a, r, x_stage));
        ring[slot * NT + tid] = (unsigned char) b;
        slot = slot + 1 == w ? 0 : slot + 1;
        if(b <= q.n_sorted) colhist[b * NT + tid]++;
    }
    for(int y0 = y_begin; y0 < y_end; y0 += q.rb) {
        const int nb = min(q.rb, y_end - y0);
        for(int rb = 0; rb < nb; rb++) {
            int b_new = qf_bin(q, load_cell(a, y0 + rb + a.hw, x_stage));
            int b_old = ring[slot * NT + tid];
            ring[slot * NT + tid] = (unsigned char) b_new;
            slot = slot + 1 == w ? 0 : slot + 1;
            if(b_new <= q.n_sorted) colhist[b_new * NT + tid]++;
            if(b_old <= q.n_sorted) colhist[b_old * NT + tid]--;
            unsigned short* l = line + (size_t) rb * NBC * NT;
            for(int b = 0; b < NBC; b++) l[b * NT + tid] = colhist[b * NT + tid];
        }
        __syncthreads();
        // horizontal pass: work items are (row, segment); a thread may take several when rb < 8
        for(int item = tid; item < nb * 32; item += NT) {
            const int rb = item >> 5, seg = item & 31;
            const int xo0 = seg * SEG;
            if(xo0 >= TX) continue;
            const unsigned short* l = line + (size_t) rb * NBC * NT;
            int hist[QF_MAX_T + 1];
            #pragma unroll 1
            for(int b = 0; b < NBC; b++) {
                int c = 0;
                for(int j = 0; j < w; j++) c += l[b * NT + xo0 + j];
                hist[b] = c;
            }
            for(int p = 0; p < SEG; p++) {
                const int xo = xo0 + p;
                if(xo >= TX) break;
                const int y = y0 + rb, x = blockIdx.x * TX + xo;
                float result = NAN;
                if(x < a.nx) {
                    int n_valid = 0;
                    for(int b = 0; b < NBC; b++) n_valid += hist[b];
                    if(n_valid > 0) {
                        // yarray[t] = F_t, clamped to [0, 1] (neighbourhood.cpp:375-390)
                        float yarray[QF_MAX_T];
                        int cum[QF_MAX_T + 1];
                        int acc = 0;
                        for(int b = 0; b < NBC; b++) { acc += hist[b]; cum[b] = acc; }
                        for(int t = 0; t < q.T; t++) {
                            int n_le = q.rank[t] >= 0 ? cum[q.rank[t]] : 0;
                            float f = __fdiv_rn((float) n_le, (float) n_valid);
                            yarray[t] = f > 1.f ? 1.f : (f < 0.f ? 0.f : f);
                        }
                        float cq = q.quantile_field ? q.quantile_field[(size_t) y * a.nx + x] : q.quantile;
                        if(cq == 1.f && yarray[0] == 1.f) result = q.thr[0];                       // neighbourhood.cpp:396-397
                        else if(cq == 0.f && yarray[q.T - 1] == 0.f) result = q.thr[q.T - 1];      // :398-399
                        else result = qf_interpolate(cq, yarray, q.thr, q.T);                      // :400-401
                    }
                }
                obuf[rb * NT + xo] = result;
                if(xo + 1 < TX && p + 1 < SEG)
                    for(int b = 0; b < NBC; b++) hist[b] += (int) l[b * NT + xo + w] - (int) l[b * NT + xo];
            }
        }
        __syncthreads();
        {
            const int x = blockIdx.x * TX + tid;
            if(tid < TX && x < a.nx)
                for(int rb = 0; rb < nb; rb++) a.out[(size_t) (y0 + rb - a.row0) * a.nx + x] = obuf[rb * NT + tid];
        }
        __syncthreads();
    }
}

__global__ void fill_kernel(float* out, size_t n, float value) {
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) out[i] = value;
}

int check_tile(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out, int halfwidth, float* d_output) {
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");   // neighbourhood.cpp:29-30
    if(n_rows_in < 0 || nx < 0 || row0 < 0 || n_rows_out < 0 || row0 + n_rows_out > n_rows_in)
        return fail(GPP_ERR_INVALID_ARGUMENT, "row window [%d, %d) outside the %d input rows", row0, row0 + n_rows_out, n_rows_in);
    if(n_rows_out > 0 && nx > 0 && (!d_input || !d_output)) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL field pointer");
    return GPP_OK;
}

template <class K>
int opt_in_smem(K kernel, size_t bytes) {
    if(bytes > 48 * 1024) GPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    return GPP_OK;
}

}  // namespace

namespace {
__global__ void square_kernel(const float* __restrict__ in, size_t n, float* __restrict__ out) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) out[i] = __fmul_rn(in[i], in[i]);
}
// neighbourhood.cpp:221-234: float arithmetic, no clamp of a negative difference (the square root of one is NaN)
__global__ void variance_kernel(const float* __restrict__ mean, const float* __restrict__ mean2, size_t n, bool want_std, float* __restrict__ out) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float var = __fsub_rn(mean2[i], __fmul_rn(mean[i], mean[i]));
    out[i] = want_std ? __fsqrt_rn(var) : var;
}
}  // namespace

extern "C" {

int gpp_neighbourhood_device(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out, int halfwidth, int statistic,
                             float* d_output, void* stream_) {
    cudaStream_t stream = (cudaStream_t) stream_;
    GPP_TRY(check_tile(d_input, n_rows_in, nx, row0, n_rows_out, halfwidth, d_output));
    if(statistic == GPP_QUANTILE) return fail(GPP_ERR_INVALID_ARGUMENT, "Use neighbourhood_quantile for computing neighbourhood quantiles");   // :31-32
    GPP_TRY(ensure_device());
    if(n_rows_out == 0 || nx == 0) return GPP_OK;
    if(statistic == GPP_STD || statistic == GPP_VARIANCE) {
        // neighbourhood.cpp:211-235: mean2 - mean * mean from two Mean filters (of the field and of its float square)
        const size_t n_in = (size_t) n_rows_in * nx, n_out = (size_t) n_rows_out * nx;
        float *sq = nullptr, *mean = nullptr, *mean2 = nullptr;
        GPP_CUDA(cudaMallocAsync((void**) &sq, sizeof(float) * n_in, stream));
        GPP_CUDA(cudaMallocAsync((void**) &mean, sizeof(float) * n_out, stream));
        GPP_CUDA(cudaMallocAsync((void**) &mean2, sizeof(float) * n_out, stream));
        square_kernel<<<(unsigned) ((n_in + 255) / 256), 256, 0, stream>>>(d_input, n_in, sq);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        int rc = gpp_neighbourhood_device(d_input, n_rows_in, nx, row0, n_rows_out, halfwidth, GPP_MEAN, mean, stream_);
        if(rc == GPP_OK) rc = gpp_neighbourhood_device(sq, n_rows_in, nx, row0, n_rows_out, halfwidth, GPP_MEAN, mean2, stream_);
        if(rc == GPP_OK) {
            variance_kernel<<<(unsigned) ((n_out + 255) / 256), 256, 0, stream>>>(mean, mean2, n_out, statistic == GPP_STD, d_output);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        cudaFreeAsync(sq, stream);
        cudaFreeAsync(mean, stream);
        cudaFreeAsync(mean2, stream);
        if(rc == GPP_OK) GPP_CUDA(cudaGetLastError());
        return rc;
    }
    if(statistic != GPP_MEAN && statistic != GPP_SUM && statistic != GPP_COUNT && statistic != GPP_MIN && statistic != GPP_MAX)
        // neighbourhood.cpp:237-238: everything else (Median, RandomChoice) gathers the window and calls calc_statistic
        return gpp_neighbourhood_brute_force_device(d_input, n_rows_in, nx, 1, row0, n_rows_out, halfwidth, statistic, 0.f, d_output, stream_);
    TileArgs a = {d_input, d_output, n_rows_in, nx, row0, n_rows_out, halfwidth};
    const bool minmax = statistic == GPP_MIN || statistic == GPP_MAX;
    {   // copy-engine (TMA) kernels first; they decline shapes they do not cover
        int handled = 0;
        GPP_TRY(nbh_tma_try(d_input, n_rows_in, nx, row0, n_rows_out, halfwidth, statistic, d_output, stream, &handled));
        if(handled) return GPP_OK;
    }
    if(halfwidth <= HW_FUSED_MAX) {
        const int w = 2 * halfwidth + 1, TX = NT - 2 * halfwidth;
        dim3 grid((nx + TX - 1) / TX, (n_rows_out + CHUNK - 1) / CHUNK);
        if(minmax) {
            size_t smem = sizeof(float) * ((size_t) RB * NT + (size_t) RB * PADW + (size_t) (w + RB - 1) * NT);
            if(statistic == GPP_MAX) {
                GPP_TRY(opt_in_smem(nbh_minmax_kernel<true>, smem));
                GPP_LAUNCH(nbh_minmax_kernel<true>, grid, NT, smem, stream, a);
            }
            else {
                GPP_TRY(opt_in_smem(nbh_minmax_kernel<false>, smem));
                GPP_LAUNCH(nbh_minmax_kernel<false>, grid, NT, smem, stream, a);
            }
        }
        else {
            size_t smem = sizeof(SumRec) * RB * PADW + sizeof(double) * RCP_TABLE + sizeof(float) * RB * PADW + sizeof(float) * (size_t) w * NT;
            if(statistic == GPP_MEAN) { GPP_TRY(opt_in_smem(nbh_sum_kernel<0>, smem)); GPP_LAUNCH(nbh_sum_kernel<0>, grid, NT, smem, stream, a); }
            else if(statistic == GPP_SUM) { GPP_TRY(opt_in_smem(nbh_sum_kernel<1>, smem)); GPP_LAUNCH(nbh_sum_kernel<1>, grid, NT, smem, stream, a); }
            else { GPP_TRY(opt_in_smem(nbh_sum_kernel<2>, smem)); GPP_LAUNCH(nbh_sum_kernel<2>, grid, NT, smem, stream, a); }
        }
        return GPP_OK;
    }
    // large half-width: two direct separable passes through a temporary plane covering the output rows
    const size_t n_tmp = (size_t) n_rows_out * nx;
    double* tsum = nullptr;
    int* tcnt = nullptr;
    float* tval = nullptr;
    if(minmax) GPP_CUDA(cudaMallocAsync((void**) &tval, sizeof(float) * n_tmp, stream));
    else {
        GPP_CUDA(cudaMallocAsync((void**) &tsum, sizeof(double) * n_tmp, stream));
        GPP_CUDA(cudaMallocAsync((void**) &tcnt, sizeof(int) * n_tmp, stream));
    }
    unsigned blocks = (unsigned) ((n_tmp + 255) / 256);
    nbh_big_vertical_kernel<<<blocks, 256, 0, stream>>>(a, statistic, tsum, tcnt, tval, row0, n_rows_out);
    nbh_big_horizontal_kernel<<<blocks, 256, 0, stream>>>(a, statistic, tsum, tcnt, tval);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    cudaError_t err = cudaGetLastError();
    if(tsum) cudaFreeAsync(tsum, stream);
    if(tcnt) cudaFreeAsync(tcnt, stream);
    if(tval) cudaFreeAsync(tval, stream);
    if(err != cudaSuccess) return fail(GPP_ERR_CUDA, "CUDA error %s: %s", cudaGetErrorName(err), cudaGetErrorString(err));
    return GPP_OK;
}

int gpp_neighbourhood_host(const float* input, int ny, int nx, int halfwidth, int statistic, float* output) {
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");
    if(statistic == GPP_QUANTILE) return fail(GPP_ERR_INVALID_ARGUMENT, "Use neighbourhood_quantile for computing neighbourhood quantiles");
    GPP_TRY(ensure_device());
    if(ny <= 0 || nx <= 0) return GPP_OK;   // neighbourhood.cpp:33-34: empty in, empty out
    const size_t n = (size_t) ny * nx;
    DeviceBuffer<float> d_in, d_out;
    GPP_TRY(d_in.upload(input, n));
    GPP_TRY(d_out.alloc(n));
    GPP_TRY(gpp_neighbourhood_device(d_in.ptr, ny, nx, 0, ny, halfwidth, statistic, d_out.ptr, nullptr));
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_neighbourhood_quantile_fast_device(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out, float quantile,
                                           const float* d_quantile_field, int halfwidth, const float* thresholds,
                                           int num_thresholds, float* d_output, void* stream_) {
    cudaStream_t stream = (cudaStream_t) stream_;
    GPP_TRY(check_tile(d_input, n_rows_in, nx, row0, n_rows_out, halfwidth, d_output));
    if(num_thresholds < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative number of thresholds");
    // neighbourhood.cpp:316-322 (the per-pixel quantile field is validated by the host wrapper, which sees it)
    if(!d_quantile_field && is_valid(quantile) && (quantile < 0 || quantile > 1))
        return fail(GPP_ERR_INVALID_ARGUMENT, "All quantiles must be >= 0 and <= 1");
    GPP_TRY(ensure_device());
    if(n_rows_out == 0 || nx == 0) return GPP_OK;
    const size_t n_out = (size_t) n_rows_out * nx;
    if(num_thresholds == 0) {   // neighbourhood.cpp:330-331: no thresholds -> all missing
        GPP_LAUNCH(fill_kernel, (unsigned) ((n_out + 255) / 256), 256, 0, stream, d_output, n_out, NAN);
        return GPP_OK;
    }
    {   // packed-counter kernel (ascending thresholds, small half-widths); declines what it does not cover
        int handled = 0;
        GPP_TRY(qf_tma_try(d_input, n_rows_in, nx, row0, n_rows_out, quantile, d_quantile_field, halfwidth, thresholds, num_thresholds,
                           d_output, stream, &handled));
        if(handled) return GPP_OK;
    }
    if(num_thresholds > QF_MAX_T)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "neighbourhood_quantile_fast supports at most %d thresholds on the device", QF_MAX_T);
    if(halfwidth > HW_FUSED_MAX)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "neighbourhood_quantile_fast supports half-widths up to %d on the device", HW_FUSED_MAX);
    QfArgs q;
    std::memset(&q, 0, sizeof(q));
    q.t = TileArgs{d_input, d_output, n_rows_in, nx, row0, n_rows_out, halfwidth};
    q.quantile_field = d_quantile_field;
    q.quantile = quantile;
    q.T = num_thresholds;
    std::vector<std::pair<float, int> > order;
    for(int t = 0; t < num_thresholds; t++) {
        q.thr[t] = thresholds[t];
        q.rank[t] = -1;
        if(!std::isnan(thresholds[t])) order.push_back(std::make_pair(thresholds[t], t));
    }
    std::stable_sort(order.begin(), order.end(), [](const std::pair<float, int>& x, const std::pair<float, int>& y) { return x.first < y.first; });
    q.n_sorted = (int) order.size();
    for(int s = 0; s < q.n_sorted; s++) {
        q.sorted[s] = order[s].first;
        q.rank[order[s].second] = s;
    }
    const int w = 2 * halfwidth + 1, TX = NT - 2 * halfwidth, NBC = q.n_sorted + 1;
    const size_t fixed = sizeof(unsigned short) * (size_t) NBC * NT + (size_t) w * NT;
    const size_t per_row = sizeof(unsigned short) * (size_t) NBC * NT + sizeof(float) * NT;
    const size_t budget = 200 * 1024;
    int rb = (int) std::min<size_t>(RB, (budget - fixed) / per_row);
    if(rb < 1) return fail(GPP_ERR_RUNTIME, "neighbourhood_quantile_fast: shared memory budget exceeded");
    q.rb = rb;
    size_t smem = fixed + per_row * rb;
    smem = (smem + 15) / 16 * 16;
    GPP_TRY(opt_in_smem(nbh_quantile_fast_kernel, smem));
    dim3 grid((nx + TX - 1) / TX, (n_rows_out + CHUNK - 1) / CHUNK);
    GPP_LAUNCH(nbh_quantile_fast_kernel, grid, NT, smem, stream, q);
    return GPP_OK;
}

int gpp_neighbourhood_quantile_fast_host(const float* input, int ny, int nx, float quantile, const float* quantile_field,
                                         int halfwidth, const float* thresholds, int num_thresholds, float* output) {
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");   // neighbourhood.cpp:303-304
    GPP_TRY(ensure_device());
    if(ny <= 0 || nx <= 0) return GPP_OK;                                                 // :306-307
    const size_t n = (size_t) ny * nx;
    if(quantile_field) {                                                                   // :316-322
        for(size_t i = 0; i < n; i++)
            if(is_valid(quantile_field[i]) && (quantile_field[i] < 0 || quantile_field[i] > 1))
                return fail(GPP_ERR_INVALID_ARGUMENT, "All quantiles must be >= 0 and <= 1");
    }
    DeviceBuffer<float> d_in, d_out, d_q;
    GPP_TRY(d_in.upload(input, n));
    if(quantile_field) GPP_TRY(d_q.upload(quantile_field, n));
    GPP_TRY(d_out.alloc(n));
    GPP_TRY(gpp_neighbourhood_quantile_fast_device(d_in.ptr, ny, nx, 0, ny, quantile, quantile_field ? d_q.ptr : nullptr, halfwidth,
                                                   thresholds, num_thresholds, d_out.ptr, nullptr));
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

}  // extern "C"
