// neighbourhood_quantile_fast with packed counters ("K3" in SURVEY.md), input staged by the copy engine (TMA).
// Replaces gridpp::neighbourhood_quantile_fast, src/api/neighbourhood.cpp:302-409, for ascending thresholds
// (T <= 31, no NaN) and half-widths <= 15; every other case takes the general kernel in neighbourhood.cu. Row
// lengths that are not a multiple of four (or narrower than one box) run the same kernel with plain loads.
//
// The reference needs, per pixel, F_t = #(valid v <= thr_t) / #(valid v) over the clipped window for every
// threshold (one summed-area table per threshold), then inverts the CDF with gridpp::interpolate. Here:
//   * each cell is classified once: b(v) = #(thresholds < v). Its contribution to all counters is a thermometer
//     code -- field 0 (the valid count) and fields t+1 >= b+1 get a one -- looked up as 4 byte-fields per word;
//   * thread t owns window column t: the column counters C (byte fields, <= 31) slide down the rows with one packed
//     add / subtract per word; the counters of the 8 rows of a batch go to a line buffer;
//   * the horizontal pass gives every thread 8 consecutive pixels of one row: the window counters n (16-bit fields,
//     <= 961) are initialised from group-of-8 sums and then slide: n += expand(C_in - C_out), 4 thresholds per op;
//   * the CDF inversion never forms the T quotients: F_t < q <=> n_t < NLT(m) with NLT(m) = min{n : fl(n / m) >= q}
//     tabulated per valid count m (fl = the reference's float division), so the number a of thresholds below the
//     quantile is one packed compare + popcount per word; only F_{a-1} and F_a are then evaluated, and
//     gridpp::interpolate's plateau / end rules (util.cpp:339-414) are applied to them. Bit-exact with the reference.
#include "stage_ring.cuh"

#include <algorithm>
#include <cstring>

using namespace gpp;
using namespace gpp::nbh;

namespace {

constexpr int SEG = 8;
constexpr int QPF = 2;                     // float stages in flight beyond the one being classified
constexpr int NSF = QPF + 1;               // float stages in the ring
constexpr int LROW = NT + NT / 8;          // words per (row, word-plane) of the line: column c at c + c / 8
constexpr int NGRP = NT / 8;               // groups of 8 window columns
constexpr int BIN_INVALID = 32;            // row of the thermometer table with no field set
constexpr int MAX_T = 31;
constexpr int MAX_HW = 15;

struct QArgs {
    const float* in;                       // the field (plain-load form)
    float* out;
    const float* qfield;                   // may be NULL (scalar quantile)
    int n_rows_in, nx, row0, n_rows_out, hw;
    int rows_per_cta, TX, P, HL, T;
    int vec_ok;                            // rows of the output are 16-byte aligned: 32-byte vector stores
    float quantile;
    float thr[32];                         // thresholds (ascending), padded with +inf
};

__device__ __forceinline__ int pc(int c) { return c + (c >> 3); }

// NLT(m) = min{ n in [0, m] : fl(n / m) >= q }, fl = float division rounded to nearest (what the reference computes:
// float(double(n) / m) == n / m rounded once, both < 2^24). m >= 1, 0 <= q <= 1.
__device__ __forceinline__ int nlt_of(float q, int m) {
    const float fm = (float) m;
    int n = max(0, min(m, (int) (q * fm) - 2));
    while(n < m && __fdiv_rn((float) n, fm) < q) n++;
    return n;
}

// byte fields -> 16-bit fields
__device__ __forceinline__ unsigned expand_lo(unsigned x) { return __byte_perm(x, 0, 0x4140); }
__device__ __forceinline__ unsigned expand_hi(unsigned x) { return __byte_perm(x, 0, 0x4342); }

// a select the compiler cannot turn back into an indexed (local-memory) array access
__device__ __forceinline__ unsigned sel(int cond, unsigned if_true, unsigned if_false) {
    unsigned r;
    asm("{\n .reg .pred p;\n setp.ne.s32 p, %3, 0;\n selp.b32 %0, %1, %2, p;\n}" : "=r"(r) : "r"(if_true), "r"(if_false), "r"(cond));
    return r;
}
// n[idx] for a run-time idx (idx >= N yields 0): a binary tree of selects over the registers
template <int N>
__device__ __forceinline__ unsigned pick(const unsigned (&n)[N], int idx) {
    constexpr int N1 = (N + 1) / 2, N2 = (N1 + 1) / 2, N3 = (N2 + 1) / 2, N4 = (N3 + 1) / 2;
    static_assert(N4 == 1, "at most 16 words");
    unsigned t1[N1], t2[N2], t3[N3];
    #pragma unroll
    for(int j = 0; j < N1; j++) t1[j] = sel(idx & 1, 2 * j + 1 < N ? n[2 * j + 1] : 0u, n[2 * j]);
    #pragma unroll
    for(int j = 0; j < N2; j++) t2[j] = sel(idx & 2, 2 * j + 1 < N1 ? t1[2 * j + 1] : 0u, t1[2 * j]);
    #pragma unroll
    for(int j = 0; j < N3; j++) t3[j] = sel(idx & 4, 2 * j + 1 < N2 ? t2[2 * j + 1] : 0u, t2[2 * j]);
    return sel(idx & 8, N3 > 1 ? t3[N3 > 1 ? 1 : 0] : 0u, t3[0]);
}
// 16-bit field f of the packed counters (any f)
template <int NW16>
__device__ __forceinline__ int get_field(const unsigned (&n)[NW16], int f) {
    const unsigned w = pick<NW16>(n, f >> 1);
    return (int) ((f & 1) ? (w >> 16) : (w & 0xffffu));
}

// number of 16-bit fields >= c (all 2 NW16 fields; fields and c are < 2^15): bit 15 of field + 0x8000 - c
template <int NW16>
__device__ __forceinline__ int count_ge(const unsigned (&n)[NW16], int c) {
    const unsigned k2 = 0x80008000u - (unsigned) c * 0x10001u;
    int cnt = 0;
    #pragma unroll
    for(int k = 0; k + 1 < NW16; k += 2)      // the sign bytes of two words gathered into one
        cnt += __popc(__byte_perm(n[k] + k2, n[k + 1] + k2, 0x7531) & 0x80808080u);
    if(NW16 & 1) cnt += __popc((n[NW16 - 1] + k2) & 0x80008000u);
    return cnt;
}

// fl(n / m) for integers 0 <= n <= m <= 961 from the correctly rounded reciprocal r = fl(1 / m): Markstein's
// residual correction. Equal to __fdiv_rn((float) n, (float) m) for every such pair (exhaustively checked in
// tests/test_capi_cpu.py::test_small_integer_quotients_are_exact).
__device__ __forceinline__ float quot(int n, float fm, float r) {
    const float fn = (float) n;
    const float q0 = __fmul_rn(fn, r);
    return __fmaf_rn(__fmaf_rn(-q0, fm, fn), r, q0);
}

// The CDF inversion for one pixel. n: window counters (field 0 = valid count m, field t + 1 = #(v <= thr_t), pad
// fields = m). neighbourhood.cpp:374-401 and gridpp::interpolate, util.cpp:377-414.
template <int NW16>
__device__ __forceinline__ float invert_cdf(const unsigned (&n)[NW16], float q, int nlt, float rm, int T, const float* thr) {
    const int m = (int) (n[0] & 0xffffu);
    const float fm = (float) m;
    // a = #(t : F_t < q): the fields below NLT are threshold fields (field 0 and the pads equal m >= NLT)
    const int a = 2 * NW16 - count_ge<NW16>(n, nlt);
    // fields a (F_{a-1}) and a + 1 (F_a) are adjacent 16-bit fields
    const unsigned w0 = pick<NW16>(n, a >> 1), w1 = pick<NW16>(n, (a >> 1) + 1);
    const unsigned pair = (a & 1) ? __funnelshift_r(w0, w1, 16) : w0;
    const int lo = (int) (pair & 0xffffu), hi = (int) (pair >> 16);
    if(q == 1.f && (int) (n[0] >> 16) == m) return thr[0];                            // neighbourhood.cpp:396-397
    if(q == 0.f && get_field<NW16>(n, T) == 0) return thr[T - 1];                     // :398-399
    if(a == T) return thr[T - 1];                                                     // util.cpp:386-387: x > iX.back()
    const float f_hi = quot(hi, fm, rm);
    if(f_hi == q) {
        // plateau at q: lower index = first F == q = a, upper index = last F == q (util.cpp:339-376,394-403)
        int nle = hi;
        while(nle < m && quot(nle + 1, fm, rm) == q) nle++;
        int b = 2 * NW16 - count_ge<NW16>(n, nle + 1);          // fields <= nle ...
        if(m <= nle) b -= 2 * NW16 - T;                         // ... without field 0 and the pads
        const int i1 = b - 1;
        const float y0 = thr[a], y1 = thr[i1];
        if(a == 0 && i1 == T - 1) return __fdiv_rn(__fadd_rn(y0, y1), 2.f);
        if(a == 0) return y1;
        if(i1 == T - 1) return y0;
        return __fdiv_rn(__fadd_rn(y0, y1), 2.f);
    }
    if(a == 0) return thr[0];                                                         // util.cpp:388-389: x < iX[0]
    const float f_lo = quot(lo, fm, rm);
    const float y0 = thr[a - 1], y1 = thr[a];
    // y0 + (y1 - y0) * (x - x0) / (x1 - x0), float, left to right (util.cpp:410)
    return __fadd_rn(y0, __fdiv_rn(__fmul_rn(__fsub_rn(y1, y0), __fsub_rn(q, f_lo)), __fsub_rn(f_hi, f_lo)));
}

// NW8 words of 4 byte fields cover fields 0 .. 4 NW8 - 1 (field 0 = valid, 1 .. T = thresholds, the rest pads)
// TMA: the input rows are staged by the copy engine (row length a multiple of 4); otherwise every thread loads its
// column's 8 values of a stage with plain coalesced loads (any row length, narrow fields).
template <int NW8, bool QFIELD, bool TMA>
__global__ void __launch_bounds__(NT, 2) qf_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ QArgs a) {
    constexpr int NW16 = 2 * NW8;
    extern __shared__ __align__(128) unsigned char smem[];
    const int hw = a.hw, w = 2 * hw + 1, P = a.P, T = a.T;
    const int NRB = RB * (P + 1);                                    // rows of the bin ring
    float* fring = reinterpret_cast<float*>(smem);                                   // [NSF][RB][NT] landed floats
    unsigned* line = reinterpret_cast<unsigned*>(fring + NSF * RB * NT);             // [RB][NW8][LROW] column counters
    unsigned* grp = line + RB * NW8 * LROW;                                          // [RB][NW16][NGRP] group sums
    unsigned* therm = grp + RB * NW16 * NGRP;                                        // [NW8][33] thermometer codes
    float* sthr = reinterpret_cast<float*>(therm + NW8 * 34);                        // [32] (34: keeps the barriers 8-byte aligned)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sthr + 32);     // [NSF]
    float2* mtab = reinterpret_cast<float2*>(bars + NSF);                            // [w * w + 1]: (NLT(m) as int bits, fl(1 / m))
    unsigned char* bring = reinterpret_cast<unsigned char*>(mtab + ((w * w + 1 + 1) & ~1));    // [NRB][NT] bins

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * a.TX;
    const int y_begin = a.row0 + blockIdx.y * a.rows_per_cta;
    const int y_end = min(y_begin + a.rows_per_cta, a.row0 + a.n_rows_out);
    const int n_batches = (y_end - y_begin + RB - 1) / RB;
    const StageRing R = {fring, bars, &in_map, x0 - a.HL, y_begin + hw - RB * P, NSF, P + n_batches};
    if(TMA) R.start();
    // ---- tables
    if(tid < 32) sthr[tid] = a.thr[tid];
    for(int e = tid; e < NW8 * 33; e += NT) {
        const int k = e / 33, b = e % 33;
        unsigned word = 0;
        if(b != BIN_INVALID)
            for(int j = 0; j < 4; j++) {
                const int f = 4 * k + j;                     // field 0: valid; 1..T: v <= thr_{f-1} <=> f >= b + 1; pads: valid
                if(f == 0 || f > T || f >= b + 1) word |= 1u << (8 * j);
            }
        therm[e] = word;
    }
    for(int m = tid; m <= w * w; m += NT)
        mtab[m] = make_float2(__int_as_float(!QFIELD && m > 0 ? nlt_of(a.quantile, m) : 0), m > 0 ? __frcp_rn((float) m) : 0.f);
    // window column tid of the strip = staged column tid + (HL - hw), see nbh_sum_tma_kernel
    const int scol = min(tid + (a.HL - hw), NT - 1);
    const bool has_col = tid + (a.HL - hw) < NT;   // the last HL - hw threads have no window column: they must not touch a shared one
    unsigned char* const my_bins = bring + scol;
    unsigned* const my_line = line + pc(tid);
    const int rel0 = RB * P - 2 * hw;        // ring row of input row y_begin - hw (earlier rows of stage 0 are not used)
    const int hb = tid >> 5, seg = tid & 31, xo0 = seg * SEG;
    const bool h_active = xo0 < a.TX && x0 + xo0 < a.nx;
    const int full = w >> 3, rem = w & 7;
    unsigned C[NW8];
    #pragma unroll
    for(int k = 0; k < NW8; k++) C[k] = 0;
    __syncthreads();

    int brow = 0;                            // bin-ring row of the first row of the current stage
    for(int s = 0; s < P + n_batches; s++) {
        float vals[RB];
        if(TMA) {
            R.wait(s);
            const float* sp = fring + (size_t) (s % NSF) * RB * NT + scol;
            #pragma unroll
            for(int b = 0; b < RB; b++) vals[b] = sp[b * NT];
        }
        else {
            const int xg = x0 - a.HL + scol;                        // field column of this thread
            const int rg = y_begin + hw - RB * P + RB * s;          // field row of the stage's first row
            const bool col_ok = xg >= 0 && xg < a.nx;
            #pragma unroll
            for(int b = 0; b < RB; b++)
                vals[b] = (col_ok && rg + b >= 0 && rg + b < a.n_rows_in) ? __ldg(a.in + (size_t) (rg + b) * a.nx + xg) : NAN;
        }
        const bool batch = s >= P;
        // bin-ring row of the row that leaves when row b of this stage enters: (8 s + b) - w + ... = the row 2 hw + 1
        // rows older; as an index: brow + b - w (mod NRB). NRB = 8 (P + 1) >= w + 7.
        int orow = brow - w + 1;             // row leaving after output row b = 0 (the window keeps 2 hw + 1 rows)
        if(orow < 0) orow += NRB;
        // ---- vertical pass: classify the 8 landed values, slide the column counters, record them
        #pragma unroll
        for(int b = 0; b < RB; b++) {
            const float v = vals[b];
            const float* tp = sthr;                       // lower bound: tp - sthr = #(thresholds < v)
            #pragma unroll
            for(int step = 16; step > 0; step >>= 1)
                if(tp[step - 1] < v) tp += step;
            int bin = (int) (tp - sthr);
            if(!finite_f(v)) bin = BIN_INVALID;
            const bool used = has_col && (batch || RB * s + b >= rel0);
            if(used) {
                my_bins[(brow + b) * NT] = (unsigned char) bin;
                #pragma unroll
                for(int k = 0; k < NW8; k++) C[k] += therm[k * 33 + bin];
            }
            if(batch) {
                #pragma unroll
                for(int k = 0; k < NW8; k++) my_line[(b * NW8 + k) * LROW] = C[k];
                const int bo = has_col ? my_bins[orow * NT] : BIN_INVALID;
                #pragma unroll
                for(int k = 0; k < NW8; k++) C[k] -= therm[k * 33 + bo];
                orow = orow + 1 == NRB ? 0 : orow + 1;
            }
        }
        brow += RB;
        if(brow == NRB) brow = 0;
        __syncthreads();
        if(TMA) R.recycle(s);                // the float stage has been classified
        if(!batch) continue;
        const int y0 = y_begin + RB * (s - P);
        // ---- group sums: thread (row hb, group seg) adds the 8 columns of its group (bytes, <= 248) and expands
        {
            const unsigned* l = line + (hb * NW8) * LROW + 9 * seg;
            unsigned* g = grp + (hb * NW16) * NGRP + seg;
            #pragma unroll
            for(int k = 0; k < NW8; k++) {
                unsigned sum = 0;
                #pragma unroll
                for(int j = 0; j < 8; j++) sum += l[k * LROW + j];
                g[(2 * k) * NGRP] = expand_lo(sum);
                g[(2 * k + 1) * NGRP] = expand_hi(sum);
            }
        }
        __syncthreads();
        // ---- horizontal pass: 8 consecutive pixels of row y0 + hb
        const int y = y0 + hb;
        if(h_active && y < y_end) {
            const unsigned* l = line + (hb * NW8) * LROW + 9 * seg;       // window column xo0 + j at l[j + j / 8]
            const unsigned* g = grp + (hb * NW16) * NGRP + seg;
            unsigned n[NW16];
            #pragma unroll
            for(int k = 0; k < NW16; k++) n[k] = 0;
            for(int mgrp = 0; mgrp < full; mgrp++) {
                #pragma unroll
                for(int k = 0; k < NW16; k++) n[k] += g[k * NGRP + mgrp];
            }
            if(rem <= 4) {
                for(int j = 8 * full; j < 8 * full + rem; j++) {
                    #pragma unroll
                    for(int k = 0; k < NW8; k++) {
                        const unsigned c = l[k * LROW + pc(j)];
                        n[2 * k] += expand_lo(c);
                        n[2 * k + 1] += expand_hi(c);
                    }
                }
            }
            else {
                #pragma unroll
                for(int k = 0; k < NW16; k++) n[k] += g[k * NGRP + full];
                for(int j = 8 * full + rem; j < 8 * full + 8; j++) {
                    #pragma unroll
                    for(int k = 0; k < NW8; k++) {
                        const unsigned c = l[k * LROW + pc(j)];
                        n[2 * k] -= expand_lo(c);
                        n[2 * k + 1] -= expand_hi(c);
                    }
                }
            }
            const int x = x0 + xo0;
            float qv[SEG];
            if(QFIELD) {
                const float* qp = a.qfield + (size_t) y * a.nx + x;
                #pragma unroll
                for(int p = 0; p < SEG; p++) qv[p] = x + p < a.nx ? qp[p] : NAN;
            }
            float o[SEG];
            #pragma unroll
            for(int p = 0; p < SEG; p++) {
                const float q = QFIELD ? qv[p] : a.quantile;
                const int m = (int) (n[0] & 0xffffu);
                float r = NAN;                                           // no valid value / invalid quantile: missing
                if(m > 0 && finite_f(q)) {
                    const float2 e = mtab[m];
                    const int nlt = QFIELD ? nlt_of(q, m) : __float_as_int(e.x);
                    r = invert_cdf<NW16>(n, q, nlt, e.y, T, sthr);
                }
                o[p] = r;
                if(p + 1 < SEG) {
                    const int jin = pc(w + p);
                    #pragma unroll
                    for(int k = 0; k < NW8; k++) {
                        // per byte: C_in - C_out + 0x40 in [33, 95]: no borrow between the fields
                        const unsigned d = l[k * LROW + jin] - l[k * LROW + p] + 0x40404040u;
                        n[2 * k] += expand_lo(d) - 0x00400040u;
                        n[2 * k + 1] += expand_hi(d) - 0x00400040u;
                    }
                }
            }
            float* dst = a.out + (size_t) (y - a.row0) * a.nx + x;
            if(a.vec_ok && x + SEG <= a.nx) {
                reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
            else {
                #pragma unroll
                for(int p = 0; p < SEG; p++)
                    if(x + p < a.nx) dst[p] = o[p];
            }
        }
        __syncthreads();   // line / grp are rewritten by the next batch
    }
}

template <int NW8>
int launch_qf(bool qfield, bool tma, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& map, const QArgs& a) {
    auto kernel = tma ? (qfield ? qf_tma_kernel<NW8, true, true> : qf_tma_kernel<NW8, false, true>)
                      : (qfield ? qf_tma_kernel<NW8, true, false> : qf_tma_kernel<NW8, false, false>);
    GPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    GPP_LAUNCH(kernel, grid, NT, smem, stream, map, a);
    return GPP_OK;
}

}  // namespace

namespace gpp {

// *handled = 0 -> the caller runs the general kernel.
int qf_tma_try(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out, float quantile, const float* d_quantile_field,
               int hw, const float* thresholds, int T, float* d_output, cudaStream_t stream, int* handled) {
    *handled = 0;
    if(getenv("GPP_NO_TMA")) return GPP_OK;
    if(hw < 1 || hw > MAX_HW || T < 1 || T > MAX_T) return GPP_OK;
    // the copy engine needs 16-byte row pitches and a field at least one box wide; other shapes load directly
    const bool tma = nx % 4 == 0 && nx >= NT && ((uintptr_t) d_input & 15) == 0;
    for(int t = 0; t < T; t++) {
        if(std::isnan(thresholds[t])) return GPP_OK;
        if(t > 0 && thresholds[t] < thresholds[t - 1]) return GPP_OK;
    }
    if(!d_quantile_field && std::isnan(quantile)) return GPP_OK;      // all-missing output: the general kernel's job
    QArgs a;
    std::memset(&a, 0, sizeof(a));
    a.in = d_input;
    a.out = d_output;
    a.vec_ok = nx % 4 == 0 && ((uintptr_t) d_output & 15) == 0;
    a.qfield = d_quantile_field;
    a.n_rows_in = n_rows_in; a.nx = nx; a.row0 = row0; a.n_rows_out = n_rows_out; a.hw = hw;
    a.HL = (hw + 3) / 4 * 4;
    a.TX = (NT - a.HL - hw) / SEG * SEG;
    a.P = (2 * hw + RB - 1) / RB;
    a.T = T;
    a.quantile = quantile;
    for(int t = 0; t < 32; t++) a.thr[t] = t < T ? thresholds[t] : INFINITY;
    const int nw8 = (T + 1 + 3) / 4;
    const int NW8 = nw8 <= 4 ? nw8 : (nw8 <= 6 ? 6 : 8);
    const int w = 2 * hw + 1;
    size_t smem = (size_t) NSF * STAGE_BYTES + sizeof(unsigned) * ((size_t) RB * NW8 * LROW + (size_t) RB * 2 * NW8 * NGRP + NW8 * 34) +
                  sizeof(float) * 32 + sizeof(unsigned long long) * NSF + sizeof(float2) * ((w * w + 1 + 1) & ~1) +
                  (size_t) RB * (a.P + 1) * NT;
    if(smem > 200 * 1024) return GPP_OK;
    const int strips = (nx + a.TX - 1) / a.TX;
    const int per_sm = std::max(1, std::min(2, (int) ((227 * 1024) / (smem + 1024))));
    const int slots = sm_count() * per_sm;
    int chunks = std::max(1, slots / strips);
    int rows = (n_rows_out + chunks - 1) / chunks;
    rows = std::max(4 * RB, (rows + RB - 1) / RB * RB);
    a.rows_per_cta = rows;
    chunks = (n_rows_out + rows - 1) / rows;
    CUtensorMap map;
    std::memset(&map, 0, sizeof(map));
    if(tma) GPP_TRY(make_field_tensor_map(&map, d_input, n_rows_in, nx, RB, NT, true));
    dim3 grid(strips, chunks);
    const bool qf = d_quantile_field != nullptr;
    switch(NW8) {
        case 1: GPP_TRY(launch_qf<1>(qf, tma, grid, smem, stream, map, a)); break;
        case 2: GPP_TRY(launch_qf<2>(qf, tma, grid, smem, stream, map, a)); break;
        case 3: GPP_TRY(launch_qf<3>(qf, tma, grid, smem, stream, map, a)); break;
        case 4: GPP_TRY(launch_qf<4>(qf, tma, grid, smem, stream, map, a)); break;
        case 6: GPP_TRY(launch_qf<6>(qf, tma, grid, smem, stream, map, a)); break;
        default: GPP_TRY(launch_qf<8>(qf, tma, grid, smem, stream, map, a)); break;
    }
    *handled = 1;
    return GPP_OK;
}

}  // namespace gpp
