// Neighbourhood Mean / Sum / Count / Min / Max with the input staged by the bulk-tensor copy engine (TMA).
// Replaces gridpp::neighbourhood, src/api/neighbourhood.cpp:28-242, for fields whose row length is a multiple of
// four (the copy engine needs 16-byte row pitches); other shapes take the kernels in neighbourhood.cu.
//
// A CTA of 256 threads owns a strip of 256 staged columns (TX output columns plus `hw` halo columns on each side)
// and walks down a chunk of output rows. One elected thread keeps a ring of NS stages of 8 rows x 256 columns in
// flight: each stage is ONE cp.async.bulk.tensor copy that lands in shared memory and completes an mbarrier; cells
// outside the field arrive as zero (sums) or NaN (extremes), which is how the window gets clipped at the domain
// edges (neighbourhood.cpp:104-107) without a special case. Thread t owns staged column t:
//   vertical pass    per row: the entering value is added to the column's window state, the leaving one (still in
//                    the ring, 2 hw + 1 rows older) removed; the state of each of the 8 rows of a batch is written
//                    to a line buffer;
//   horizontal pass  each thread slides along 8 consecutive pixels of one row of the batch and stores 32 bytes.
// HBM traffic is the algorithmic 8 B/pixel plus the halo rows/columns of each CTA (L2 hits for the most part).
#include "stage_ring.cuh"

#include <algorithm>
#include <cstring>

using namespace gpp;
using namespace gpp::nbh;

namespace gpp {

int make_field_tensor_map(CUtensorMap* map, const float* base, int rows, int nx, int box_rows, int box_cols, bool nan_fill) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if(!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if(err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return fail(GPP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        }
        encode = (EncodeFn) fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t) nx, (cuuint64_t) rows};
    cuuint64_t strides[1] = {(cuuint64_t) nx * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t) box_cols, (cuuint32_t) box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if(rc != CUDA_SUCCESS) return fail(GPP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int) rc);
    return GPP_OK;
}

}  // namespace gpp

namespace {

constexpr int SEG = 8;               // consecutive pixels per thread in the horizontal pass
#ifndef NBH_PF
#define NBH_PF 2
#endif
#ifndef NBH_MINB_SUM
#define NBH_MINB_SUM 3
#endif
#ifndef NBH_MINB_MM
#define NBH_MINB_MM 4
#endif
#ifndef NBH_PF_MM
#define NBH_PF_MM NBH_PF
#endif
constexpr int PREFETCH = NBH_PF;     // stages in flight beyond the ones the window needs (sums)
constexpr int PREFETCH_MM = NBH_PF_MM;   // ... (extremes)
constexpr int MINB_SUM = NBH_MINB_SUM, MINB_MM = NBH_MINB_MM;   // resident CTAs per SM the kernels are built for
constexpr int RCP_MAX = 1024;
// Line buffers hold one record per window column and batch row. The vertical pass writes them with consecutive
// lanes on consecutive columns, the horizontal pass reads them with lane l starting at column 8 l.
//  * kernels that hold the row window in registers read 16-byte vectors; an XOR swizzle of the 16-byte slots inside
//    each 128-byte row makes both accesses conflict-free: slot s = 8 r + c is stored at 8 r + (c ^ (r & 7));
//  * the sliding kernels (large or run-time half-widths) read single records at run-time offsets, where a padded
//    layout is cheaper to address: 8-byte records get 2 pad records per 8, 4-byte records 4 per 8.
constexpr int LROW_D = NT + NT / 4;  // doubles per line row (padded form; the swizzled form uses the first NT)
constexpr int LROW_F = NT + NT / 2;  // floats per line row
__device__ __forceinline__ int swz_slot(int s) { return (s & ~7) | ((s ^ (s >> 3)) & 7); }
template <bool SWZ> __device__ __forceinline__ int idx_d(int e) { return SWZ ? ((swz_slot(e >> 1) << 1) | (e & 1)) : e + 2 * (e >> 3); }
template <bool SWZ> __device__ __forceinline__ int idx_f(int e) { return SWZ ? ((swz_slot(e >> 2) << 2) | (e & 3)) : e + 4 * (e >> 3); }

struct TmaArgs {
    float* out;
    int n_rows_in, nx, row0, n_rows_out, hw;
    int rows_per_cta;    // output rows per CTA, a multiple of RB
    int TX;              // output columns per strip, a multiple of SEG
    int P;               // stages that must have landed before the first output row: ceil(2 hw / RB)
    int NS;              // stages in the ring: P + 1 + PREFETCH
    int n_rcp;           // entries of the reciprocal table
    int HL;              // staged columns to the left of the strip: hw rounded up to a multiple of 4 (the copy engine
                         // needs 16-byte aligned box origins); the first HL - hw of them are not part of any window
};

__device__ __forceinline__ void store_segment(const TmaArgs& a, int y, int x, const float (&o)[SEG]) {
    float* dst = a.out + (size_t) (y - a.row0) * a.nx + x;
    if(x + SEG <= a.nx) {
        reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    else {
        #pragma unroll
        for(int p = 0; p < SEG; p++)
            if(x + p < a.nx) dst[p] = o[p];
    }
}

// ------------------------------------------------------------------ mean / sum / count ----------------
// neighbourhood.cpp:45-145. The reference takes four corners of a double summed-area table and of an int count
// table; here the clipped-window sum is accumulated directly in fp64 (a running column sum that never holds more
// than 2 hw + 1 values, then a row-window sum over 2 hw + 1 columns) and divided by the valid count.
//
// The copy engine fills cells outside the field with ZERO for these kernels, so a clipped window sums correctly by
// itself and its valid count is (rows inside) x (columns inside) as long as no cell of the window is missing.
// Missing values (NaN, +-inf) are handled by poisoning: the running column sum is updated without any test; a
// non-finite value makes it non-finite for good (inf - inf = NaN), which is checked once per 8-row batch. Only then
// is the column rebuilt from the rows still in the ring, this time counting the invalid cells, and the CTA takes the
// horizontal pass that slides the valid counts along with the sums.
struct ColumnState {
    double csum;   // sum of the valid values of rows y0 - hw .. y0 + hw - 1 of the column (the next batch's start)
    int ninv;      // invalid (non-finite) values among them
};

// Rebuild (poisoned) / annotate (clean) the records of one column for the batch whose first window starts at ring
// row `first`. Records: line[b] = sum of the valid values of the window of output row y0 + b, cline[b] = their count.
__device__ __noinline__ void fix_column(const float* ring_col, int NR, int first, int w, bool poisoned, ColumnState& st,
                                        double* my_line, unsigned char* my_cline, int y0, int hw, int n_rows_in, bool col_ok) {
    if(poisoned) {
        double s = 0.0;
        int nv = 0, slot = first;
        for(int j = 0; j < w - 1; j++) {
            const float v = ring_col[slot * NT];
            if(finite_f(v)) s += (double) v; else nv++;
            slot = slot + 1 == NR ? 0 : slot + 1;
        }
        int lead = first;
        for(int b = 0; b < RB; b++) {
            const float vn = ring_col[slot * NT];
            if(finite_f(vn)) s += (double) vn; else nv++;
            slot = slot + 1 == NR ? 0 : slot + 1;
            const int y = y0 + b;
            const int ch = min(y + hw, n_rows_in - 1) - max(y - hw, 0) + 1;
            my_line[b * LROW_D] = s;
            my_cline[b * LROW_D] = (unsigned char) (col_ok ? max(ch - nv, 0) : 0);
            const float vo = ring_col[lead * NT];
            if(finite_f(vo)) s -= (double) vo; else nv--;
            lead = lead + 1 == NR ? 0 : lead + 1;
        }
        st.csum = s;
        st.ninv = nv;
    }
    else {
        for(int b = 0; b < RB; b++) {
            const int y = y0 + b;
            const int ch = min(y + hw, n_rows_in - 1) - max(y - hw, 0) + 1;
            my_cline[b * LROW_D] = (unsigned char) (col_ok ? max(ch - st.ninv, 0) : 0);
        }
    }
}

// value / count as the reference forms it (double division rounded to float, neighbourhood.cpp:133-142); counts
// below n_rcp multiply by a tabulated reciprocal instead (differs from the division by at most one rounding of the
// double quotient, invisible after the rounding to float but for ~1e-9 of the cases)
template <bool TABLE_ONLY>
__device__ __forceinline__ float mean_of(double s, int c, const double* rcp, int n_rcp) {
    if(TABLE_ONLY || c < n_rcp) return (float) (s * rcp[c]);
    return (float) (s / (double) c);
}

template <int N>
__device__ __forceinline__ double tree_sum(const double* v) {
    if constexpr(N == 1) return v[0];
    else return tree_sum<N / 2>(v) + tree_sum<N - N / 2>(v + N / 2);
}

// STAT: 0 = Mean, 1 = Sum, 2 = Count.  HW > 0: half-width known at compile time (loops unrolled, the row window held
// in registers); HW == 0: any half-width (a.hw).
template <int STAT, int HW>
__global__ void __launch_bounds__(NT, MINB_SUM) nbh_sum_tma_kernel(const __grid_constant__ CUtensorMap in_map, const TmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool STATIC = HW > 0;
    constexpr bool TAB = STATIC && (2 * HW + 1) * (2 * HW + 1) < RCP_MAX;   // every possible count is tabulated
    constexpr bool REGWIN = STATIC && HW <= 8;                               // row window in registers, swizzled line
    const int hw = STATIC ? HW : a.hw, w = 2 * hw + 1;
    const int P = STATIC ? (2 * HW + RB - 1) / RB : a.P;
    const int NS = STATIC ? P + 1 + PREFETCH : a.NS;
    const int NR = NS * RB;
    float* ring = reinterpret_cast<float*>(smem);                                 // [NS * RB][NT]
    double* line = reinterpret_cast<double*>(ring + (size_t) NR * NT);           // [RB][LROW_D] column sums
    double* rcp = line + RB * LROW_D;                                             // [n_rcp]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(rcp + a.n_rcp);   // [NS]
    unsigned char* cline = reinterpret_cast<unsigned char*>(bars + NS);           // [RB][LROW_D] column valid counts (<= w)

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * a.TX;
    const int y_begin = a.row0 + blockIdx.y * a.rows_per_cta;
    const int y_end = min(y_begin + a.rows_per_cta, a.row0 + a.n_rows_out);
    const int n_batches = (y_end - y_begin + RB - 1) / RB;
    const StageRing R = {ring, bars, &in_map, x0 - a.HL, y_begin + hw - RB * P, NS, P + n_batches};
    R.start();
    for(int c = tid; c < a.n_rcp; c += NT) rcp[c] = c > 0 ? 1.0 / (double) c : 0.0;

    // thread t owns window column t of the strip = staged column t + (HL - hw); the HL - hw rightmost threads have
    // no column (they re-read the last one; their records are never used)
    const int scol = min(tid + (a.HL - hw), NT - 1);
    const int x_stage = x0 - a.HL + scol;
    const bool col_ok = x_stage >= 0 && x_stage < a.nx;
    double* const my_line = line + idx_d<REGWIN>(tid);
    unsigned char* const my_cline = cline + tid;
    const float* const ring_col = ring + scol;
    const int rel0 = RB * P - 2 * hw;     // ring row of input row y_begin - hw; 0 <= rel0 < RB
    // ---- prime the column with rows y_begin - hw .. y_begin + hw - 1
    ColumnState st = {0.0, 0};
    for(int k = 0; k < P; k++) R.wait(k);
    for(int rel = rel0; rel < RB * P; rel++) {
        const float v = ring_col[rel * NT];
        if(finite_f(v)) st.csum += (double) v; else st.ninv++;
    }
    // register-window form: rows y0 - hw .. y0 + hw - 1 of the column, already converted (each value is converted to
    // double once: the conversions run on the 16-lane XU pipe)
    double hist[REGWIN ? 2 * HW : 1];
    if constexpr(REGWIN) {
        #pragma unroll
        for(int j = 0; j < 2 * HW; j++) hist[j] = (double) ring_col[(rel0 + j) * NT];
    }
    const int hb = tid >> 5, seg = tid & 31, xo0 = seg * SEG;
    const bool h_active = xo0 < a.TX && x0 + xo0 < a.nx;
    const bool strip_inside = x0 - hw >= 0 && x0 + a.TX - 1 + hw < a.nx;   // no window of the strip is clipped sideways
    const double rc_full = 1.0 / (double) (w * w);
    int s_new = P % NS, s_old = 0;        // ring slots of the entering stage and of the stage of the first leaving row

    for(int i = 0; i < n_batches; i++) {
        const int y0 = y_begin + RB * i;
        R.wait(P + i);
        // ---- vertical pass: row y0 + b + hw enters, the record of output row y0 + b is written, row y0 + b - hw leaves
        {
            const float* newp = ring_col + s_new * (RB * NT);
            double dn[RB], dd[RB], c[RB + 1];
            if constexpr(REGWIN) {
                // the 2 hw rows above the entering stage are kept in registers: the leaving rows are hist[0 .. 7]
                #pragma unroll
                for(int b = 0; b < RB; b++) dn[b] = (double) newp[b * NT];
                #pragma unroll
                for(int b = 0; b < RB; b++) dd[b] = dn[b] - (b < 2 * HW ? hist[b] : dn[b - 2 * HW]);
                #pragma unroll
                for(int j = 0; j < 2 * HW; j++) hist[j] = j + RB < 2 * HW ? hist[j + RB] : dn[j + RB - 2 * HW];
            }
            else {
                const int s_old2 = s_old + 1 == NS ? 0 : s_old + 1;
                const float* oldA = ring_col + (s_old * RB + rel0) * NT;           // leaving rows b < RB - rel0
                const float* oldB = ring_col + (s_old2 * RB + rel0 - RB) * NT;     // leaving rows b >= RB - rel0
                #pragma unroll
                for(int b = 0; b < RB; b++) {
                    dn[b] = (double) newp[b * NT];
                    dd[b] = dn[b] - (double) (b < RB - rel0 ? oldA[b * NT] : oldB[b * NT]);
                }
            }
            c[0] = st.csum;
            #pragma unroll
            for(int b = 0; b < RB; b += 2) {
                c[b + 1] = c[b] + dd[b];
                c[b + 2] = c[b] + (dd[b] + dd[b + 1]);
            }
            #pragma unroll
            for(int b = 0; b < RB; b++) my_line[b * LROW_D] = c[b] + dn[b];
            st.csum = c[RB];
        }
        const bool poisoned = !(fabs(st.csum) < INFINITY);
        const bool any_missing = __syncthreads_or(poisoned || st.ninv > 0) != 0;
        if(any_missing) {
            int first = s_old * RB + rel0;
            // through a copy: passing `st` itself by reference would keep it in local memory in the fast path too
            // (one LDL/STL pair per batch at the head of the dependent chain; mean hw 7: 45.8 -> 42.8 us)
            ColumnState fixed = st;
            fix_column(ring_col, NR, first, w, poisoned, fixed, my_line, my_cline, y0, hw, a.n_rows_in, col_ok);
            st = fixed;
            __syncthreads();
        }
        R.recycle(i);          // stage i is dead: every leaving row of later batches lives in a later stage
        s_new = s_new + 1 == NS ? 0 : s_new + 1;
        s_old = s_old + 1 == NS ? 0 : s_old + 1;
        // ---- horizontal pass
        const int y = y0 + hb;
        if(h_active && y < y_end) {
            const int x = x0 + xo0;
            const double* l = line + hb * LROW_D;     // window column e at l[idx_d(e)]
            const unsigned char* lc = cline + hb * LROW_D + xo0;
            float o[SEG];
            const int ch = min(y + hw, a.n_rows_in - 1) - max(y - hw, 0) + 1;   // rows of the window inside the field
            const bool full = strip_inside && ch == w;
            if constexpr(REGWIN) {
                constexpr int W = 2 * HW + 1, NV = W + SEG - 1;
                double v[NV + 1];
                #pragma unroll
                for(int q = 0; q < (NV + 1) / 2; q++) {
                    const double2 t = reinterpret_cast<const double2*>(l)[swz_slot(4 * seg + q)];
                    v[2 * q] = t.x;
                    v[2 * q + 1] = t.y;
                }
                // window sums of the 8 pixels: sw[0] by a pairwise tree, sw[p+1] = sw[p] + (v[W+p] - v[p]) two at a time
                double sw[SEG], t[SEG - 1];
                sw[0] = tree_sum<W>(v);
                #pragma unroll
                for(int p = 0; p < SEG - 1; p++) t[p] = v[W + p] - v[p];
                #pragma unroll
                for(int p = 0; p + 1 < SEG; p += 2) {
                    sw[p + 1] = sw[p] + t[p];
                    if(p + 2 < SEG) sw[p + 2] = sw[p] + (t[p] + t[p + 1]);
                }
                if(!any_missing) {
                    #pragma unroll
                    for(int p = 0; p < SEG; p++) {
                        const double s = sw[p];
                        if(STAT == 1) o[p] = (float) s;
                        else if(full) o[p] = STAT == 2 ? (float) (W * W) : (float) (s * rc_full);
                        else {
                            // count = (rows inside the field) x (columns inside the field), neighbourhood.cpp:104-107
                            const int c = ch * (min(x + p + hw, a.nx - 1) - max(x + p - hw, 0) + 1);
                            if(STAT == 2) o[p] = (float) c;
                            else o[p] = mean_of<TAB>(s, c, rcp, a.n_rcp);
                        }
                    }
                }
                else {
                    int cv[NV];
                    #pragma unroll
                    for(int j = 0; j < NV; j++) cv[j] = lc[j];
                    int c = 0;
                    #pragma unroll
                    for(int j = 0; j < W; j++) c += cv[j];
                    #pragma unroll
                    for(int p = 0; p < SEG; p++) {
                        const double s = sw[p];
                        if(STAT == 2) o[p] = (float) c;
                        else if(STAT == 1) o[p] = c > 0 ? (float) s : NAN;
                        else o[p] = c > 0 ? mean_of<TAB>(s, c, rcp, a.n_rcp) : NAN;
                        if(p + 1 < SEG) c += cv[W + p] - cv[p];
                    }
                }
            }
            else {
                // sliding sums, the records re-read from the (padded) line as they enter and leave
                const double* lp = l + seg * (SEG + 2);   // window column xo0 + j at lp[j + 2 (j / 8)]
                double s = 0.0;
                int c = 0;
                #pragma unroll 8
                for(int j = 0; j < w; j++) s += lp[idx_d<false>(j)];
                if(any_missing)
                    for(int j = 0; j < w; j++) c += lc[j];
                #pragma unroll
                for(int p = 0; p < SEG; p++) {
                    if(!any_missing && full) o[p] = STAT == 2 ? (float) (w * w) : (STAT == 1 ? (float) s : (float) (s * rc_full));
                    else {
                        const int cc = any_missing ? c : ch * (min(x + p + hw, a.nx - 1) - max(x + p - hw, 0) + 1);
                        if(STAT == 2) o[p] = (float) cc;
                        else if(STAT == 1) o[p] = cc > 0 ? (float) s : NAN;
                        else o[p] = cc > 0 ? mean_of<TAB>(s, cc, rcp, a.n_rcp) : NAN;
                    }
                    if(p + 1 < SEG) {
                        s += lp[idx_d<false>(w + p)] - lp[p];
                        if(any_missing) c += lc[w + p] - lc[p];
                    }
                }
            }
            store_segment(a, y, x, o);
        }
        __syncthreads();   // the next batch overwrites the line buffers
    }
}

// ------------------------------------------------------------------ min / max -------------------------
// neighbourhood.cpp:146-210: extreme of the valid values in the clipped window. fminf / fmaxf return the other
// operand when one is NaN, so missing cells (and the NaN the copy engine writes outside the field for these kernels)
// drop out by themselves and an all-missing window yields NaN; infinite inputs are "invalid" too (util.cpp:16-18)
// and are turned into NaN when their stage lands.
//
// Eight consecutive windows of width w (w >= 8) over v[0 .. w+6] all contain the core v[7 .. w-1]; window i is
// ext(suffix-extreme of v[i..6], core, prefix-extreme of v[w .. w+i-1]).
template <bool IS_MAX>
__device__ __forceinline__ float ext(float x, float y) { return IS_MAX ? fmaxf(x, y) : fminf(x, y); }

template <bool IS_MAX, class F>
__device__ __forceinline__ void eight_windows(F v, int w, float (&out)[RB]) {
    if(w >= RB) {
        float suf[RB], pre[RB];
        suf[RB - 1] = NAN;
        #pragma unroll
        for(int i = RB - 2; i >= 0; i--) suf[i] = ext<IS_MAX>(suf[i + 1], v(i));
        float core = v(RB - 1);
        #pragma unroll
        for(int j = RB; j < w; j++) core = ext<IS_MAX>(core, v(j));
        pre[0] = NAN;
        #pragma unroll
        for(int i = 1; i < RB; i++) pre[i] = ext<IS_MAX>(pre[i - 1], v(w + i - 1));
        #pragma unroll
        for(int i = 0; i < RB; i++) out[i] = ext<IS_MAX>(ext<IS_MAX>(suf[i], core), pre[i]);
    }
    else {
        #pragma unroll
        for(int i = 0; i < RB; i++) {
            float m = NAN;
            #pragma unroll
            for(int j = 0; j < w; j++) m = ext<IS_MAX>(m, v(i + j));
            out[i] = m;
        }
    }
}

template <bool IS_MAX, int HW>
__global__ void __launch_bounds__(NT, MINB_MM) nbh_minmax_tma_kernel(const __grid_constant__ CUtensorMap in_map, const TmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool STATIC = HW > 0;
    const int hw = STATIC ? HW : a.hw, w = 2 * hw + 1;
    const int P = STATIC ? (2 * HW + RB - 1) / RB : a.P;
    const int NS = STATIC ? P + 1 + PREFETCH_MM : a.NS;
    const int NR = NS * RB;
    float* ring = reinterpret_cast<float*>(smem);                       // [NS * RB][NT]
    float* line = ring + (size_t) NR * NT;                              // [RB][LROW_F] column extremes
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(line + RB * LROW_F);

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * a.TX;
    const int y_begin = a.row0 + blockIdx.y * a.rows_per_cta;
    const int y_end = min(y_begin + a.rows_per_cta, a.row0 + a.n_rows_out);
    const int n_batches = (y_end - y_begin + RB - 1) / RB;
    const StageRing R = {ring, bars, &in_map, x0 - a.HL, y_begin + hw - RB * P, NS, P + n_batches};
    R.start();
    const int scol = min(tid + (a.HL - hw), NT - 1);   // staged column of window column tid, see nbh_sum_tma_kernel
    const bool has_col = tid + (a.HL - hw) < NT;       // only the owner of a staged column rewrites its infinities

    const int rel0 = RB * P - 2 * hw;     // ring row of input row y_begin - hw
    for(int k = 0; k < P; k++) {
        R.wait(k);
        float* sp = ring + (size_t) k * RB * NT + scol;
        #pragma unroll
        for(int b = 0; b < RB; b++)
            if(has_col && fabsf(sp[b * NT]) == INFINITY) sp[b * NT] = NAN;
    }
    const int hb = tid >> 5, seg = tid & 31, xo0 = seg * SEG;
    const bool h_active = xo0 < a.TX && x0 + xo0 < a.nx;
    constexpr bool REGWIN = STATIC && HW <= 10;
    float* const my_line = line + idx_f<REGWIN>(tid);
    int first = rel0;                     // ring row of the first row of the window of output row y0
    int s_new = P % NS;
    // register-window form: the 2 hw rows above the entering stage stay in registers from batch to batch (14 moves
    // instead of 14 shared-memory loads per batch and thread; the kernel is bound by the L1/shared data pipe)
    float hist[REGWIN ? 2 * HW : 1];
    if constexpr(REGWIN) {
        #pragma unroll
        for(int j = 0; j < 2 * HW; j++) hist[j] = ring[(rel0 + j) * NT + scol];   // rows y_begin - hw .. y_begin + hw - 1 (sanitised above)
    }

    for(int i = 0; i < n_batches; i++) {
        const int y0 = y_begin + RB * i;
        R.wait(P + i);
        // ---- vertical pass: the windows of the 8 rows of the batch span ring rows first .. first + w + 6, of which
        // the last 8 are the stage that has just landed (its infinities are replaced on the way)
        {
            float* sp = ring + (size_t) s_new * RB * NT + scol;
            float vnew[RB];
            #pragma unroll
            for(int b = 0; b < RB; b++) {
                vnew[b] = sp[b * NT];
                if(fabsf(vnew[b]) == INFINITY) { vnew[b] = NAN; if(has_col) sp[b * NT] = NAN; }
            }
            float out[RB];
            const float* base = ring + scol;
            if constexpr(REGWIN) {
                constexpr int W = 2 * HW + 1;
                float v[W + RB - 1];
                #pragma unroll
                for(int j = 0; j < W - 1; j++) v[j] = hist[j];
                #pragma unroll
                for(int b = 0; b < RB; b++) v[W - 1 + b] = vnew[b];
                eight_windows<IS_MAX>([&](int j) { return v[j]; }, W, out);
                #pragma unroll
                for(int j = 0; j < W - 1; j++) hist[j] = v[j + RB];
                (void) base;
            }
            else
                eight_windows<IS_MAX>([&](int j) { int s = first + j; if(s >= NR) s -= NR; return base[s * NT]; }, w, out);
            #pragma unroll
            for(int b = 0; b < RB; b++) my_line[b * LROW_F] = out[b];
        }
        first += RB;
        if(first >= NR) first -= NR;
        s_new = s_new + 1 == NS ? 0 : s_new + 1;
        __syncthreads();
        R.recycle(i);
        // ---- horizontal pass over window columns xo0 .. xo0 + w + 6
        const int y = y0 + hb;
        if(h_active && y < y_end) {
            const float* l = line + hb * LROW_F;      // window column e at l[idx_f(e)]
            float o[SEG];
            if constexpr(REGWIN) {
                constexpr int W = 2 * HW + 1, NV = (W + SEG - 1 + 3) / 4 * 4;
                float v[NV];
                #pragma unroll
                for(int q = 0; q < NV / 4; q++) {
                    const float4 t = reinterpret_cast<const float4*>(l)[swz_slot(2 * seg + q)];
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                eight_windows<IS_MAX>([&](int j) { return v[j]; }, W, o);
            }
            else
                {
                const float* lp = l + seg * (SEG + 4);    // window column xo0 + j at lp[j + 4 (j / 8)]
                eight_windows<IS_MAX>([&](int j) { return lp[idx_f<false>(j)]; }, w, o);
            }
            store_segment(a, y, x0 + xo0, o);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ mean / sum in float ---------------
// The same filter with every accumulation in fp32 and NO running sums, for the unrolled half-widths up to 7 (the
// metric's configuration). What made the fp64 kernel slow was not arithmetic but pipes and barriers: three
// f32 <-> f64 conversions per pixel on the 16-lane XU pipe, 8-byte line records, two barriers per batch. Here:
//   * vertical pass: the 2 hw rows above the entering stage stay in registers; the column sum of every output row is a
//     FRESH pairwise tree over its 2 hw + 1 rows in a fixed order (top row first). No value is ever subtracted, so a
//     large value leaves no residue behind (the failure of a float running sum), and the sum of a window does not
//     depend on where the CTA's chunk or the 8-row batch starts: row tiles reproduce the whole field bit for bit;
//   * horizontal pass: 8 consecutive pixels per thread from 2 hw + 8 column sums with shared suffix / core / prefix
//     partial sums. The grouping depends on x mod 8 only, which row tiling does not change;
//   * the line buffer is double-buffered: one barrier per batch;
//   * the ring is only read for entering rows, so a stage is handed back to the copy engine as soon as its batch
//     has read it: NS - 1 stages in flight.
// Missing values: the sum of a batch's 8 column sums is non-finite iff one of the 22 rows is; then (uniformly for the
// CTA) the sums are redone over masked values with a validity bit mask per column, and the horizontal pass also
// slides the valid counts. Error: two pairwise trees of depth <= 4 + 5 and one division, <= ~5e-7 relative to the
// mean of |v| over the window (the reference accumulates in double, neighbourhood.cpp:45-99); the bar is 1e-5.
template <int N>
__device__ __forceinline__ float tree_sum_f(const float* v) {
    if constexpr(N == 1) return v[0];
    else return __fadd_rn(tree_sum_f<N / 2>(v), tree_sum_f<N - N / 2>(v + N / 2));
}
// s / c for a float s and a small positive integer c from r = fl(1 / c): Markstein's correction of s * r
__device__ __forceinline__ float quot_f(float s, float fc, float r) {
    const float q0 = __fmul_rn(s, r);
    return __fmaf_rn(__fmaf_rn(-q0, fc, s), r, q0);
}

// two fp32 values in one 64-bit register and the packed IEEE add of sm_100 (FADD2): each half is rounded exactly like a
// scalar add, so packing changes the instruction count, not the result
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float((unsigned) v); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float((unsigned) (v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
template <int N>
__device__ __forceinline__ f32x2 tree_sum_2(const f32x2* v) {
    if constexpr(N == 1) return v[0];
    else return add2(tree_sum_2<N / 2>(v), tree_sum_2<N - N / 2>(v + N / 2));
}

// STAT: 0 = Mean, 1 = Sum
template <int STAT, int HW>
__global__ void __launch_bounds__(NT, MINB_SUM) nbh_sumf_tma_kernel(const __grid_constant__ CUtensorMap in_map, const TmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int W = 2 * HW + 1, P = (2 * HW + RB - 1) / RB, NS = P + 1 + PREFETCH, NR = NS * RB, NH = 2 * HW + RB, NV = W + SEG - 1;
    constexpr int HR = RB / 2, NP = NH - HR;            // output rows b and b + 4 share a packed tree: pairs (h[i], h[i + 4]), i < NP
    float* ring = reinterpret_cast<float*>(smem);                                 // [NS * RB][NT]
    float* fline = ring + (size_t) NR * NT;                                      // [2][RB][NT] column sums (swizzled 16-byte slots)
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(fline + 2 * RB * NT);   // [NS]
    unsigned char* cline = reinterpret_cast<unsigned char*>(bars + NS);           // [2][RB][NT] column valid counts (<= W)

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * a.TX;
    const int y_begin = a.row0 + blockIdx.y * a.rows_per_cta;
    const int y_end = min(y_begin + a.rows_per_cta, a.row0 + a.n_rows_out);
    const int n_batches = (y_end - y_begin + RB - 1) / RB;
    const StageRing R = {ring, bars, &in_map, x0 - a.HL, y_begin + HW - RB * P, NS, P + n_batches};
    R.start();
    const int scol = min(tid + (a.HL - HW), NT - 1);      // see nbh_sum_tma_kernel
    const int x_stage = x0 - a.HL + scol;
    const bool col_ok = x_stage >= 0 && x_stage < a.nx;
    const float* const ring_col = ring + scol;
    const int my_slot = idx_f<true>(tid);
    constexpr int rel0 = RB * P - 2 * HW;                 // ring row of input row y_begin - hw
    // ---- prime the register window with rows y_begin - hw .. y_begin + hw - 1: h[0 .. 2 HW), kept as the pairs
    // hp[i] = (h[i], h[i + 4]); the rows that still wait for their partner are carry[b] = h[2 HW - 4 + b]
    f32x2 hp[NP];
    float carry[HR];
    for(int k = 0; k < P; k++) R.wait(k);
    {
        float h0[2 * HW];
        #pragma unroll
        for(int j = 0; j < 2 * HW; j++) h0[j] = ring_col[(rel0 + j) * NT];
        #pragma unroll
        for(int i = 0; i < NP - RB; i++) hp[i] = pack2(h0[i], h0[i + HR]);
        #pragma unroll
        for(int b = 0; b < HR; b++) carry[b] = h0[2 * HW - HR + b];
    }
    __syncthreads();
    for(int k = 0; k < P; k++) R.recycle(k);
    const int hb = tid >> 5, seg = tid & 31, xo0 = seg * SEG;
    const bool h_active = xo0 < a.TX && x0 + xo0 < a.nx;
    const bool strip_inside = x0 - HW >= 0 && x0 + a.TX - 1 + HW < a.nx;   // no window of the strip is clipped sideways
    const float rc_full = __frcp_rn((float) (W * W));
    int slot[(NV + 3) / 4];                                // this thread's 16-byte slots of a line row
    #pragma unroll
    for(int q = 0; q < (NV + 3) / 4; q++) slot[q] = swz_slot(2 * seg + q);
    auto vertical = [&](int batch, float* fl, float& chk) {
        const int k_new = P + batch;
        R.wait(k_new);
        float nv[RB];
        {
            const float* newp = ring_col + (k_new % NS) * (RB * NT);
            #pragma unroll
            for(int b = 0; b < RB; b++) nv[b] = newp[b * NT];
        }
        #pragma unroll
        for(int b = 0; b < HR; b++) {
            hp[NP - RB + b] = pack2(carry[b], nv[b]);      // (h[2 HW - 4 + b], h[2 HW + b])
            hp[NP - HR + b] = pack2(nv[b], nv[HR + b]);    // (h[2 HW + b], h[2 HW + 4 + b])
            carry[b] = nv[HR + b];
        }
        // a fresh tree per output row, rows b and b + 4 in one packed tree
        f32x2 cs[HR];
        #pragma unroll
        for(int b = 0; b < HR; b++) cs[b] = tree_sum_2<W>(hp + b);
        const f32x2 chk2 = tree_sum_2<HR>(cs);
        #pragma unroll
        for(int b = 0; b < HR; b++) {
            fl[b * NT + my_slot] = lo2(cs[b]);
            fl[(b + HR) * NT + my_slot] = hi2(cs[b]);
        }
        chk = __fadd_rn(lo2(chk2), hi2(chk2));
    };
    // the masked sums and valid counts of the batch whose rows are in hp (before the shift), for a CTA that saw a missing value
    auto vertical_masked = [&](int batch, float* fl, unsigned char* cl) {
        const int y0 = y_begin + RB * batch;
        float mh[NH];
        unsigned mask = 0;
        #pragma unroll
        for(int j = 0; j < NH; j++) {
            const float hj = j < NP ? lo2(hp[j]) : hi2(hp[j - HR]);
            const bool ok = finite_f(hj);
            mh[j] = ok ? hj : 0.f;
            mask |= (ok ? 1u : 0u) << j;
        }
        #pragma unroll
        for(int b = 0; b < RB; b++) {
            const int y = y0 + b;
            const int outside = W - (min(y + HW, a.n_rows_in - 1) - max(y - HW, 0) + 1);   // rows of the window the copy engine zero-filled
            fl[b * NT + my_slot] = tree_sum_f<W>(mh + b);
            cl[b * NT + tid] = (unsigned char) (col_ok ? max(__popc((mask >> b) & ((1u << W) - 1u)) - outside, 0) : 0);
        }
    };
    auto shift = [&]() {
        #pragma unroll
        for(int j = 0; j < NP - RB; j++) hp[j] = hp[j + RB];
    };
    auto horizontal = [&](int batch, const float* fl, const unsigned char* cl, bool any_missing) {
        const int y = y_begin + RB * batch + hb;
        if(!(h_active && y < y_end)) return;
        const int x = x0 + xo0;
        const float4* l4 = reinterpret_cast<const float4*>(fl + hb * NT);     // window column e in slot swz_slot(e / 4)
        float v[(NV + 3) / 4 * 4];
        #pragma unroll
        for(int q = 0; q < (NV + 3) / 4; q++) {
            const float4 t = l4[slot[q]];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        float sw[SEG];
        if constexpr(W >= SEG) {
            // window p = v[p .. p + W - 1] = suffix(v[p .. 6]) + core(v[7 .. W - 1]) + prefix(v[W .. W + p - 1])
            float suf[SEG], pre[SEG];
            suf[SEG - 1] = 0.f;
            #pragma unroll
            for(int p = SEG - 2; p >= 0; p--) suf[p] = p == SEG - 2 ? v[p] : __fadd_rn(v[p], suf[p + 1]);
            const float core = tree_sum_f<W - SEG + 1>(v + SEG - 1);
            pre[0] = 0.f;
            #pragma unroll
            for(int p = 1; p < SEG; p++) pre[p] = p == 1 ? v[W] : __fadd_rn(pre[p - 1], v[W + p - 1]);
            sw[0] = __fadd_rn(suf[0], core);
            sw[SEG - 1] = __fadd_rn(core, pre[SEG - 1]);
            #pragma unroll
            for(int p = 1; p < SEG - 1; p++) sw[p] = __fadd_rn(__fadd_rn(suf[p], core), pre[p]);
        }
        else {
            #pragma unroll
            for(int p = 0; p < SEG; p++) sw[p] = tree_sum_f<W>(v + p);
        }
        float o[SEG];
        const int ch = min(y + HW, a.n_rows_in - 1) - max(y - HW, 0) + 1;   // rows of the window inside the field
        if(!any_missing) {
            if(STAT == 1) {
                #pragma unroll
                for(int p = 0; p < SEG; p++) o[p] = sw[p];
            }
            else if(strip_inside && ch == W) {
                #pragma unroll
                for(int p = 0; p < SEG; p++) o[p] = quot_f(sw[p], (float) (W * W), rc_full);
            }
            else {
                // count = (rows inside the field) x (columns inside the field), neighbourhood.cpp:104-107
                #pragma unroll
                for(int p = 0; p < SEG; p++) {
                    const float fc = (float) (ch * (min(x + p + HW, a.nx - 1) - max(x + p - HW, 0) + 1));
                    o[p] = quot_f(sw[p], fc, __frcp_rn(fc));
                }
            }
        }
        else {
            const unsigned char* lc = cl + hb * NT + xo0;
            int cv[NV];
            #pragma unroll
            for(int j = 0; j < NV; j++) cv[j] = xo0 + j < NT ? lc[j] : 0;
            int c = 0;
            #pragma unroll
            for(int j = 0; j < W; j++) c += cv[j];
            #pragma unroll
            for(int p = 0; p < SEG; p++) {
                const float fc = (float) c;
                if(STAT == 1) o[p] = c > 0 ? sw[p] : NAN;
                else o[p] = c > 0 ? quot_f(sw[p], fc, __frcp_rn(fc)) : NAN;
                if(p + 1 < SEG) c += cv[W + p] - cv[p];
            }
        }
        store_segment(a, y, x, o);
    };

    // One barrier per batch: the line is double-buffered, so the next batch's vertical pass may overwrite the other buffer
    // while slower warps still read this one. (Measured: forming the column sums of batch i + 1 BEFORE the outputs of batch i,
    // so that the two independent streams overlap inside a thread, was slower -- 43.1 against 39.2 us at 4000 x 4000.)
    for(int i = 0; i < n_batches; i++) {
        float* const fl = fline + (i & 1) * (RB * NT);
        unsigned char* const cl = cline + (i & 1) * (RB * NT);
        float chk;
        vertical(i, fl, chk);
        const bool any_missing = __syncthreads_or(!finite_f(chk)) != 0;   // also publishes the line
        R.recycle(P + i);                                                 // every thread has read the entering stage
        if(any_missing) {
            vertical_masked(i, fl, cl);
            __syncthreads();
        }
        shift();
        horizontal(i, fl, cl, any_missing);
    }
}

template <class K>
int run_kernel(K kernel, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& map, const TmaArgs& a) {
    GPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    GPP_LAUNCH(kernel, grid, NT, smem, stream, map, a);
    return GPP_OK;
}

// half-widths with a fully unrolled instantiation; everything else runs the HW = 0 form
#define GPP_FOR_STATIC_HW(X) X(1) X(2) X(3) X(5) X(7) X(10) X(15)

template <int STAT>
int run_sum(int hw, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& map, const TmaArgs& a) {
    switch(hw) {
#define X(H) case H: return run_kernel(nbh_sum_tma_kernel<STAT, H>, grid, smem, stream, map, a);
        GPP_FOR_STATIC_HW(X)
#undef X
        default: return run_kernel(nbh_sum_tma_kernel<STAT, 0>, grid, smem, stream, map, a);
    }
}
// the float kernels: Mean / Sum with an unrolled half-width up to 7
template <int STAT>
int run_sumf(int hw, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& map, const TmaArgs& a) {
    switch(hw) {
        case 1: return run_kernel(nbh_sumf_tma_kernel<STAT, 1>, grid, smem, stream, map, a);
        case 2: return run_kernel(nbh_sumf_tma_kernel<STAT, 2>, grid, smem, stream, map, a);
        case 3: return run_kernel(nbh_sumf_tma_kernel<STAT, 3>, grid, smem, stream, map, a);
        case 5: return run_kernel(nbh_sumf_tma_kernel<STAT, 5>, grid, smem, stream, map, a);
        default: return run_kernel(nbh_sumf_tma_kernel<STAT, 7>, grid, smem, stream, map, a);
    }
}
template <bool IS_MAX>
int run_minmax(int hw, dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap& map, const TmaArgs& a) {
    switch(hw) {
#define X(H) case H: return run_kernel(nbh_minmax_tma_kernel<IS_MAX, H>, grid, smem, stream, map, a);
        GPP_FOR_STATIC_HW(X)
#undef X
        default: return run_kernel(nbh_minmax_tma_kernel<IS_MAX, 0>, grid, smem, stream, map, a);
    }
}

}  // namespace

namespace gpp {

// Runs the statistic through the copy-engine kernels when the shape allows it. *handled = 0 -> the caller falls
// back to the plain kernels (row length not a multiple of 4, narrow fields, very large half-widths).
int nbh_tma_try(const float* d_input, int n_rows_in, int nx, int row0, int n_rows_out, int hw, int statistic, float* d_output,
                cudaStream_t stream, int* handled) {
    *handled = 0;
    if(getenv("GPP_NO_TMA")) return GPP_OK;
    if(nx % 4 != 0 || nx < NT || hw < 1) return GPP_OK;
    if(((uintptr_t) d_input & 15) != 0 || ((uintptr_t) d_output & 15) != 0) return GPP_OK;
    TmaArgs a;
    a.out = d_output;
    a.n_rows_in = n_rows_in; a.nx = nx; a.row0 = row0; a.n_rows_out = n_rows_out; a.hw = hw;
    a.HL = (hw + 3) / 4 * 4;
    a.TX = (NT - a.HL - hw) / SEG * SEG;
    if(a.TX < NT / 2) return GPP_OK;
    a.P = (2 * hw + RB - 1) / RB;
    a.NS = a.P + 1 + (statistic == GPP_MIN || statistic == GPP_MAX ? PREFETCH_MM : PREFETCH);
    const int w = 2 * hw + 1;
    a.n_rcp = std::min(RCP_MAX, w * w + 1);
    const bool minmax = statistic == GPP_MIN || statistic == GPP_MAX;
    static const bool no_float = getenv("GPP_NBH_F64") != nullptr;    // A/B switch: the fp64 kernels for every half-width
    const bool sumf = !no_float && (statistic == GPP_MEAN || statistic == GPP_SUM) && (hw == 1 || hw == 2 || hw == 3 || hw == 5 || hw == 7);
    size_t smem = (size_t) a.NS * STAGE_BYTES + sizeof(unsigned long long) * a.NS;
    if(minmax) smem += sizeof(float) * RB * LROW_F;
    else if(sumf) smem += (sizeof(float) + 1) * 2 * RB * NT;
    else smem += (sizeof(double) + 1) * RB * LROW_D + sizeof(double) * a.n_rcp;
    if(smem > 100 * 1024) return GPP_OK;
    // one wave: strips x chunks <= resident CTAs
    const int strips = (nx + a.TX - 1) / a.TX;
    const int per_sm = std::max(1, std::min(minmax ? MINB_MM : MINB_SUM, (int) ((227 * 1024) / (smem + 1024))));
    const int slots = sm_count() * per_sm;
    int chunks = std::max(1, slots / strips);
    int rows = (n_rows_out + chunks - 1) / chunks;
    // the float kernels take any chunk height (the last batch of a CTA may be partial), so the CTAs can fill the resident
    // slots evenly; the others want whole batches
    rows = sumf ? std::max(4 * RB, rows) : std::max(4 * RB, (rows + RB - 1) / RB * RB);
    a.rows_per_cta = rows;
    chunks = (n_rows_out + rows - 1) / rows;
    CUtensorMap map;
    // sums: zero fill outside the field; extremes: NaN fill (ignored by fminf / fmaxf)
    GPP_TRY(make_field_tensor_map(&map, d_input, n_rows_in, nx, RB, NT, minmax));
    dim3 grid(strips, chunks);
    switch(statistic) {
        case GPP_MEAN: GPP_TRY(sumf ? run_sumf<0>(hw, grid, smem, stream, map, a) : run_sum<0>(hw, grid, smem, stream, map, a)); break;
        case GPP_SUM: GPP_TRY(sumf ? run_sumf<1>(hw, grid, smem, stream, map, a) : run_sum<1>(hw, grid, smem, stream, map, a)); break;
        case GPP_COUNT: GPP_TRY(run_sum<2>(hw, grid, smem, stream, map, a)); break;
        case GPP_MIN: GPP_TRY(run_minmax<false>(hw, grid, smem, stream, map, a)); break;
        case GPP_MAX: GPP_TRY(run_minmax<true>(hw, grid, smem, stream, map, a)); break;
        default: return GPP_OK;
    }
    *handled = 1;
    return GPP_OK;
}

}  // namespace gpp
