// Neighbourhood Mean / Sum / Count / Min / Max with the input staged by the bulk-tensor copy engine (TMA).
// Replaces gridpp::neighbourhood, src/api/neighbourhood.cpp:28-242, for fields whose row length is a multiple of
// four (the copy engine needs 16-byte row pitches); other shapes take the kernels in neighbourhood.cu.
//
// A CTA of 256 threads owns a strip of 256 staged columns (TX output columns plus `hw` halo columns on each side)
// and walks down a chunk of output rows. One elected thread keeps a ring of NS stages of 8 rows x 256 columns in
// flight: each stage is ONE cp.async.bulk.tensor copy that lands in shared memory and completes an mbarrier; cells
// outside the field arrive as NaN (= gridpp's missing value), which is how the window gets clipped at the domain
// edges (neighbourhood.cpp:104-107) without a special case. Thread t owns staged column t:
//   vertical pass    per row: the entering value is added to the column's window state, the leaving one (still in
//                    the ring, 2 hw + 1 rows older) removed; the state of each of the 8 rows of a batch is written
//                    to a line buffer;
//   horizontal pass  each thread slides along 8 consecutive pixels of one row of the batch and stores 32 bytes.
// HBM traffic is the algorithmic 8 B/pixel plus the halo rows/columns of each CTA (L2 hits for the most part).
#include "tma.cuh"

#include <algorithm>
#include <cstring>

using namespace gpp;

namespace gpp {

int make_field_tensor_map(CUtensorMap* map, const float* base, int rows, int nx, int box_rows, int box_cols) {
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
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if(rc != CUDA_SUCCESS) return fail(GPP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int) rc);
    return GPP_OK;
}

}  // namespace gpp

namespace {

constexpr int NT = 256;              // threads per CTA = staged columns per strip
constexpr int RB = 8;                // rows per stage = output rows per batch
constexpr int SEG = 8;               // consecutive pixels per thread in the horizontal pass
constexpr int LROW = NT + NT / 8;    // line-buffer row length: element e lives at e + e / 8 (bank-conflict padding)
constexpr int PREFETCH = 2;          // stages in flight beyond the ones the window needs
constexpr int RCP_MAX = 1024;
constexpr unsigned STAGE_BYTES = RB * NT * sizeof(float);

struct TmaArgs {
    float* out;
    int n_rows_in, nx, row0, n_rows_out, hw;
    int rows_per_cta;    // output rows per CTA, a multiple of RB
    int TX;              // output columns per strip, a multiple of SEG
    int P;               // stages that must have landed before the first output row: ceil(2 hw / RB)
    int NS;              // stages in the ring: P + 1 + PREFETCH
    int n_rcp;           // entries of the reciprocal table (0: divide)
    int HL;              // staged columns to the left of the strip: hw rounded up to a multiple of 4 (the copy engine
                         // needs 16-byte aligned box origins); the first HL - hw of them are not part of any window
};

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e38f; }   // == is_valid(v), util.cpp:16-18

// The ring of stages. Stage k holds input rows r0 + RB k .. r0 + RB k + RB - 1 of staged columns xs0 .. xs0 + NT - 1
// in ring slot k % NS; its barrier completes when the copy has landed. r0 = y_begin + hw - RB P, so that the rows
// entering the window during batch i are exactly stage P + i.
struct StageRing {
    float* ring;
    unsigned long long* bars;
    const CUtensorMap* map;
    int xs0, r0, NS, total;

    __device__ __forceinline__ void issue(int k) const {   // one thread
        const int slot = k % NS;
        mbar_expect_tx(&bars[slot], STAGE_BYTES);
        tma_load_2d(ring + (size_t) slot * RB * NT, map, xs0, r0 + RB * k, &bars[slot]);
    }
    __device__ __forceinline__ void start() const {
        if(threadIdx.x == 0) {
            tma_prefetch_descriptor(map);
            for(int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
            mbar_fence_init();
            for(int k = 0; k < min(NS, total); k++) issue(k);
        }
        __syncthreads();
    }
    __device__ __forceinline__ void wait(int k) const { mbar_wait(&bars[k % NS], (unsigned) (k / NS) & 1u); }
    // after every thread is done with stage k (a __syncthreads separates its last read from this call)
    __device__ __forceinline__ void recycle(int k) const {
        if(threadIdx.x == 0 && k + NS < total) issue(k + NS);
    }
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
// than 2 hw + 1 values, then a sliding row sum over at most 2 hw + 8 columns) and divided by the valid count.
// STAT: 0 = Mean, 1 = Sum, 2 = Count
template <int STAT>
__global__ void __launch_bounds__(NT, 3) nbh_sum_tma_kernel(const __grid_constant__ CUtensorMap in_map, const TmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int NR = a.NS * RB;
    float* ring = reinterpret_cast<float*>(smem);                                 // [NS * RB][NT]
    double* line = reinterpret_cast<double*>(ring + (size_t) NR * NT);           // [RB][LROW] column sums
    int* cline = reinterpret_cast<int*>(line + RB * LROW);                       // [RB][LROW] column valid counts
    double* rcp = reinterpret_cast<double*>(cline + RB * LROW);                  // [n_rcp]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(rcp + a.n_rcp);

    const int tid = threadIdx.x;
    const int hw = a.hw, w = 2 * hw + 1;
    const int x0 = blockIdx.x * a.TX;
    const int y_begin = a.row0 + blockIdx.y * a.rows_per_cta;
    const int y_end = min(y_begin + a.rows_per_cta, a.row0 + a.n_rows_out);
    const int n_batches = (y_end - y_begin + RB - 1) / RB;
    const StageRing R = {ring, bars, &in_map, x0 - a.HL, y_begin + hw - RB * a.P, a.NS, a.P + n_batches};
    R.start();
    for(int c = tid; c < a.n_rcp; c += NT) rcp[c] = c > 0 ? 1.0 / (double) c : 0.0;

    const int x_stage = x0 - a.HL + tid;
    const bool col_ok = x_stage >= 0 && x_stage < a.nx;
    // window column e = tid - (HL - hw) of the strip; the HL - hw leftmost staged columns park their records in the
    // unused tail of the line
    const int ecol = tid >= a.HL - hw ? tid - (a.HL - hw) : NT - (a.HL - hw) + tid;
    const bool in_window = tid >= a.HL - hw;
    // ---- prime the column window with rows y_begin - hw .. y_begin + hw - 1
    double csum = 0.0;
    int ccnt = 0;
    int rows_in = 0;                 // rows of the current window that lie inside the field (uniform)
    const int rel0 = RB * a.P - 2 * hw;   // ring row of input row y_begin - hw
    for(int k = 0; k < a.P; k++) R.wait(k);
    for(int rel = rel0; rel < RB * a.P; rel++) {
        const float v = ring[rel * NT + tid];
        if(finite_f(v)) { csum += (double) v; ccnt++; }
        const int r = R.r0 + rel;
        rows_in += (r >= 0 && r < a.n_rows_in) ? 1 : 0;
    }
    int old_slot = rel0;             // ring row of the row that leaves next; rel0 < RB <= NR
    double* const my_line = line + ecol + (ecol >> 3);
    int* const my_cline = cline + ecol + (ecol >> 3);
    const int hb = tid >> 5, seg = tid & 31, xo0 = seg * SEG;
    const bool h_active = xo0 < a.TX && x0 + xo0 < a.nx;
    // staged column (relative to xo0) entering the row window when it slides from pixel p to p + 1
    int in_off[SEG - 1];
    #pragma unroll
    for(int p = 0; p < SEG - 1; p++) in_off[p] = (w + p) + ((w + p) >> 3);

    for(int i = 0; i < n_batches; i++) {
        const int y0 = y_begin + RB * i;
        R.wait(a.P + i);
        const float* newp = ring + (size_t) ((a.P + i) % a.NS) * RB * NT + tid;
        // ---- vertical pass: row y0 + b + hw enters, the record of output row y0 + b is written, row y0 + b - hw leaves
        bool missing = false;
        #pragma unroll
        for(int b = 0; b < RB; b++) {
            const float vn = newp[b * NT];
            if(finite_f(vn)) { csum += (double) vn; ccnt++; }
            const int rn = y0 + b + hw;
            rows_in += (rn >= 0 && rn < a.n_rows_in) ? 1 : 0;
            my_line[b * LROW] = csum;
            my_cline[b * LROW] = ccnt;
            missing = missing || (in_window && ccnt != (col_ok ? rows_in : 0));
            const float vo = ring[old_slot * NT + tid];
            if(finite_f(vo)) { csum -= (double) vo; ccnt--; }
            const int ro = y0 + b - hw;
            rows_in -= (ro >= 0 && ro < a.n_rows_in) ? 1 : 0;
            old_slot = old_slot + 1 == NR ? 0 : old_slot + 1;
        }
        // line buffer complete; stage i is dead. "missing": some window of this batch holds a missing value inside
        // the field, so the counts are not the clipped window areas.
        const bool any_missing = __syncthreads_or(missing) != 0;
        R.recycle(i);
        // ---- horizontal pass
        const int y = y0 + hb;
        if(h_active && y < y_end) {
            const double* l = line + hb * LROW + seg * (SEG + 1);    // staged column xo0 + j at l[j + j / 8]
            const int* lc = cline + hb * LROW + seg * (SEG + 1);
            double s = 0.0;
            {
                int j = 0;
                const double* lj = l;
                for(; j + 8 <= w; j += 8, lj += 9) {
                    #pragma unroll
                    for(int u = 0; u < 8; u++) s += lj[u];
                }
                for(int u = 0; j + u < w; u++) s += lj[u];
            }
            float o[SEG];
            const int x = x0 + xo0;
            if(!any_missing) {
                // count = (rows of the window inside the field) x (columns inside the field), neighbourhood.cpp:104-107
                const int ch = min(y + hw, a.n_rows_in - 1) - max(y - hw, 0) + 1;
                #pragma unroll
                for(int p = 0; p < SEG; p++) {
                    const int cw = min(x + p + hw, a.nx - 1) - max(x + p - hw, 0) + 1;
                    const int c = ch * cw;
                    if(STAT == 2) o[p] = (float) c;
                    else if(STAT == 1) o[p] = (float) s;
                    else o[p] = c < a.n_rcp ? (float) (s * rcp[c]) : (float) (s / (double) c);   // neighbourhood.cpp:133-142
                    if(p + 1 < SEG) s += l[in_off[p]] - l[p];
                }
            }
            else {
                int c = 0;
                {
                    int j = 0;
                    const int* lj = lc;
                    for(; j + 8 <= w; j += 8, lj += 9) {
                        #pragma unroll
                        for(int u = 0; u < 8; u++) c += lj[u];
                    }
                    for(int u = 0; j + u < w; u++) c += lj[u];
                }
                #pragma unroll
                for(int p = 0; p < SEG; p++) {
                    if(STAT == 2) o[p] = (float) c;
                    else if(STAT == 1) o[p] = c > 0 ? (float) s : NAN;
                    else o[p] = c > 0 ? (c < a.n_rcp ? (float) (s * rcp[c]) : (float) (s / (double) c)) : NAN;
                    if(p + 1 < SEG) {
                        s += l[in_off[p]] - l[p];
                        c += lc[in_off[p]] - lc[p];
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
// operand when one is NaN, so missing cells (and the NaN the copy engine writes outside the field) drop out by
// themselves and an all-missing window yields NaN; infinite inputs are "invalid" too (util.cpp:16-18) and are
// turned into NaN when their stage lands.
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
            for(int j = 0; j < w; j++) m = ext<IS_MAX>(m, v(i + j));
            out[i] = m;
        }
    }
}

template <bool IS_MAX>
__global__ void __launch_bounds__(NT, 3) nbh_minmax_tma_kernel(const __grid_constant__ CUtensorMap in_map, const TmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int NR = a.NS * RB;
    float* ring = reinterpret_cast<float*>(smem);                       // [NS * RB][NT]
    float* line = ring + (size_t) NR * NT;                              // [RB][LROW] column extremes
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(line + RB * LROW);

    const int tid = threadIdx.x;
    const int hw = a.hw, w = 2 * hw + 1;
    const int x0 = blockIdx.x * a.TX;
    const int y_begin = a.row0 + blockIdx.y * a.rows_per_cta;
    const int y_end = min(y_begin + a.rows_per_cta, a.row0 + a.n_rows_out);
    const int n_batches = (y_end - y_begin + RB - 1) / RB;
    const StageRing R = {ring, bars, &in_map, x0 - a.HL, y_begin + hw - RB * a.P, a.NS, a.P + n_batches};
    R.start();
    const int ecol = tid >= a.HL - hw ? tid - (a.HL - hw) : NT - (a.HL - hw) + tid;   // see nbh_sum_tma_kernel

    const int rel0 = RB * a.P - 2 * hw;   // ring row of input row y_begin - hw
    for(int k = 0; k < a.P; k++) {
        R.wait(k);
        float* sp = ring + (size_t) k * RB * NT + tid;
        #pragma unroll
        for(int b = 0; b < RB; b++)
            if(fabsf(sp[b * NT]) == INFINITY) sp[b * NT] = NAN;
    }
    const int hb = tid >> 5, seg = tid & 31, xo0 = seg * SEG;
    const bool h_active = xo0 < a.TX && x0 + xo0 < a.nx;
    float* const my_line = line + ecol + (ecol >> 3);
    int first = rel0;                     // ring row of the first row of the window of output row y0

    for(int i = 0; i < n_batches; i++) {
        const int y0 = y_begin + RB * i;
        R.wait(a.P + i);
        {
            float* sp = ring + (size_t) ((a.P + i) % a.NS) * RB * NT + tid;
            #pragma unroll
            for(int b = 0; b < RB; b++)
                if(fabsf(sp[b * NT]) == INFINITY) sp[b * NT] = NAN;
        }
        // ---- vertical pass: the windows of the 8 rows of the batch over ring rows first .. first + w + 6
        {
            float out[RB];
            const float* base = ring + tid;
            eight_windows<IS_MAX>([&](int j) { int s = first + j; if(s >= NR) s -= NR; return base[s * NT]; }, w, out);
            #pragma unroll
            for(int b = 0; b < RB; b++) my_line[b * LROW] = out[b];
        }
        first += RB;
        if(first >= NR) first -= NR;
        __syncthreads();
        R.recycle(i);
        // ---- horizontal pass over staged columns xo0 .. xo0 + w + 6
        const int y = y0 + hb;
        if(h_active && y < y_end) {
            const float* l = line + hb * LROW + seg * (SEG + 1);
            float o[SEG];
            eight_windows<IS_MAX>([&](int j) { return l[j + (j >> 3)]; }, w, o);
            store_segment(a, y, x0 + xo0, o);
        }
        __syncthreads();
    }
}

template <class K>
int prepare_kernel(K kernel, size_t smem) {
    GPP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    return GPP_OK;
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
    a.NS = a.P + 1 + PREFETCH;
    const int w = 2 * hw + 1;
    a.n_rcp = std::min(RCP_MAX, w * w + 1);
    const bool minmax = statistic == GPP_MIN || statistic == GPP_MAX;
    size_t smem = (size_t) a.NS * STAGE_BYTES + sizeof(unsigned long long) * a.NS;
    if(minmax) smem += sizeof(float) * RB * LROW;
    else smem += (sizeof(double) + sizeof(int)) * RB * LROW + sizeof(double) * a.n_rcp;
    if(smem > 100 * 1024) return GPP_OK;
    // one wave: strips x chunks <= resident CTAs
    const int strips = (nx + a.TX - 1) / a.TX;
    const int per_sm = std::max(1, std::min(3, (int) ((227 * 1024) / (smem + 1024))));
    const int slots = sm_count() * per_sm;
    int chunks = std::max(1, slots / strips);
    int rows = (n_rows_out + chunks - 1) / chunks;
    rows = std::max(4 * RB, (rows + RB - 1) / RB * RB);
    a.rows_per_cta = rows;
    chunks = (n_rows_out + rows - 1) / rows;
    CUtensorMap map;
    GPP_TRY(make_field_tensor_map(&map, d_input, n_rows_in, nx, RB, NT));
    dim3 grid(strips, chunks);
    switch(statistic) {
        case GPP_MEAN: GPP_TRY(prepare_kernel(nbh_sum_tma_kernel<0>, smem)); GPP_LAUNCH(nbh_sum_tma_kernel<0>, grid, NT, smem, stream, map, a); break;
        case GPP_SUM: GPP_TRY(prepare_kernel(nbh_sum_tma_kernel<1>, smem)); GPP_LAUNCH(nbh_sum_tma_kernel<1>, grid, NT, smem, stream, map, a); break;
        case GPP_COUNT: GPP_TRY(prepare_kernel(nbh_sum_tma_kernel<2>, smem)); GPP_LAUNCH(nbh_sum_tma_kernel<2>, grid, NT, smem, stream, map, a); break;
        case GPP_MIN: GPP_TRY(prepare_kernel(nbh_minmax_tma_kernel<false>, smem)); GPP_LAUNCH(nbh_minmax_tma_kernel<false>, grid, NT, smem, stream, map, a); break;
        case GPP_MAX: GPP_TRY(prepare_kernel(nbh_minmax_tma_kernel<true>, smem)); GPP_LAUNCH(nbh_minmax_tma_kernel<true>, grid, NT, smem, stream, map, a); break;
        default: return GPP_OK;
    }
    *handled = 1;
    return GPP_OK;
}

}  // namespace gpp
