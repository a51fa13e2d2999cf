// Row statistics shared by stats.cu (calc_statistic / calc_quantile / brute-force neighbourhoods) and gridding.cu (the values
// of the neighbours of every output point): one warp per "row" of values, read through an accessor.
#pragma once

#include "common.cuh"

#include <algorithm>
#include <atomic>

namespace gpp {
namespace rowstats {

constexpr int STAT_WARPS = 4;
constexpr int STAT_STAGE = 1024;   // values of a row staged in shared memory per warp; longer rows are read from global


struct RowArgs {
    long long n_rows;
    int statistic;            // GPP_* (Quantile: `quantile` / `q_rows`)
    float quantile;
    const float* q_rows;      // per-row quantile levels or NULL
    float* out;
    int* bad_quantile;        // set when a quantile level is outside [0, 1] (util.cpp:113-115 throws)
    unsigned seed;            // RandomChoice
};

// Where the values of a "row" come from.
// LinearRows: row r = a[r * T .. r * T + T).
struct LinearRows {
    const float* a;
    int T;
    __device__ int length(long long) const { return T; }
    __device__ void locate(long long, int*) const {}
    __device__ float get(long long row, const int*, int i) const { return a[(size_t) row * T + i]; }
};
// WindowRows: row r = the (2 hw + 1)^2 window of pixel r = y * nx + x of an ny x nx (x ne) field, clipped to the domain, in the
// order the reference fills its `neighbourhood` vector (rows, then columns, then members; neighbourhood.cpp:560-573,607-622).
struct WindowRows {
    const float* in;
    int ny, nx, ne, hw;
    int row0;                 // output row r of the launch is row row0 + r of the ny rows given
    // w = {first row, first column, columns in the window}
    __device__ void locate(long long row, int* w) const {
        const int r = (int) (row / nx), x = (int) (row - (long long) r * nx), y = row0 + r;
        w[0] = max(0, y - hw);
        w[1] = max(0, x - hw);
        w[2] = min(nx - 1, x + hw) - w[1] + 1;
        w[3] = min(ny - 1, y + hw) - w[0] + 1;
    }
    __device__ int length(long long row) const {
        int w[4];
        locate(row, w);
        return w[2] * w[3] * ne;
    }
    __device__ float get(long long, const int* w, int i) const {
        const int e = i % ne, c = i / ne;
        const int jj = c % w[2], ii = c / w[2];
        return in[((size_t) (w[0] + ii) * nx + (w[1] + jj)) * ne + e];
    }
};

// calc_quantile(array, quantile), util.cpp:111-178, for one row; warp-synchronous, every lane returns the value
template <class V>
__device__ float row_quantile(const V& v, int T, float quantile, int* bad) {
    const int lane = (int) lane_id();
    if(quantile < 0.f || quantile > 1.f) {
        if(lane == 0) atomicExch(bad, 1);
        return NAN;
    }
    if(!is_valid(quantile) || T == 0) return NAN;
    if(quantile == 0.f || quantile == 1.f) {   // util.cpp:121-146: extreme of the valid values
        const bool want_min = quantile == 0.f;
        float best = NAN;
        for(int i = lane; i < T; i += 32) {
            const float x = v(i);
            if(!is_valid(x)) continue;
            if(!is_valid(best) || (want_min ? x < best : x > best)) best = x;
        }
        #pragma unroll
        for(int off = 16; off > 0; off >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, best, off);
            if(is_valid(o) && (!is_valid(best) || (want_min ? o < best : o > best))) best = o;
        }
        return best;
    }
    int n_valid = 0;
    for(int i = lane; i < T; i += 32) n_valid += is_valid(v(i)) ? 1 : 0;
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    if(n_valid == 0) return NAN;
    // util.cpp:160-163: indices and their quantile levels in float arithmetic
    const float span = (float) (n_valid - 1);
    const int lower = (int) floorf(__fmul_rn(quantile, span)), upper = (int) ceilf(__fmul_rn(quantile, span));
    float lower_value = 0.f, upper_value = 0.f;
    for(int i0 = 0; i0 < T; i0 += 32) {
        const int i = i0 + lane;
        const float x = i < T ? v(i) : NAN;
        const bool ok = is_valid(x);
        int pos = 0;
        if(__any_sync(0xffffffffu, ok)) {
            for(int j = 0; j < T; j++) {
                const float y = v(j);
                pos += (is_valid(y) && (y < x || (y == x && j < i))) ? 1 : 0;
            }
        }
        const unsigned ml = __ballot_sync(0xffffffffu, ok && pos == lower), mu = __ballot_sync(0xffffffffu, ok && pos == upper);
        if(ml) lower_value = __shfl_sync(0xffffffffu, x, __ffs(ml) - 1);
        if(mu) upper_value = __shfl_sync(0xffffffffu, x, __ffs(mu) - 1);
    }
    if(lower == upper) return lower_value;
    const float lower_q = __fdiv_rn((float) lower, span), upper_q = __fdiv_rn((float) upper, span);
    const float f = __fdiv_rn(__fsub_rn(quantile, lower_q), __fsub_rn(upper_q, lower_q));
    return __fadd_rn(lower_value, __fmul_rn(__fsub_rn(upper_value, lower_value), f));   // util.cpp:174
}

// calc_statistic / calc_quantile of one row whose values are read through v(i); warp-synchronous
template <class V>
__device__ float row_value(const V& v, int T, const RowArgs& A, long long row) {
    const int lane = (int) lane_id();
    float value = NAN;
    const int st = A.statistic;
    if(st == GPP_MEAN || st == GPP_SUM || st == GPP_COUNT) {   // util.cpp:22-38
        if(st == GPP_COUNT) {
            int count = 0;
            for(int i = lane; i < T; i += 32) count += is_valid(v(i)) ? 1 : 0;
            value = (float) __reduce_add_sync(0xffffffffu, count);
        }
        else {
            if(lane == 0) {
                float total = 0.f;
                int count = 0;
                for(int i = 0; i < T; i++) {
                    const float x = v(i);
                    if(is_valid(x)) { total = __fadd_rn(total, x); count++; }
                }
                if(count > 0) value = st == GPP_MEAN ? __fdiv_rn(total, (float) count) : total;
            }
            value = __shfl_sync(0xffffffffu, value, 0);
        }
    }
    else if(st == GPP_STD || st == GPP_VARIANCE) {             // util.cpp:40-73
        if(lane == 0) {
            float total = 0.f, total2 = 0.f, K = NAN;
            int count = 0;
            for(int i = 0; i < T; i++) {
                const float x = v(i);
                if(!is_valid(x)) continue;
                if(!is_valid(K)) K = x;
                const float d = __fsub_rn(x, K);
                total = __fadd_rn(total, d);
                total2 = __fadd_rn(total2, __fmul_rn(d, d));
                count++;
            }
            if(count > 0) {
                const float mean = __fdiv_rn(total, (float) count), mean2 = __fdiv_rn(total2, (float) count);
                float var = __fsub_rn(mean2, __fmul_rn(mean, mean));
                if(var < 0.f) var = 0.f;
                value = st == GPP_STD ? __fsqrt_rn(var) : var;
            }
        }
        value = __shfl_sync(0xffffffffu, value, 0);
    }
    else if(st == GPP_RANDOMCHOICE) {
        // util.cpp:75-96 picks the (rand() % num_valid)-th valid value; the draw here is a hash of (seed, row): any valid
        // value of the row is a correct outcome, the sequence of the C library's rand() is not reproduced
        int n_valid = 0;
        for(int i = lane; i < T; i += 32) n_valid += is_valid(v(i)) ? 1 : 0;
        n_valid = __reduce_add_sync(0xffffffffu, n_valid);
        if(n_valid > 0) {
            unsigned long long h = ((unsigned long long) A.seed << 32) ^ (unsigned long long) row;
            h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
            const int want = (int) (h % (unsigned long long) n_valid);
            int before = 0;
            for(int i0 = 0; i0 < T; i0 += 32) {
                const int i = i0 + lane;
                const float x = i < T ? v(i) : NAN;
                const unsigned m = __ballot_sync(0xffffffffu, is_valid(x));
                const int here = __popc(m);
                if(want < before + here) {
                    unsigned mm = m;
                    for(int s = 0; s < want - before; s++) mm &= mm - 1;
                    value = __shfl_sync(0xffffffffu, x, __ffs(mm) - 1);
                    break;
                }
                before += here;
            }
        }
    }
    else {                                                     // util.cpp:97-108 and calc_quantile
        float q = A.quantile;
        if(st == GPP_MIN) q = 0.f;
        else if(st == GPP_MEDIAN) q = 0.5f;
        else if(st == GPP_MAX) q = 1.f;
        else if(A.q_rows) q = A.q_rows[row];
        value = row_quantile(v, T, q, A.bad_quantile);
    }
    return value;
}

template <class Rows>
__global__ void __launch_bounds__(STAT_WARPS * 32) row_statistic_kernel(const __grid_constant__ RowArgs A, const __grid_constant__ Rows R) {
    __shared__ float stage[STAT_WARPS][STAT_STAGE];
    const int lane = (int) lane_id(), warp = threadIdx.x >> 5;
    const long long warps_total = (long long) gridDim.x * STAT_WARPS;
    for(long long row = (long long) blockIdx.x * STAT_WARPS + warp; row < A.n_rows; row += warps_total) {
        int where[4] = {0, 0, 0, 0};
        R.locate(row, where);
        const int T = R.length(row);
        float value;
        if(T <= STAT_STAGE) {
            __syncwarp();
            for(int i = lane; i < T; i += 32) stage[warp][i] = R.get(row, where, i);
            __syncwarp();
            const float* st = stage[warp];
            value = row_value([st](int i) { return st[i]; }, T, A, row);
        }
        else
            value = row_value([&](int i) { return R.get(row, where, i); }, T, A, row);
        if(lane == 0) A.out[row] = value;
    }
}

inline std::atomic<unsigned>& choice_calls() { static std::atomic<unsigned> c{0}; return c; }

template <class Rows>
inline int run_rows(const Rows& R, long long n_rows, int statistic, float quantile, const float* d_q_rows, float* d_out, cudaStream_t stream, bool* bad_quantile) {
    if(n_rows <= 0) return GPP_OK;
    DeviceBuffer<int> flag;
    GPP_TRY(flag.alloc(1));
    GPP_CUDA(cudaMemsetAsync(flag.ptr, 0, sizeof(int), stream));
    RowArgs A = {n_rows, statistic, quantile, d_q_rows, d_out, flag.ptr, choice_calls().fetch_add(1) * 2654435761u + 12345u};
    const long long want = (n_rows + STAT_WARPS - 1) / STAT_WARPS;
    const unsigned grid = (unsigned) std::max<long long>(1, std::min<long long>(want, (long long) sm_count() * 16));
    GPP_LAUNCH(row_statistic_kernel<Rows>, grid, STAT_WARPS * 32, 0, stream, A, R);
    if(bad_quantile) {
        int h = 0;
        GPP_TRY(flag.download(&h, 1, stream));
        GPP_CUDA(cudaStreamSynchronize(stream));
        *bad_quantile = h != 0;
    }
    return GPP_OK;
}

// Rows given by segments of an index list: row r = values[index[offset[r] + i]], i < offset[r + 1] - offset[r] (the low 32 bits of
// a 64-bit (row, index) key; the keys are sorted, so a row's values come in ascending index)
struct SegmentRows {
    const float* values;
    const unsigned long long* keys;
    const long long* offset;
    __device__ int length(long long row) const { return (int) (offset[row + 1] - offset[row]); }
    __device__ void locate(long long, int*) const {}
    __device__ float get(long long row, const int*, int i) const { return values[(unsigned) keys[offset[row] + i]]; }
};

}  // namespace rowstats
}  // namespace gpp
