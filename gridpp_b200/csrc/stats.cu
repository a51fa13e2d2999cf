// Row-wise statistics on the device: gridpp::calc_statistic (src/api/util.cpp:19-110,208-215), gridpp::calc_quantile
// (util.cpp:111-207) and gridpp::interpolate (util.cpp:339-431) for many rows / values at once.
//
// One warp per row of T values (an ensemble at one grid point, the members of one location, ...). The reference's float
// accumulations are sequential in element order, so the sums are formed by one lane from the row staged in shared
// memory; counts, extremes and the order statistics (rank by counting: sorted position of element i = #(v_j < v_i) +
// #(v_j == v_i, j < i)) use all lanes. No sort, no scratch allocation: any T.
#include "common.cuh"

#include <algorithm>
#include <atomic>

using namespace gpp;

namespace {

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

// gridpp::interpolate(x, iX, iY), util.cpp:377-414 with get_lower_index / get_upper_index (:339-376); one thread per x
__global__ void interpolate_kernel(const float* __restrict__ x, long long n, const float* __restrict__ iX, const float* __restrict__ iY, int m,
                                   float* __restrict__ out) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= n) return;
    const float xv = x[t];
    float y = NAN;
    if(is_valid(xv) && m > 0) {
        if(xv > iX[m - 1]) y = iY[m - 1];
        else if(xv < iX[0]) y = iY[0];
        else {
            int i0 = -1, i1 = -1;
            for(int i = 0; i < m; i++) {
                const float c = iX[i];
                if(!is_valid(c)) continue;
                if(c < xv) i0 = i;
                else if(c == xv) { i0 = i; break; }
                else break;
            }
            for(int i = m - 1; i >= 0; i--) {
                const float c = iX[i];
                if(!is_valid(c)) continue;
                if(c > xv) i1 = i;
                else if(c == xv) { i1 = i; break; }
                else break;
            }
            if(i0 >= 0 && i1 >= 0) {
                const float x0 = iX[i0], x1 = iX[i1], y0 = iY[i0], y1 = iY[i1];
                if(x0 == x1) {
                    if(i0 == 0 && i1 == m - 1) y = __fdiv_rn(__fadd_rn(y0, y1), 2.f);
                    else if(i0 == 0) y = y1;
                    else if(i1 == m - 1) y = y0;
                    else y = __fdiv_rn(__fadd_rn(y0, y1), 2.f);
                }
                else y = __fadd_rn(y0, __fdiv_rn(__fmul_rn(__fsub_rn(y1, y0), __fsub_rn(xv, x0)), __fsub_rn(x1, x0)));
            }
        }
    }
    out[t] = y;
}

std::atomic<unsigned> g_choice_calls{0};

template <class Rows>
int run_rows(const Rows& R, long long n_rows, int statistic, float quantile, const float* d_q_rows, float* d_out, cudaStream_t stream, bool* bad_quantile) {
    if(n_rows <= 0) return GPP_OK;
    DeviceBuffer<int> flag;
    GPP_TRY(flag.alloc(1));
    GPP_CUDA(cudaMemsetAsync(flag.ptr, 0, sizeof(int), stream));
    RowArgs A = {n_rows, statistic, quantile, d_q_rows, d_out, flag.ptr, g_choice_calls.fetch_add(1) * 2654435761u + 12345u};
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
int run_rows(const float* d_a, long long n_rows, int T, int statistic, float quantile, const float* d_q_rows, float* d_out, cudaStream_t stream,
             bool* bad_quantile) {
    return run_rows(LinearRows{d_a, T}, n_rows, statistic, quantile, d_q_rows, d_out, stream, bad_quantile);
}

bool known_statistic(int s) {
    return s == GPP_MEAN || s == GPP_MIN || s == GPP_MEDIAN || s == GPP_MAX || s == GPP_STD || s == GPP_VARIANCE || s == GPP_SUM || s == GPP_COUNT ||
           s == GPP_RANDOMCHOICE;
}

}  // namespace

extern "C" {

int gpp_calc_statistic_device(const float* d_array, long long n_rows, int row_length, int statistic, float* d_out, void* stream) {
    if(n_rows < 0 || row_length < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative size");
    if(!known_statistic(statistic)) return fail(GPP_ERR_RUNTIME, "Internal error. Cannot compute statistic");   // util.cpp:105
    GPP_TRY(ensure_device());
    return run_rows(d_array, n_rows, row_length, statistic, NAN, nullptr, d_out, (cudaStream_t) stream, nullptr);
}

int gpp_calc_statistic_host(const float* array, long long n_rows, int row_length, int statistic, float* out) {
    if(n_rows < 0 || row_length < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative size");
    if(!known_statistic(statistic)) return fail(GPP_ERR_RUNTIME, "Internal error. Cannot compute statistic");
    GPP_TRY(ensure_device());
    if(n_rows == 0) return GPP_OK;
    DeviceBuffer<float> d_a, d_out;
    GPP_TRY(d_a.upload(array, (size_t) n_rows * row_length));
    GPP_TRY(d_out.alloc((size_t) n_rows));
    GPP_TRY(run_rows(d_a.ptr, n_rows, row_length, statistic, NAN, nullptr, d_out.ptr, 0, nullptr));
    GPP_TRY(d_out.download(out, (size_t) n_rows));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_calc_quantile_host(const float* array, long long n_rows, int row_length, float quantile, const float* quantile_rows, float* out) {
    if(n_rows < 0 || row_length < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative size");
    if(!quantile_rows && (quantile < 0.f || quantile > 1.f))
        return fail(GPP_ERR_INVALID_ARGUMENT, "calc_quantile: Quantile must be between 0 and 1 inclusive");   // util.cpp:113-115
    GPP_TRY(ensure_device());
    if(n_rows == 0) return GPP_OK;
    DeviceBuffer<float> d_a, d_q, d_out;
    GPP_TRY(d_a.upload(array, (size_t) n_rows * row_length));
    if(quantile_rows) GPP_TRY(d_q.upload(quantile_rows, (size_t) n_rows));
    GPP_TRY(d_out.alloc((size_t) n_rows));
    bool bad = false;
    GPP_TRY(run_rows(d_a.ptr, n_rows, row_length, GPP_QUANTILE, quantile, quantile_rows ? d_q.ptr : nullptr, d_out.ptr, 0, &bad));
    if(bad) return fail(GPP_ERR_INVALID_ARGUMENT, "calc_quantile: Quantile must be between 0 and 1 inclusive");
    GPP_TRY(d_out.download(out, (size_t) n_rows));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

/* neighbourhood_brute_force / neighbourhood_quantile, neighbourhood.cpp:528-535 and the helpers :547-630 */
int gpp_neighbourhood_brute_force_device(const float* d_input, int n_rows_in, int nx, int ne, int row0, int n_rows_out, int halfwidth,
                                         int statistic, float quantile, float* d_output, void* stream) {
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");   // neighbourhood.cpp:548-549
    if(n_rows_in < 0 || nx < 0 || ne < 0 || row0 < 0 || n_rows_out < 0 || row0 + n_rows_out > n_rows_in)
        return fail(GPP_ERR_INVALID_ARGUMENT, "row window out of bounds");
    if(statistic == GPP_QUANTILE) {
        if(quantile < 0.f || quantile > 1.f) return fail(GPP_ERR_INVALID_ARGUMENT, "calc_quantile: Quantile must be between 0 and 1 inclusive");
    }
    else if(!known_statistic(statistic)) return fail(GPP_ERR_RUNTIME, "Internal error. Cannot compute statistic");
    GPP_TRY(ensure_device());
    if(n_rows_out == 0 || nx == 0 || ne == 0) return GPP_OK;
    const WindowRows R = {d_input, n_rows_in, nx, ne, halfwidth, row0};
    return run_rows(R, (long long) n_rows_out * nx, statistic, quantile, nullptr, d_output, (cudaStream_t) stream, nullptr);
}

int gpp_neighbourhood_brute_force_host(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float quantile, float* output) {
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");
    if(statistic == GPP_QUANTILE && (quantile < 0.f || quantile > 1.f))
        return fail(GPP_ERR_INVALID_ARGUMENT, "calc_quantile: Quantile must be between 0 and 1 inclusive");
    GPP_TRY(ensure_device());
    if(ny <= 0 || nx <= 0 || ne <= 0) return GPP_OK;   // neighbourhood.cpp:550-551,591-592: empty in, empty out
    const size_t n = (size_t) ny * nx;
    DeviceBuffer<float> d_in, d_out;
    GPP_TRY(d_in.upload(input, n * ne));
    GPP_TRY(d_out.alloc(n));
    GPP_TRY(gpp_neighbourhood_brute_force_device(d_in.ptr, ny, nx, ne, 0, ny, halfwidth, statistic, quantile, d_out.ptr, nullptr));
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_interpolate_host(const float* x, long long n, const float* iX, const float* iY, int m, float* out) {
    if(n < 0 || m < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative size");
    GPP_TRY(ensure_device());
    if(n == 0) return GPP_OK;
    DeviceBuffer<float> d_x, d_ix, d_iy, d_out;
    GPP_TRY(d_x.upload(x, (size_t) n));
    GPP_TRY(d_ix.upload(iX, (size_t) m));
    GPP_TRY(d_iy.upload(iY, (size_t) m));
    GPP_TRY(d_out.alloc((size_t) n));
    GPP_LAUNCH(interpolate_kernel, (unsigned) ((n + 255) / 256), 256, 0, 0, d_x.ptr, n, d_ix.ptr, d_iy.ptr, m, d_out.ptr);
    GPP_TRY(d_out.download(out, (size_t) n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

}  // extern "C"
