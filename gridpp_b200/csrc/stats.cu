// Row-wise statistics on the device: gridpp::calc_statistic (src/api/util.cpp:19-110,208-215), gridpp::calc_quantile
// (util.cpp:111-207) and gridpp::interpolate (util.cpp:339-431) for many rows / values at once.
//
// One warp per row of T values (an ensemble at one grid point, the members of one location, ...). The reference's float
// accumulations are sequential in element order, so the sums are formed by one lane from the row staged in shared
// memory; counts, extremes and the order statistics (rank by counting: sorted position of element i = #(v_j < v_i) +
// #(v_j == v_i, j < i)) use all lanes. No sort, no scratch allocation: any T.
#include "rowstats.cuh"

using namespace gpp;

using namespace gpp::rowstats;

namespace {

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
