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

// gridpp::neighbourhood_search, neighbourhood_search.cpp:7-113: one thread per pixel walks its clipped window in the reference's
// (row-major) order -- the float accumulation, the "first hit wins" nearest-target rule and the `counter > 0` short-cut all
// depend on that order.
__global__ void neighbourhood_search_kernel(const float* __restrict__ array, const float* __restrict__ search, const int* __restrict__ apply,
                                            int ny, int nx, int hw, float tmin, float tmax, float delta, float* __restrict__ out) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= (long long) ny * nx) return;
    const int y = (int) (t / nx), x = (int) (t - (long long) y * nx);
    const float here = array[t], s_here = search[t];
    float result = here;
    const bool skip = !is_valid(s_here) || (apply && apply[t] == 0);              // :40-50
    if(!skip && (!apply || apply[t] == 1)) {                                     // :63
        float nearest = NAN, nearest_value = 0.f, accum = 0.f;
        int counter = 0;
        for(int yy = max(0, y - hw); yy <= min(ny - 1, y + hw); yy++)
            for(int xx = max(0, x - hw); xx <= min(nx - 1, x + hw); xx++) {
                const float sv = search[(size_t) yy * nx + xx], av = array[(size_t) yy * nx + xx];
                if(!is_valid(sv) || !is_valid(av)) continue;
                if(sv >= tmin && sv <= tmax) {
                    counter++;
                    accum = __fadd_rn(accum, av);
                }
                else if(counter > 0) continue;
                else if(fabsf(__fsub_rn(sv, s_here)) >= delta) {
                    if(!is_valid(nearest)) { nearest = sv; nearest_value = av; }
                    else {
                        const float curr = fminf(fabsf(__fsub_rn(sv, tmin)), fabsf(__fsub_rn(sv, tmax)));
                        const float best = fminf(fabsf(__fsub_rn(nearest, tmin)), fabsf(__fsub_rn(nearest, tmax)));
                        if(curr < best) { nearest = sv; nearest_value = av; }
                    }
                }
            }
        if(counter > 0) result = __fdiv_rn(accum, (float) counter);
        else if(is_valid(nearest)) result = nearest_value;
    }
    out[t] = result;
}

// gridpp::calc_gradient, MinMax branch (calc_gradient.cpp:27-75): the values at the window's lowest and highest base
__global__ void gradient_minmax_kernel(const float* __restrict__ base, const float* __restrict__ values, int ny, int nx, int hw, int num_min,
                                       float min_range, float default_gradient, float* __restrict__ out) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= (long long) ny * nx) return;
    const int y = (int) (t / nx), x = (int) (t - (long long) y * nx);
    float cmax = NAN, cmin = NAN, vmax = 0.f, vmin = 0.f;
    int count = 0;
    for(int yy = max(0, y - hw); yy <= min(ny - 1, y + hw); yy++)
        for(int xx = max(0, x - hw); xx <= min(nx - 1, x + hw); xx++) {
            const float b = base[(size_t) yy * nx + xx], v = values[(size_t) yy * nx + xx];
            if(!is_valid(b) || !is_valid(v)) continue;
            if(!is_valid(cmax) || b > cmax) { cmax = b; vmax = v; }
            if(!is_valid(cmin) || b < cmin) { cmin = b; vmin = v; }
            count++;
        }
    float r = default_gradient;
    if(count >= num_min && is_valid(cmax) && is_valid(cmin) && !(fabsf(__fsub_rn(cmax, cmin)) <= min_range))
        r = __fdiv_rn(__fsub_rn(vmax, vmin), __fsub_rn(cmax, cmin));
    out[t] = r;
}

// LinearRegression branch (calc_gradient.cpp:77-124): the moment fields where both inputs are valid ...
__global__ void gradient_moments_kernel(const float* __restrict__ base, const float* __restrict__ values, long long n, float* __restrict__ base0,
                                        float* __restrict__ values0, float* __restrict__ bb, float* __restrict__ bv, float* __restrict__ valid) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= n) return;
    const float b = base[t], v = values[t];
    const bool ok = is_valid(b) && is_valid(v);
    base0[t] = ok ? b : NAN;
    values0[t] = ok ? v : NAN;
    bb[t] = ok ? __fmul_rn(b, b) : NAN;      // pow(base, 2) in double, stored as float: one rounding of the exact square
    bv[t] = ok ? __fmul_rn(b, v) : NAN;
    valid[t] = ok ? 1.f : 0.f;
}
// ... and the regression slope from their neighbourhood means
__global__ void gradient_regression_kernel(const float* __restrict__ mX, const float* __restrict__ mY, const float* __restrict__ mXX,
                                           const float* __restrict__ mXY, const float* __restrict__ count, long long n, int num_min, float min_range,
                                           float default_gradient, float* __restrict__ out) {
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= n) return;
    float r = default_gradient;
    const float var = __fsub_rn(mXX[t], __fmul_rn(mX[t], mX[t]));
    if(count[t] >= (float) num_min && is_valid(mXX[t]) && is_valid(mXY[t]) && is_valid(mX[t]) && var != 0.f) {
        bool valid_range = true;
        if(is_valid(min_range)) {
            const float range = __fsqrt_rn(var);
            valid_range = is_valid(range) && !(range < min_range);
        }
        if(valid_range) r = __fdiv_rn(__fsub_rn(mXY[t], __fmul_rn(mX[t], mY[t])), var);
    }
    out[t] = r;
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

/* gridpp::neighbourhood_search, neighbourhood_search.cpp:7-113 */
int gpp_neighbourhood_search_host(const float* array, const float* search_array, int ny, int nx, int halfwidth, float search_target_min,
                                  float search_target_max, float search_delta, const int* apply_array, float* output) {
    if(search_target_min > search_target_max)
        return fail(GPP_ERR_INVALID_ARGUMENT, "Search_target_min must be smaller than search_target_max");   // :10-12
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "halfwidth must be positive");                   // :13-15
    if(ny < 0 || nx < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative size");
    GPP_TRY(ensure_device());
    const size_t n = (size_t) ny * nx;
    if(n == 0) return GPP_OK;
    DeviceBuffer<float> d_a, d_s, d_out;
    DeviceBuffer<int> d_apply;
    GPP_TRY(d_a.upload(array, n));
    GPP_TRY(d_s.upload(search_array, n));
    if(apply_array) GPP_TRY(d_apply.upload(apply_array, n));
    GPP_TRY(d_out.alloc(n));
    GPP_LAUNCH(neighbourhood_search_kernel, (unsigned) ((n + 127) / 128), 128, 0, 0, d_a.ptr, d_s.ptr, apply_array ? d_apply.ptr : nullptr, ny, nx,
               halfwidth, search_target_min, search_target_max, search_delta, d_out.ptr);
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

/* gridpp::calc_gradient, calc_gradient.cpp:6-126. gradient_type: GPP_GRADIENT_MINMAX (0) or GPP_GRADIENT_LINEAR_REGRESSION (10) */
int gpp_calc_gradient_host(const float* base, const float* values, int ny, int nx, int gradient_type, int halfwidth, int num_min, float min_range,
                           float default_gradient, float* output) {
    if(halfwidth <= 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Halwidth cannot be <= 0; must be positive integer");   // :9-10
    if(is_valid(min_range) && min_range < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "min_range must be >= 0");
    if(num_min < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "num_min must be >= 0");
    if(ny <= 0 || nx < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "base input has no size");
    GPP_TRY(ensure_device());
    const size_t n = (size_t) ny * nx;
    if(n == 0) return GPP_OK;
    const unsigned blocks = (unsigned) ((n + 127) / 128);
    DeviceBuffer<float> d_b, d_v, d_out;
    GPP_TRY(d_b.upload(base, n));
    GPP_TRY(d_v.upload(values, n));
    GPP_TRY(d_out.alloc(n));
    if(gradient_type == GPP_GRADIENT_MINMAX) {
        GPP_LAUNCH(gradient_minmax_kernel, blocks, 128, 0, 0, d_b.ptr, d_v.ptr, ny, nx, halfwidth, num_min, min_range, default_gradient, d_out.ptr);
    }
    else if(gradient_type == GPP_GRADIENT_LINEAR_REGRESSION) {
        DeviceBuffer<float> in[5], mean[5];
        for(int i = 0; i < 5; i++) {
            GPP_TRY(in[i].alloc(n));
            GPP_TRY(mean[i].alloc(n));
        }
        GPP_LAUNCH(gradient_moments_kernel, blocks, 128, 0, 0, d_b.ptr, d_v.ptr, (long long) n, in[0].ptr, in[1].ptr, in[2].ptr, in[3].ptr, in[4].ptr);
        for(int i = 0; i < 5; i++)   // meanX, meanY, meanXX, meanXY (Mean) and the number of valid pairs (Sum), :98-103
            GPP_TRY(gpp_neighbourhood_device(in[i].ptr, ny, nx, 0, ny, halfwidth, i < 4 ? GPP_MEAN : GPP_SUM, mean[i].ptr, nullptr));
        GPP_LAUNCH(gradient_regression_kernel, blocks, 128, 0, 0, mean[0].ptr, mean[1].ptr, mean[2].ptr, mean[3].ptr, mean[4].ptr, (long long) n, num_min,
                   min_range, default_gradient, d_out.ptr);
    }
    else {   // any other value: the reference returns the field of default gradients (calc_gradient.cpp:24,125)
        std::vector<float> fill(n, default_gradient);
        GPP_TRY(d_out.upload(fill.data(), n));
    }
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
