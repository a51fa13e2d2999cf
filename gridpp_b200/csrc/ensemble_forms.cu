// The ensemble (vec3) forms of the neighbourhood filters.
//   gridpp::neighbourhood(vec3, halfwidth, statistic), src/api/neighbourhood.cpp:12-27: every cell's members are
//     reduced with calc_statistic (util.cpp:19-110) and the 2-D filter runs on the result;
//   gridpp::neighbourhood_quantile_fast(vec3, quantile | vec2, halfwidth, thresholds), neighbourhood.cpp:411-527:
//     per threshold the fraction of valid members <= threshold (a float) is averaged over the window with the Mean
//     filter, and the CDF is inverted with gridpp::interpolate.
// Both are thin kernels around the 2-D device entry points (which pick the TMA-staged kernels when the shape allows).
#include "common.cuh"

#include <algorithm>
#include <cstring>

using namespace gpp;

namespace {

// calc_statistic(members, statistic) for Mean / Sum / Count (float accumulation in member order, util.cpp:23-38)
// and Min / Max (calc_quantile(., 0 | 1), util.cpp:121-146). One thread per cell; members are contiguous.
__global__ void ens_statistic_kernel(const float* __restrict__ in, size_t n, int ne, int statistic, float* __restrict__ out) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float* a = in + i * ne;
    float value = NAN;
    if(statistic == GPP_MIN || statistic == GPP_MAX) {
        for(int e = 0; e < ne; e++) {
            const float v = a[e];
            if(!is_valid(v)) continue;
            if(!is_valid(value)) value = v;
            else if(statistic == GPP_MIN ? v < value : v > value) value = v;
        }
    }
    else {
        float total = 0.f;
        int count = 0;
        for(int e = 0; e < ne; e++) {
            const float v = a[e];
            if(is_valid(v)) { total = __fadd_rn(total, v); count++; }
        }
        if(statistic == GPP_COUNT) value = (float) count;
        else if(count > 0) value = statistic == GPP_MEAN ? __fdiv_rn(total, (float) count) : total;
    }
    out[i] = value;
}

// neighbourhood.cpp:452-470: temp = float(#valid members <= threshold) / #valid members, missing without valid members
__global__ void ens_fraction_kernel(const float* __restrict__ in, size_t n, int ne, float threshold, float* __restrict__ out) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float* a = in + i * ne;
    int sum = 0, count = 0;
    for(int e = 0; e < ne; e++) {
        const float v = a[e];
        if(is_valid(v)) { sum += v <= threshold ? 1 : 0; count++; }
    }
    out[i] = count > 0 ? __fdiv_rn((float) sum, (float) count) : NAN;
}

constexpr int ENS_MAX_T = 64;
struct EnsQuantileArgs {
    const float* stats;        // [T][n] window means of the fractions
    const float* qfield;       // may be NULL
    float* out;
    size_t n;
    int T, ne;
    float quantile;
    float thr[ENS_MAX_T];
};

// gridpp::interpolate, util.cpp:377-414, with get_lower_index / get_upper_index (util.cpp:339-376); iX has no missing values
__device__ float ens_interpolate(float x, const float* iX, const float* iY, int n) {
    if(!is_valid(x)) return NAN;
    if(x > iX[n - 1]) return iY[n - 1];
    if(x < iX[0]) return iY[0];
    int i0 = -1, i1 = -1;
    for(int i = 0; i < n; i++) {
        const float c = iX[i];
        if(c < x) i0 = i;
        else if(c == x) { i0 = i; break; }
        else if(c > x) break;
    }
    for(int i = n - 1; i >= 0; i--) {
        const float c = iX[i];
        if(c > x) i1 = i;
        else if(c == x) { i1 = i; break; }
        else if(c < x) break;
    }
    if(i0 < 0 || i1 < 0) return NAN;
    const float x0 = iX[i0], x1 = iX[i1], y0 = iY[i0], y1 = iY[i1];
    if(x0 == x1) {
        if(i0 == 0 && i1 == n - 1) return __fdiv_rn(__fadd_rn(y0, y1), 2.f);
        if(i0 == 0) return y1;
        if(i1 == n - 1) return y0;
        return __fdiv_rn(__fadd_rn(y0, y1), 2.f);
    }
    return __fadd_rn(y0, __fdiv_rn(__fmul_rn(__fsub_rn(y1, y0), __fsub_rn(x, x0)), __fsub_rn(x1, x0)));
}

// neighbourhood.cpp:483-521
__global__ void ens_quantile_finish_kernel(const __grid_constant__ EnsQuantileArgs a) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= a.n) return;
    float yarray[ENS_MAX_T];
    bool is_missing = false;
    for(int t = 0; t < a.T; t++) {
        const float st = a.stats[(size_t) t * a.n + i];
        if(is_valid(st)) {
            // the reference adds the same window mean once per member and divides by the member count (:496-503)
            float sum = 0.f;
            for(int e = 0; e < a.ne; e++) sum = __fadd_rn(sum, st);
            float y = __fdiv_rn(sum, (float) a.ne);
            y = y > 1.f ? 1.f : (y < 0.f ? 0.f : y);
            yarray[t] = y;
        }
        else {
            yarray[t] = NAN;
            is_missing = true;
        }
    }
    float result = NAN;
    if(!is_missing) {
        const float q = a.qfield ? a.qfield[i] : a.quantile;
        if(q == 1.f && yarray[0] == 1.f) result = a.thr[0];
        else if(q == 0.f && yarray[a.T - 1] == 0.f) result = a.thr[a.T - 1];
        else result = ens_interpolate(q, yarray, a.thr, a.T);
    }
    a.out[i] = result;
}

__global__ void ens_fill_kernel(float* out, size_t n, float value) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) out[i] = value;
}

unsigned blocks_for(size_t n) { return (unsigned) ((n + 255) / 256); }

}  // namespace

extern "C" {

int gpp_neighbourhood_ens_device(const float* d_input, int ny, int nx, int ne, int halfwidth, int statistic, float* d_output,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t) stream_;
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");                       // neighbourhood.cpp:29-30
    if(statistic == GPP_QUANTILE) return fail(GPP_ERR_INVALID_ARGUMENT, "Use neighbourhood_quantile for computing neighbourhood quantiles");
    GPP_TRY(ensure_device());
    if(ny <= 0 || nx <= 0 || ne <= 0) return GPP_OK;
    const size_t n = (size_t) ny * nx;
    float* flat = nullptr;
    GPP_CUDA(cudaMallocAsync((void**) &flat, sizeof(float) * n, stream));
    int rc = GPP_OK;
    if(statistic == GPP_MEAN || statistic == GPP_SUM || statistic == GPP_COUNT || statistic == GPP_MIN || statistic == GPP_MAX) {
        ens_statistic_kernel<<<blocks_for(n), 256, 0, stream>>>(d_input, n, ne, statistic, flat);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    else rc = gpp_calc_statistic_device(d_input, (long long) n, ne, statistic, flat, stream_);   // Median, Std, Variance, RandomChoice
    if(rc == GPP_OK) rc = gpp_neighbourhood_device(flat, ny, nx, 0, ny, halfwidth, statistic, d_output, stream_);
    cudaFreeAsync(flat, stream);
    if(rc == GPP_OK) GPP_CUDA(cudaGetLastError());
    return rc;
}

int gpp_neighbourhood_ens_host(const float* input, int ny, int nx, int ne, int halfwidth, int statistic, float* output) {
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");
    GPP_TRY(ensure_device());
    if(ny <= 0 || nx <= 0 || ne <= 0) return GPP_OK;
    const size_t n = (size_t) ny * nx;
    DeviceBuffer<float> d_in, d_out;
    GPP_TRY(d_in.upload(input, n * ne));
    GPP_TRY(d_out.alloc(n));
    GPP_TRY(gpp_neighbourhood_ens_device(d_in.ptr, ny, nx, ne, halfwidth, statistic, d_out.ptr, nullptr));
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

int gpp_neighbourhood_quantile_fast_ens_device(const float* d_input, int ny, int nx, int ne, float quantile, const float* d_quantile_field,
                                               int halfwidth, const float* thresholds, int num_thresholds, float* d_output, void* stream_) {
    cudaStream_t stream = (cudaStream_t) stream_;
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");                       // neighbourhood.cpp:418-419
    if(num_thresholds < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "negative number of thresholds");
    if(!d_quantile_field && is_valid(quantile) && (quantile < 0 || quantile > 1))
        return fail(GPP_ERR_INVALID_ARGUMENT, "All quantiles must be >= 0 and <= 1");                        // :433-440
    GPP_TRY(ensure_device());
    if(ny <= 0 || nx <= 0 || ne <= 0) return GPP_OK;                                                          // :421-422
    const size_t n = (size_t) ny * nx;
    if(num_thresholds == 0) {                                                                                  // :446-447
        ens_fill_kernel<<<blocks_for(n), 256, 0, stream>>>(d_output, n, NAN);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        GPP_CUDA(cudaGetLastError());
        return GPP_OK;
    }
    if(num_thresholds > ENS_MAX_T)
        return fail(GPP_ERR_NOT_IMPLEMENTED, "the ensemble form of neighbourhood_quantile_fast supports at most %d thresholds on the device", ENS_MAX_T);
    float *temp = nullptr, *stats = nullptr;
    GPP_CUDA(cudaMallocAsync((void**) &temp, sizeof(float) * n, stream));
    GPP_CUDA(cudaMallocAsync((void**) &stats, sizeof(float) * n * num_thresholds, stream));
    int rc = GPP_OK;
    for(int t = 0; t < num_thresholds && rc == GPP_OK; t++) {
        ens_fraction_kernel<<<blocks_for(n), 256, 0, stream>>>(d_input, n, ne, thresholds[t], temp);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = gpp_neighbourhood_device(temp, ny, nx, 0, ny, halfwidth, GPP_MEAN, stats + (size_t) t * n, stream_);   // :471
    }
    if(rc == GPP_OK) {
        EnsQuantileArgs a;
        std::memset(&a, 0, sizeof(a));
        a.stats = stats; a.qfield = d_quantile_field; a.out = d_output; a.n = n; a.T = num_thresholds; a.ne = ne; a.quantile = quantile;
        for(int t = 0; t < num_thresholds; t++) a.thr[t] = thresholds[t];
        ens_quantile_finish_kernel<<<blocks_for(n), 256, 0, stream>>>(a);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    cudaFreeAsync(temp, stream);
    cudaFreeAsync(stats, stream);
    if(rc == GPP_OK) GPP_CUDA(cudaGetLastError());
    return rc;
}

int gpp_neighbourhood_quantile_fast_ens_host(const float* input, int ny, int nx, int ne, float quantile, const float* quantile_field,
                                             int halfwidth, const float* thresholds, int num_thresholds, float* output) {
    if(halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "Half width must be > 0");
    GPP_TRY(ensure_device());
    if(ny <= 0 || nx <= 0 || ne <= 0) return GPP_OK;
    const size_t n = (size_t) ny * nx;
    if(quantile_field)
        for(size_t i = 0; i < n; i++)
            if(is_valid(quantile_field[i]) && (quantile_field[i] < 0 || quantile_field[i] > 1))
                return fail(GPP_ERR_INVALID_ARGUMENT, "All quantiles must be >= 0 and <= 1");
    DeviceBuffer<float> d_in, d_out, d_q;
    GPP_TRY(d_in.upload(input, n * ne));
    if(quantile_field) GPP_TRY(d_q.upload(quantile_field, n));
    GPP_TRY(d_out.alloc(n));
    GPP_TRY(gpp_neighbourhood_quantile_fast_ens_device(d_in.ptr, ny, nx, ne, quantile, quantile_field ? d_q.ptr : nullptr, halfwidth,
                                                       thresholds, num_thresholds, d_out.ptr, nullptr));
    GPP_TRY(d_out.download(output, n));
    GPP_CUDA(cudaStreamSynchronize(0));
    return GPP_OK;
}

}  // extern "C"
