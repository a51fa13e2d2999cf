// gridpp::get_neighbourhood_thresholds(vec2 | vec3, num_thresholds), src/api/neighbourhood.cpp:243-295: pool every valid
// value of the field, sort, and pick evenly spaced quantiles among the distinct values with gridpp::calc_even_quantiles
// (util.cpp:261-338). The pooling, the sort and the de-duplication (everything that touches all Y X values) run on the
// device (thrust / cub radix sort); the selection itself reads `num_thresholds` entries.
#include "common.cuh"

#include <thrust/binary_search.h>
#include <thrust/copy.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

using namespace gpp;

namespace {
struct IsValid {
    __device__ bool operator()(float v) const { return is_valid(v); }
};
__global__ void gather_kernel(const float* __restrict__ src, const int* __restrict__ index, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) out[i] = src[index[i]];
}
}  // namespace

extern "C" int gpp_get_neighbourhood_thresholds_host(const float* input, long long n_values, int num_thresholds, float* thresholds,
                                                     int* num_out) {
    if(num_out) *num_out = 0;
    if(num_thresholds <= 0) return fail(GPP_ERR_INVALID_ARGUMENT, "num_thresholds must be > 0");   // neighbourhood.cpp:244-245
    if(!num_out || !thresholds) return fail(GPP_ERR_INVALID_ARGUMENT, "NULL output");
    GPP_TRY(ensure_device());
    if(n_values <= 0) return GPP_OK;                                                                // :247-248
    const size_t n = (size_t) n_values;
    DeviceBuffer<float> d_in, d_valid, d_uniq, d_pick;
    DeviceBuffer<int> d_index;
    GPP_TRY(d_in.upload(input, n));
    GPP_TRY(d_valid.alloc(n));
    GPP_TRY(d_uniq.alloc(n));
    cudaStream_t stream = 0;
    auto policy = thrust::cuda::par.on(stream);
    // all_values: the valid values, sorted (neighbourhood.cpp:255-263)
    float* v_end = thrust::copy_if(policy, d_in.ptr, d_in.ptr + n, d_valid.ptr, IsValid());
    const long long size = v_end - d_valid.ptr;
    g_launches.fetch_add(3, std::memory_order_relaxed);
    if(size == 0) return GPP_OK;                       // calc_even_quantiles: nothing to pick from (util.cpp:264-265)
    thrust::sort(policy, d_valid.ptr, v_end);
    float* u_end = thrust::unique_copy(policy, d_valid.ptr, v_end, d_uniq.ptr);
    const long long n_uniq = u_end - d_uniq.ptr;
    GPP_CUDA(cudaGetLastError());
    const int num = num_thresholds;
    std::vector<int> index;                            // positions in the distinct-value array, in output order
    if((long long) num >= size) {                      // util.cpp:271-280: every distinct value
        for(long long i = 0; i < n_uniq; i++) index.push_back((int) i);
    }
    else {
        // sorted[0] = uniq[0], sorted[size-1] = uniq[n_uniq-1]; count_lower = multiplicity of the lowest value
        long long count_lower = size;
        if(n_uniq > 1) {
            float second;
            GPP_CUDA(cudaMemcpyAsync(&second, d_uniq.ptr + 1, sizeof(float), cudaMemcpyDeviceToHost, stream));
            GPP_CUDA(cudaStreamSynchronize(stream));
            count_lower = thrust::lower_bound(policy, d_valid.ptr, v_end, second) - d_valid.ptr;
        }
        index.push_back(0);                            // lowest
        if(num == 2) {                                 // util.cpp:296-300
            if(n_uniq > 1) index.push_back((int) (n_uniq - 1));
        }
        else {
            // util.cpp:302-308: the first value past a long run of the lowest value
            const bool repeated_at_beginning = count_lower < size && count_lower > size / num;
            long long first_remaining = 1;             // distinct values strictly above last_added start here
            if(repeated_at_beginning) { index.push_back(1); first_remaining = 2; }
            const long long n_remaining = n_uniq - first_remaining;   // remaining_unique_values.size(), util.cpp:310-316
            if(n_remaining > 0) {
                const int num_left = num - (int) index.size();
                for(int i = 1; i <= num_left; i++) {
                    const float f = float(i) / (num_left);
                    const int idx = (int) ((float) n_remaining * f - 1);   // util.cpp:322 (int * float - 1, truncated)
                    if(idx < 0) return fail(GPP_ERR_RUNTIME, "Internal error in calc_even_quantiles.");
                    index.push_back((int) (first_remaining + idx));
                }
            }
        }
    }
    const int n_out = (int) index.size();
    if(n_out > 0) {
        GPP_TRY(d_index.upload(index.data(), n_out));
        GPP_TRY(d_pick.alloc(n_out));
        GPP_LAUNCH(gather_kernel, (unsigned) ((n_out + 255) / 256), 256, 0, stream, d_uniq.ptr, d_index.ptr, n_out, d_pick.ptr);
        GPP_TRY(d_pick.download(thresholds, n_out));
        GPP_CUDA(cudaStreamSynchronize(stream));
    }
    *num_out = n_out;
    return GPP_OK;
}
