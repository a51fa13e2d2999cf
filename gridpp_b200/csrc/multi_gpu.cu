// Multi-GPU pieces under the C ABI.
//
//  * gpp_optimal_interpolation_multi_gpu_host: one PROCESS, several devices. Every output point of an OI analysis is
//    independent (oi.cpp:221-338), so the rows of the background grid are split into contiguous blocks, one per device; each
//    device gets its own copy of the (small) observation table and runs the ordinary single-device host path on its
//    block from its own host thread. No data-path collective. This is what a C++ program written against
//    include/gridpp.h uses to drive all the GPUs of a box (gridpp::b200::use_devices).
//  * gpp_halo_pull_device: the halo exchange of the row-tiled stencil filters as ONE kernel that reads the neighbours'
//    boundary rows straight out of their memory over NVLink (peer-mapped pointers: symmetric memory / CUDA IPC between
//    the one-process-per-GPU ranks, cudaDeviceEnablePeerAccess inside one process), instead of a grouped ncclSend /
//    ncclRecv whose launch overhead (~150 us) dwarfs the 480 KB it moves.
#include "points.cuh"

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

using namespace gpp;

namespace {

// 16-byte copies of 2 x hw boundary rows; rows are nx floats, nx % 4 == 0 and 16-byte aligned pointers take the vector path
__global__ void halo_pull_kernel(float* __restrict__ top_dst, const float* __restrict__ top_src, float* __restrict__ bottom_dst,
                                 const float* __restrict__ bottom_src, size_t n_floats, int vec) {
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    const size_t t0 = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(vec) {
        const size_t n4 = n_floats / 4;
        for(size_t i = t0; i < n4; i += stride) {
            if(top_src) reinterpret_cast<float4*>(top_dst)[i] = reinterpret_cast<const float4*>(top_src)[i];
            if(bottom_src) reinterpret_cast<float4*>(bottom_dst)[i] = reinterpret_cast<const float4*>(bottom_src)[i];
        }
    }
    else {
        for(size_t i = t0; i < n_floats; i += stride) {
            if(top_src) top_dst[i] = top_src[i];
            if(bottom_src) bottom_dst[i] = bottom_src[i];
        }
    }
}

}  // namespace

extern "C" {

int gpp_halo_pull_device(float* d_buf, int rows, int nx, int halfwidth, const float* d_from_above, const float* d_from_below, void* stream_) {
    cudaStream_t stream = (cudaStream_t) stream_;
    if(!d_buf || rows < 0 || nx < 0 || halfwidth < 0) return fail(GPP_ERR_INVALID_ARGUMENT, "bad argument");
    GPP_TRY(ensure_device());
    const size_t n = (size_t) halfwidth * nx;
    if(n == 0 || (!d_from_above && !d_from_below)) return GPP_OK;
    float* top = d_buf;                                             // halo rows [0, hw)
    float* bottom = d_buf + ((size_t) halfwidth + rows) * nx;      // halo rows [hw + rows, hw + rows + hw)
    const int vec = nx % 4 == 0 && (((uintptr_t) top | (uintptr_t) bottom | (uintptr_t) d_from_above | (uintptr_t) d_from_below) & 15) == 0;
    const size_t work = vec ? n / 4 : n;
    const unsigned grid = (unsigned) std::max<size_t>(1, std::min<size_t>((work + 255) / 256, (size_t) sm_count() * 4));
    GPP_LAUNCH(halo_pull_kernel, grid, 256, 0, stream, top, d_from_above, bottom, d_from_below, n, vec);
    return GPP_OK;
}

int gpp_optimal_interpolation_multi_gpu_host(int n_devices, const gpp_points* bpoints, const float* background, const float* bvariance,
                                             const gpp_points* opoints, const float* pobs, const float* obs_variance, const float* pbackground,
                                             const float* bvariance_at_points, const gpp_structure* structure, int max_points,
                                             int allow_extrapolation, float* analysis, float* analysis_variance) {
    if(!bpoints || !opoints) return fail(GPP_ERR_INVALID_ARGUMENT, "points must not be NULL");
    int available = 0;
    GPP_TRY(gpp_device_count(&available));
    if(available == 0) return fail(GPP_ERR_CUDA, "no usable CUDA device; libgridpp_b200 has no CPU fallback");
    int nd = n_devices <= 0 ? available : std::min(n_devices, available);
    const int nB = bpoints->n;
    const int nx = bpoints->shape_nx > 0 ? bpoints->shape_nx : 1;       // whole rows of a grid, single points of a point set
    const int n_rows = nB / nx;
    nd = std::max(1, std::min(nd, n_rows));
    if(nd == 1)
        return gpp_optimal_interpolation_host(bpoints, background, bvariance, opoints, pobs, obs_variance, pbackground, bvariance_at_points, structure,
                                              max_points, allow_extrapolation, analysis, analysis_variance);
    int home = 0;
    cudaGetDevice(&home);
    std::vector<int> rc((size_t) nd, GPP_OK);
    std::vector<std::string> msg((size_t) nd);
    std::vector<std::thread> workers;
    for(int d = 0; d < nd; d++) {
        workers.emplace_back([&, d]() {
            const int r0 = (int) ((long long) n_rows * d / nd), r1 = (int) ((long long) n_rows * (d + 1) / nd);
            const size_t first = (size_t) r0 * nx, count = (size_t) (r1 - r0) * nx;
            auto run = [&]() -> int {
                GPP_CUDA(cudaSetDevice(d));
                gpp_points *bp = nullptr, *op = nullptr;
                // this device's rows of the grid, and its own copy of the observation points
                GPP_TRY(gpp_points_create(bpoints->lats.data() + first, bpoints->lons.data() + first, bpoints->has_elevs ? bpoints->elevs.data() + first : nullptr,
                                          bpoints->has_lafs ? bpoints->lafs.data() + first : nullptr, (int) count, bpoints->type, &bp));
                int r = bpoints->shape_nx > 0 ? gpp_points_set_shape(bp, r1 - r0, nx) : GPP_OK;
                if(r == GPP_OK)
                    r = gpp_points_create(opoints->lats.data(), opoints->lons.data(), opoints->has_elevs ? opoints->elevs.data() : nullptr,
                                          opoints->has_lafs ? opoints->lafs.data() : nullptr, opoints->n, opoints->type, &op);
                if(r == GPP_OK)
                    r = gpp_optimal_interpolation_host(bp, background + first, bvariance ? bvariance + first : nullptr, op, pobs, obs_variance, pbackground,
                                                       bvariance_at_points, structure, max_points, allow_extrapolation, analysis + first,
                                                       analysis_variance ? analysis_variance + first : nullptr);
                gpp_points_destroy(bp);
                gpp_points_destroy(op);
                return r;
            };
            rc[d] = run();
            if(rc[d] != GPP_OK) msg[d] = gpp_last_error();      // the message lives in this thread
        });
    }
    for(std::thread& t : workers) t.join();
    cudaSetDevice(home);
    for(int d = 0; d < nd; d++)
        if(rc[d] != GPP_OK) return fail(rc[d], "device %d: %s", d, msg[d].c_str());
    return GPP_OK;
}

}  // extern "C"
