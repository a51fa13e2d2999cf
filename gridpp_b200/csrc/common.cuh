// Shared host/device helpers for libgridpp_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include "gridpp_b200.h"

namespace gpp {

// ------------------------------------------------------------------ errors ----------------------------
extern thread_local std::string g_last_error;
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define GPP_CUDA(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t err__ = (expr);                                                                       \
        if(err__ != cudaSuccess)                                                                          \
            return gpp::fail(GPP_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(err__), __FILE__, \
                             __LINE__, cudaGetErrorString(err__));                                        \
    } while(0)

#define GPP_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if(rc__ != GPP_OK) return rc__; \
    } while(0)

// every kernel launch of the library goes through this so that bench.py can report gpu_launches
#define GPP_LAUNCH(kernel, grid, block, smem, stream, ...)                \
    do {                                                                  \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);       \
        gpp::g_launches.fetch_add(1, std::memory_order_relaxed);          \
        GPP_CUDA(cudaGetLastError());                                     \
    } while(0)

// Opt-in phase timing of the host entry points: GPP_TRACE=1 prints wall-clock milliseconds per phase to stderr.
struct Trace {
    bool on;
    double t0;
    const char* what;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit Trace(const char* w) : what(w) {
        static const bool enabled = getenv("GPP_TRACE") != nullptr;
        on = enabled;
        t0 = on ? now() : 0;
    }
    void lap(const char* phase) {
        if(!on) return;
        double t = now();
        fprintf(stderr, "[gpp trace] %s: %s %.3f ms\n", what, phase, t - t0);
        t0 = t;
    }
};

int ensure_device();       // GPP_OK when a CUDA device is usable, else GPP_ERR_CUDA (no CPU fallback exists)
int sm_count();            // multiprocessors of the current device (148 on B200)

// RAII device buffer. Allocations come from the device's default stream-ordered memory pool, whose release
// threshold ensure_device() raises so that freed blocks stay cached: cudaMalloc/cudaFree of 64 MB fields were
// measured at up to 130 ms per call on the B200 box (profiles/e2e_probe.py), which dwarfed the kernels.
template <class T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t count = 0;
    DeviceBuffer() {}
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { release(); }
    void release() {
        if(ptr) cudaFreeAsync(ptr, 0);
        ptr = nullptr;
        count = 0;
    }
    int alloc(size_t n) {
        if(n <= count && ptr) return GPP_OK;
        release();
        if(n == 0) n = 1;
        GPP_CUDA(cudaMallocAsync((void**) &ptr, n * sizeof(T), 0));
        count = n;
        return GPP_OK;
    }
    int upload(const T* host, size_t n, cudaStream_t stream = 0) {
        GPP_TRY(alloc(n));
        if(n) GPP_CUDA(cudaMemcpyAsync(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        return GPP_OK;
    }
    int download(T* host, size_t n, cudaStream_t stream = 0) const {
        if(n) GPP_CUDA(cudaMemcpyAsync(host, ptr, n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        return GPP_OK;
    }
};

// ------------------------------------------------------------------ device numerics -------------------
// gridpp::is_valid, util.cpp:16-18 (MV is NaN, so `value != MV` is always true)
__host__ __device__ __forceinline__ bool is_valid(float v) { return !isnan(v) && !isinf(v); }

#ifdef __CUDACC__
// KDTree::calc_straight_distance, kdtree.cpp:192-194. Evaluated in float with individually rounded
// operations (no FMA contraction) so that it matches the reference's x86 build bit for bit.
__device__ __forceinline__ float straight_distance(float x0, float y0, float z0, float x1, float y1, float z1) {
    float dx = __fsub_rn(x0, x1), dy = __fsub_rn(y0, y1), dz = __fsub_rn(z0, z1);
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
// 1/x and 1/sqrt(x) from the hardware approximations plus two Newton steps: full double accuracy to a few ulp, a
// fraction of the cost of the correctly rounded forms. For pivots and rotation angles, whose last bit does not matter.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * fma(-0.5 * x * y, y, 1.5);
    y = y * fma(-0.5 * x * y, y, 1.5);
    return y;
}
__device__ __forceinline__ double shfl_double(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
#endif

}  // namespace gpp
