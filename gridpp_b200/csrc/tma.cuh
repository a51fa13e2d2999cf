// Thin wrappers over the sm_100a bulk-tensor copy engine (TMA) and mbarrier PTX, plus host-side creation of the
// tensor map. The driver entry point cuTensorMapEncodeTiled is resolved at run time through the runtime API, so
// the library does not link against libcuda.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace gpp {

// 2-D row-major fp32 tensor of `rows` x `nx` elements at `base`; boxes of box_rows x box_cols elements. Elements of
// a box that fall outside the tensor are filled with NaN (nan_fill: gridpp's missing value) or zero, so a window
// clipped at the domain edge (neighbourhood.cpp:104-107) needs no special case in the kernels.
// Requirements of the copy engine: base 16-byte aligned, nx * 4 a multiple of 16, box_cols * 4 a multiple of 16,
// box dims <= 256, and the x coordinate of every box origin a multiple of 4 elements (16 bytes; measured: an
// unaligned origin raises an illegal-instruction fault).
int make_field_tensor_map(CUtensorMap* map, const float* base, int rows, int nx, int box_rows, int box_cols, bool nan_fill);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    while(!mbar_try_wait(bar, parity)) {}
}
// box whose first element is (x, y) of the tensor -> dense [box_rows][box_cols] tile at `dst` (128-byte aligned)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_descriptor(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace gpp
