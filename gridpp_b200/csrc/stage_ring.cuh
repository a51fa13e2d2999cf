// The staging ring shared by the copy-engine (TMA) neighbourhood kernels: a CTA of NT threads owns a strip of NT
// staged columns and keeps NS stages of RB rows in flight, one cp.async.bulk.tensor copy and one mbarrier per stage.
#pragma once

#include "tma.cuh"

namespace gpp {
namespace nbh {

constexpr int NT = 256;              // threads per CTA = staged columns per strip
constexpr int RB = 8;                // rows per stage = output rows per batch
constexpr unsigned STAGE_BYTES = RB * NT * sizeof(float);

#ifdef __CUDACC__
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

#endif

}  // namespace nbh
}  // namespace gpp
