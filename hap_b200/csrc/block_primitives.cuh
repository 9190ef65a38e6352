// hap_b200/csrc/block_primitives.cuh -- block-wide scans used by the Snappy kernels.
#pragma once
#include "simt.h"

namespace hapb200 {

// Exclusive prefix sum over the NT threads of a block; *total receives the block sum.
// scratch: NT/32 words of shared memory.  Ends with a barrier, so scratch may be reused at once.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t *total, uint32_t *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(HAP_FULL_MASK, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) {
        uint32_t s = scratch[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

// Exclusive prefix maximum (identity 0); *total receives the block maximum.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_max(uint32_t v, uint32_t *total, uint32_t *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(HAP_FULL_MASK, incl, d);
        if (lane >= d) incl = incl > o ? incl : o;
    }
    uint32_t prev = __shfl_up_sync(HAP_FULL_MASK, incl, 1);
    if (lane == 0) prev = 0;
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) {
        uint32_t s = scratch[w];
        if (w < warp) base = base > s ? base : s;
        tot = tot > s ? tot : s;
    }
    __syncthreads();
    *total = tot;
    return base > prev ? base : prev;
}

}  // namespace hapb200
