// Deterministic multi-value reductions shared by the fused vector kernels: every CTA reduces NACC per-thread accumulators
// with independent shuffle trees + one barrier, stores one partial per value, and the last CTA to finish (atomic ticket)
// adds the partials of every value in a fixed order.  Result j lands in res[j] (shared memory) of the last CTA.
#pragma once
#include <cuda_runtime.h>

namespace dotgpu {

// 256-thread CTAs.  sh: [8 * NACC] doubles of shared memory.  Returns true in the last CTA, with res[0..nacc) valid
// for all its threads; false elsewhere.
template <int NACC>
__device__ __forceinline__ bool multi_reduce_256(const double (&acc)[NACC], int nacc, double* sh, double* res, bool* last_flag,
                                                 double* __restrict__ partial, unsigned* __restrict__ counter) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
        if (j < nacc) {
            double v = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) sh[warp * NACC + j] = v;
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < nacc) {
        const int j = threadIdx.x;
        const double s = ((sh[j] + sh[NACC + j]) + (sh[2 * NACC + j] + sh[3 * NACC + j])) +
                         ((sh[4 * NACC + j] + sh[5 * NACC + j]) + (sh[6 * NACC + j] + sh[7 * NACC + j]));
        partial[(size_t)j * gridDim.x + blockIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *last_flag = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!*last_flag) return false;
    __threadfence();
    // warp w adds the partials of values w, w+8, ...: lanes stride over the CTAs, then a shuffle tree
    for (int j = warp; j < nacc; j += 8) {
        double v = 0.0;
        for (int i = lane; i < (int)gridDim.x; i += 32) v += __ldcg(partial + (size_t)j * gridDim.x + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) res[j] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0u;
    return true;
}

}  // namespace dotgpu
