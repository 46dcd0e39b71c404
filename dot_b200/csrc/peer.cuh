// Device side of the peer-memory all-reduce (peer_reduce.cu): the views a producer / consumer kernel needs to fuse the push or the
// rank-ordered sum into its own pass, and the flag primitives.
#pragma once
#include "comm.h"

namespace dotgpu {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// every CTA of a consumer: wait until all ranks have published `epoch` (call with all threads of the CTA)
__device__ __forceinline__ void peer_wait_flags(const PeerSrc& S) {
    if ((int)threadIdx.x < S.world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(S.flags + threadIdx.x) - S.epoch) < 0) {
            __nanosleep(100);
            if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a peer died; fail instead of hanging the GPU
        }
    }
    __syncthreads();
}
// sum over the ranks in rank order (bit-identical on every rank); peers wrote these lines: never through this SM's L1
__device__ __forceinline__ double peer_sum(const PeerSrc& S, long long i) {
    double s = __ldcg(S.slots + i);
    for (int r = 1; r < S.nsum; ++r) s += __ldcg(S.slots + (long long)r * S.cap + i);
    return s;
}
// a producer's store of entry i of this rank's vector
__device__ __forceinline__ void peer_store(const PeerDst& D, long long i, double v) {
    if (D.slice) {
        const int s = (int)(i / D.slice);
        D.slot[s][i - (long long)s * D.slice] = v;
    } else {
        for (int r = 0; r < D.world; ++r) D.slot[r][i] = v;
    }
}
// last step of a producer kernel: every CTA calls it with all threads AFTER its stores into the peers' slots
__device__ __forceinline__ void peer_publish(const PeerDst& D) {
    __threadfence_system();
    __syncthreads();
    __shared__ bool peer_last;
    if (threadIdx.x == 0) peer_last = atomicAdd(D.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!peer_last) return;
    if ((int)threadIdx.x < D.world) {
        __threadfence_system();
        st_release_sys(D.flag[threadIdx.x], D.epoch);
    }
    if (threadIdx.x == 0) *D.counter = 0u;
}

}  // namespace dotgpu
