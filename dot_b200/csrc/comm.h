// NCCL all-reduce over NVLink for the one exchange step of the path: the sum of the subdomains'
// search-direction contributions (DOTTimeStepper.cpp:434-450 is a serial shared-memory loop in the
// reference).  libnccl is loaded at run time (dlopen) so that single-GPU use has no NCCL dependency.
#pragma once
#include "common.h"

namespace dotgpu {
struct Comm {
    void* lib = nullptr;
    void* comm = nullptr;
    int rank = 0, world = 1;
    ~Comm();
    static void unique_id(void* out128);
    void init(const void* unique_id128, int rank, int world);
    void all_reduce_sum(double* buf, long long n, cudaStream_t st);
};
}  // namespace dotgpu
