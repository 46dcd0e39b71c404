// Shared host-side helpers of libdotgpu: error codes, CUDA checks, device buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dotgpu.h"

namespace dotgpu {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);

#define DG_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            throw ::dotgpu::Error(DOTGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + \
                                                       " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define DG_REQUIRE(cond, msg)                                            \
    do {                                                                 \
        if (!(cond)) throw ::dotgpu::Error(DOTGPU_ERR_INVALID, (msg));   \
    } while (0)

// Every kernel launch of the library goes through this counter (bench.py reports it as gpu_launches).
extern thread_local int64_t g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) DG_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
    }
    void upload(const T* h, size_t cnt, cudaStream_t st = 0) {
        if (cnt > n) alloc(cnt);
        if (cnt) DG_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    void upload(const std::vector<T>& h, cudaStream_t st = 0) {
        upload(h.data(), h.size(), st);
        DG_CUDA(cudaStreamSynchronize(st));  // the host vector may die right after
    }
    void download(T* h, size_t cnt, cudaStream_t st = 0) const {
        if (cnt) DG_CUDA(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, st));
        DG_CUDA(cudaStreamSynchronize(st));
        DG_CUDA(cudaGetLastError());  // a kernel launch rejected earlier on this thread must not pass silently
    }
    void zero(cudaStream_t st = 0) {
        if (n) DG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    }
    size_t bytes() const { return n * sizeof(T); }
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace dotgpu
