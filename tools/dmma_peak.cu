// fp64 peaks of the GPU this runs on, for the "tensor-pipe % of peak" figures of K6 (profiles/README.md):
//   dmma   : mma.sync.aligned.m8n8k4.f64 issued back to back from registers (16 independent accumulator pairs per warp)
//   dfma   : plain FMA.f64 from registers (8 independent chains per thread)
//   syrk   : C[64x64 tile] -= A A^T with the K6 tile mainloop variants (operands through shared memory), one CTA per tile,
//            long K, as `k_update_cb` runs it: v0 = single-buffered (the round-1/2 kernel), v1 = cp.async double-buffered
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_peak tools/dmma_peak.cu ; run: tools/dmma_peak [json]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0) {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a0, double b0) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = i;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------ tile mainloops ------------------------------------------------
constexpr int SLD = 36, KC = 32;
__device__ __forceinline__ void load_chunk(double* sX, const double* __restrict__ src, long long ld, int nrows, int kw) {
    const int k = threadIdx.x & 31, r0 = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int r = r0 + 4 * i;
        double v = 0.0;
        if (r < nrows && k < kw) v = src[(long long)r * ld + k];
        sX[r * SLD + k] = v;
    }
}
__device__ __forceinline__ void mma_chunk(double (&acc)[4][4][2], const double* sA, const double* sB, int wr, int wc, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int k0 = 0; k0 < KC; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sA[(wr * 32 + i * 8 + g) * SLD + k0 + q];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sB[(wc * 32 + j * 8 + g) * SLD + k0 + q];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// v0: the kernel as shipped in rounds 1-2 (64x64 tile, 4 warps, single buffer)
__global__ void __launch_bounds__(128) k_syrk_v0(const double* __restrict__ P, int nb, int ns, double* __restrict__ C) {
    __shared__ double sA[64 * SLD], sB[64 * SLD];
    const int ta = blockIdx.x / (nb / 64), tb = blockIdx.x % (nb / 64);
    if (tb > ta) return;
    const double* __restrict__ A = P + (long long)ta * 64 * ns;
    const double* __restrict__ B = P + (long long)tb * 64 * ns;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp >> 1, wc = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < ns; k0 += KC) {
        __syncthreads();
        load_chunk(sA, A + k0, ns, 64, ns - k0);
        load_chunk(sB, B + k0, ns, 64, ns - k0);
        __syncthreads();
        mma_chunk(acc, sA, sB, wr, wc, lane);
    }
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int r = ta * 64 + wr * 32 + i * 8 + g, c = tb * 64 + wc * 32 + j * 8 + 2 * q + e;
                if (r >= c) C[(long long)r * nb + c] -= acc[i][j][e];
            }
}

// v1: TM x 64 tile (TM = 64 or 128), TM/16 warps (each 32x32), cp.async ring of NST stages of [TM + 64][KC] operand chunks
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int ROWS, int NT>
__device__ __forceinline__ void cp_chunk(double* sX, const double* __restrict__ src, long long ld, int nrows, int kw) {
    // ROWS x KC doubles = ROWS * 16 copies of 16 B; rows / columns past the end are zero-filled (src-size 0)
    for (int e = threadIdx.x; e < ROWS * 16; e += NT) {
        const int r = e >> 4, k = (e & 15) * 2;
        const bool in = r < nrows && k < kw;   // kw even (ns even) in this testbed; the library version handles odd tails with 8-B copies
        const double* s = in ? src + (long long)r * ld + k : src;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sX + r * SLD + k)), "l"(s), "r"(in ? 16 : 0) : "memory");
    }
}
template <int TM, int NST>
__global__ void __launch_bounds__(TM * 2) k_syrk_v1(const double* __restrict__ P, int nb, int ns, double* __restrict__ C) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = TM * 2, STG = (TM + 64) * SLD;
    const int nta = nb / TM, ntb = nb / 64;
    const int ta = blockIdx.x / ntb, tb = blockIdx.x % ntb;
    if (tb * 64 > ta * TM + TM - 1 || ta >= nta) return;
    const double* __restrict__ A = P + (long long)ta * TM * ns;
    const double* __restrict__ B = P + (long long)tb * 64 * ns;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp >> 1, wc = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nk = (ns + KC - 1) / KC;
#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nk) {
            cp_chunk<TM, NT>(sm + s * STG, A + s * KC, ns, TM, ns - s * KC);
            cp_chunk<64, NT>(sm + s * STG + TM * SLD, B + s * KC, ns, 64, ns - s * KC);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kc = 0; kc < nk; ++kc) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
        __syncthreads();   // chunk kc landed for everybody; everybody is done with chunk kc-1 (whose buffer is refilled next)
        const int nx = kc + NST - 1;
        if (nx < nk) {
            double* dst = sm + (nx % NST) * STG;
            cp_chunk<TM, NT>(dst, A + nx * KC, ns, TM, ns - nx * KC);
            cp_chunk<64, NT>(dst + TM * SLD, B + nx * KC, ns, 64, ns - nx * KC);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const double* sA = sm + (kc % NST) * STG;
        mma_chunk(acc, sA, sA + TM * SLD, wr, wc, lane);
    }
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int r = ta * TM + wr * 32 + i * 8 + g, c = tb * 64 + wc * 32 + j * 8 + 2 * q + e;
                if (r >= c) C[(long long)r * nb + c] -= acc[i][j][e];
            }
}

template <class F>
static float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    double* out;
    CK(cudaMalloc(&out, (size_t)nsm * 8 * 256 * sizeof(double)));
    const int iters = 4096;
    double best_dmma = 0, best_dfma = 0;
    int best_w = 0;
    for (int cps = 1; cps <= 8; cps *= 2) {  // CTAs of 8 warps per SM
        float ms = time_ms([&] { k_dmma<<<nsm * cps, 256>>>(out, iters, 1.0, 1e-3); }, 5);
        double tf = (double)nsm * cps * 8 * iters * 16 * 512.0 / (ms * 1e-3) / 1e12;
        std::printf("dmma m8n8k4 f64: %d warps/SM  %.3f ms  %.2f TFLOP/s\n", cps * 8, ms, tf);
        if (tf > best_dmma) { best_dmma = tf; best_w = cps * 8; }
        ms = time_ms([&] { k_dfma<<<nsm * cps, 256>>>(out, iters * 4, 1.0000001, 1e-3); }, 5);
        tf = (double)nsm * cps * 256 * (iters * 4.0) * 8 * 2.0 / (ms * 1e-3) / 1e12;
        std::printf("dfma f64        : %d warps/SM  %.3f ms  %.2f TFLOP/s\n", cps * 8, ms, tf);
        if (tf > best_dfma) best_dfma = tf;
    }
    // ---- tile mainloops: nb x nb contribution block (lower tiles), K = ns ----
    struct Res { const char* name; int nb, ns; double tf; };
    std::vector<Res> res;
    const int shapes[][2] = {{2048, 512}, {4096, 1024}, {1024, 192}, {8192, 64}};
    for (auto& sh : shapes) {
        const int nb = sh[0], ns = sh[1];
        double *P, *C;
        CK(cudaMalloc(&P, (size_t)nb * ns * 8));
        CK(cudaMalloc(&C, (size_t)nb * nb * 8));
        std::vector<double> h((size_t)nb * ns);
        for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 2654435761u) % 1000) * 1e-3 - 0.5;
        CK(cudaMemcpy(P, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemset(C, 0, (size_t)nb * nb * 8));
        const double flops = (double)nb * nb * ns;  // lower triangle only: nb^2/2 entries x 2 ns
        auto rec = [&](const char* name, float ms) {
            std::printf("syrk %-22s nb %5d ns %5d  %.3f ms  %.2f TFLOP/s\n", name, nb, ns, ms, flops / (ms * 1e-3) / 1e12);
            res.push_back({name, nb, ns, flops / (ms * 1e-3) / 1e12});
        };
        rec("v0 64x64 single", time_ms([&] { k_syrk_v0<<<(nb / 64) * (nb / 64), 128>>>(P, nb, ns, C); }, 5));
        std::vector<double> c0((size_t)nb * nb), c1((size_t)nb * nb);
        CK(cudaMemset(C, 0, (size_t)nb * nb * 8));
        k_syrk_v0<<<(nb / 64) * (nb / 64), 128>>>(P, nb, ns, C);
        CK(cudaMemcpy(c0.data(), C, c0.size() * 8, cudaMemcpyDeviceToHost));
        auto check = [&](const char* name) {
            CK(cudaMemcpy(c1.data(), C, c1.size() * 8, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < c0.size(); ++i) bad += c0[i] != c1[i];
            if (bad) std::printf("   !! %s differs from v0 in %zu entries\n", name, bad);
        };
#define RUN_V1(TM, NST, NAME)                                                                                              \
    {                                                                                                                      \
        const size_t smem = (size_t)NST * (TM + 64) * SLD * 8;                                                             \
        CK(cudaFuncSetAttribute(k_syrk_v1<TM, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
        rec(NAME, time_ms([&] { k_syrk_v1<TM, NST><<<(nb / TM) * (nb / 64), TM * 2, smem>>>(P, nb, ns, C); }, 5));          \
        CK(cudaMemset(C, 0, (size_t)nb * nb * 8));                                                                          \
        k_syrk_v1<TM, NST><<<(nb / TM) * (nb / 64), TM * 2, smem>>>(P, nb, ns, C);                                          \
        check(NAME);                                                                                                       \
    }
        RUN_V1(64, 2, "v1 64x64 cp.async x2");
        RUN_V1(64, 3, "v1 64x64 cp.async x3");
        RUN_V1(64, 4, "v1 64x64 cp.async x4");
        RUN_V1(128, 3, "v1 128x64 cp.async x3");
        RUN_V1(128, 4, "v1 128x64 cp.async x4");
        CK(cudaFree(P));
        CK(cudaFree(C));
    }
    if (argc > 1) {
        FILE* f = std::fopen(argv[1], "w");
        std::fprintf(f, "{\"device\": \"%s\", \"sms\": %d, \"dmma_m8n8k4_f64_tflops\": %.3f, \"dmma_best_warps_per_sm\": %d, \"dfma_f64_tflops\": %.3f, \"syrk\": [",
                     prop.name, nsm, best_dmma, best_w, best_dfma);
        for (size_t i = 0; i < res.size(); ++i)
            std::fprintf(f, "%s{\"kernel\": \"%s\", \"nb\": %d, \"ns\": %d, \"tflops\": %.3f}", i ? ", " : "", res[i].name, res[i].nb, res[i].ns, res[i].tf);
        std::fprintf(f, "]}\n");
        std::fclose(f);
    }
    return 0;
}
