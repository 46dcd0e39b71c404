#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "linalg.h"
#include "peer.cuh"
#include "reduce.cuh"

namespace dotgpu {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// deterministic CTA sum for 256 threads, valid in thread 0
__device__ __forceinline__ double cta_sum256(double v, double* sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) r = ((sh[0] + sh[1]) + (sh[2] + sh[3])) + ((sh[4] + sh[5]) + (sh[6] + sh[7]));
    return r;
}

__global__ void __launch_bounds__(128) k_fill(long long nblk, const long long* __restrict__ ptr, const int* __restrict__ src,
                                              const int* __restrict__ row, const int* __restrict__ bj, const double* __restrict__ consts,
                                              const int* __restrict__ ia, const double* __restrict__ He, double* __restrict__ a) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.0;
    for (long long e = ptr[b]; e < ptr[b + 1]; ++e) {
        int code = src[e];
        if (code >= 0) {
            // code = 16*tet + 4*a + b; only blocks a <= b are stored ([nT][10][9]), (a,b) with a > b is the transpose of (b,a)
            const int tet = code >> 4, ba = (code >> 2) & 3, bb = code & 3;
            const int lo = min(ba, bb), hi = max(ba, bb);
            const double* __restrict__ h = He + (long long)tet * 90 + (lo * 4 - lo * (lo - 1) / 2 + (hi - lo)) * 9;
            if (ba <= bb) {
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += h[i];
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int r = 0; r < 3; ++r) acc[3 * i + r] += h[3 * r + i];
            }
        } else {
            double c = consts[-code - 1];
            acc[0] += c; acc[4] += c; acc[8] += c;
        }
    }
    const int r = row[b], j = bj[b];
    if (j < 0) {  // fixed vertex: three 1x1 rows
        a[ia[r]] = acc[0];
        a[ia[r + 1]] = acc[4];
        a[ia[r + 2]] = acc[8];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int base = ia[r + i] + 3 * j - i;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (j > 0 || c >= i) a[base + c] = acc[3 * i + c];
        }
    }
}

__global__ void __launch_bounds__(256) k_quadform(int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ a,
                                                  const double* __restrict__ p, double* __restrict__ partial) {
    __shared__ double sh[8];
    int i = blockIdx.x * 256 + threadIdx.x;
    double s = 0.0;
    if (i < n) {
        const double pi = p[i];
        int b = ia[i], e = ia[i + 1];
        double off = 0.0;
        for (int k = b + 1; k < e; ++k) off += a[k] * p[ja[k]];
        s = pi * (a[b] * pi + 2.0 * off);  // first entry of every row is the diagonal
    }
    double r = cta_sum256(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    double r = cta_sum256(s, sh);
    if (threadIdx.x == 0) out[0] = r;
}

__global__ void __launch_bounds__(256) k_spmv_sym(int n, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ a,
                                                  const int* __restrict__ tp, const int* __restrict__ tslot, const int* __restrict__ trow,
                                                  const double* __restrict__ x, double* __restrict__ y) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = tp[i]; k < tp[i + 1]; ++k) s += a[tslot[k]] * x[trow[k]];  // strictly-lower part via the transposed index
    for (int k = ia[i]; k < ia[i + 1]; ++k) s += a[k] * x[ja[k]];
    y[i] = s;
}

constexpr int DOT_TPB = 256;
constexpr int DOT_MAX_BLOCKS = 592;  // 4 CTAs per SM

__global__ void __launch_bounds__(DOT_TPB) k_dot(long long n, const double* __restrict__ a, const double* __restrict__ b,
                                                 double* __restrict__ partial, unsigned* __restrict__ counter, double* __restrict__ out) {
    __shared__ double sh[8];
    __shared__ bool last;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * DOT_TPB + threadIdx.x; i < n; i += (long long)gridDim.x * DOT_TPB) s += a[i] * b[i];
    double r = cta_sum256(s, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = r;
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {  // the last CTA to finish sums the partials in a fixed order
        __threadfence();
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += DOT_TPB) v += ((volatile double*)partial)[i];
        __syncthreads();
        double tot = cta_sum256(v, sh);
        if (threadIdx.x == 0) {
            out[0] = tot;
            *counter = 0u;
        }
    }
}




__global__ void k_axpy(long long n, double* __restrict__ out, const double* __restrict__ x0, const double* __restrict__ p, double alpha) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x0[i] + alpha * p[i];
}
__global__ void k_warm_start(int nV, double* __restrict__ x, const double* __restrict__ vel, const unsigned char* __restrict__ fixed, double dt,
                             double gx, double gy, double gz) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV || fixed[v]) return;
    x[3 * (size_t)v] = x[3 * (size_t)v] + 1.0 * (dt * vel[3 * (size_t)v] + gx);
    x[3 * (size_t)v + 1] = x[3 * (size_t)v + 1] + 1.0 * (dt * vel[3 * (size_t)v + 1] + gy);
    x[3 * (size_t)v + 2] = x[3 * (size_t)v + 2] + 1.0 * (dt * vel[3 * (size_t)v + 2] + gz);
}
__global__ void k_xtilde(int nV, double* __restrict__ xt, const double* __restrict__ xn, const double* __restrict__ vel,
                         const unsigned char* __restrict__ fixed, double dt, double gx, double gy, double gz) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nV) return;
    const bool f = fixed[v];
    xt[3 * (size_t)v] = f ? xn[3 * (size_t)v] : xn[3 * (size_t)v] + (vel[3 * (size_t)v] * dt + gx);
    xt[3 * (size_t)v + 1] = f ? xn[3 * (size_t)v + 1] : xn[3 * (size_t)v + 1] + (vel[3 * (size_t)v + 1] * dt + gy);
    xt[3 * (size_t)v + 2] = f ? xn[3 * (size_t)v + 2] : xn[3 * (size_t)v + 2] + (vel[3 * (size_t)v + 2] * dt + gz);
}
__global__ void k_velocity(long long n, double* __restrict__ vel, const double* __restrict__ x, const double* __restrict__ xn, double dt) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vel[i] = (x[i] - xn[i]) / dt;
}
__global__ void k_scatter_avg(int ndof, const int* __restrict__ cptr, const int* __restrict__ cidx, const double* __restrict__ xs,
                              const int* __restrict__ dup, double* __restrict__ p) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndof) return;
    double s = 0.0;
    for (int k = cptr[d]; k < cptr[d + 1]; ++k) s += xs[cidx[k]];
    int du = dup ? dup[d / 3] : 1;
    if (du > 1) s /= (double)du;
    p[d] = s;
}

}  // namespace

void launch_fill(const DeviceFill& f, const double* He, double* a, cudaStream_t st) {
    if (!f.nblk) return;
    k_fill<<<ceil_div(f.nblk, 128), 128, 0, st>>>(f.nblk, f.ptr.p, f.src.p, f.row.p, f.j.p, f.consts.p, f.ia.p, He, a);
    count_launch();
}

void launch_quadform(int n, const int* ia, const int* ja, const double* a, const double* p, double* partial, double* out, cudaStream_t st) {
    int nb = ceil_div(n, 256);
    k_quadform<<<nb, 256, 0, st>>>(n, ia, ja, a, p, partial);
    k_sum_partials<<<1, 256, 0, st>>>(partial, nb, out);
    count_launch(2);
}

void launch_spmv_sym(int n, const int* ia, const int* ja, const double* a, const int* tp, const int* tslot, const int* trow, const double* x,
                     double* y, cudaStream_t st) {
    k_spmv_sym<<<ceil_div(n, 256), 256, 0, st>>>(n, ia, ja, a, tp, tslot, trow, x, y);
    count_launch();
}

int dot_partial_count(long long) { return DOT_MAX_BLOCKS; }

void launch_dot(long long n, const double* a, const double* b, double* partial, unsigned* counter, double* sc_out, cudaStream_t st) {
    int nb = (int)std::min<long long>(DOT_MAX_BLOCKS, std::max<long long>(1, (n + DOT_TPB * 4 - 1) / (DOT_TPB * 4)));
    k_dot<<<nb, DOT_TPB, 0, st>>>(n, a, b, partial, counter, sc_out);
    count_launch();
}

// ------------------------------------------------------------------------------------------------------------------
// fused L-BFGS kernels
constexpr int MD_TPB = 256;
static int md_max_blocks() {  // grid of the multi-reduction kernels (DOTGPU_MD_BLOCKS: experiments)
    static const int v = [] {
        const char* e = std::getenv("DOTGPU_MD_BLOCKS");
        return e && *e ? std::max(1, std::atoi(e)) : 592;  // 4 CTAs per SM (B200 sweep r2: 296 -> 592 -> 1184: 76.9 -> 76.1 -> 76.7 ms per frame at 1M tets)
    }();
    return v;
}

__global__ void __launch_bounds__(MD_TPB) k_dots(long long n, DotPairs P, double* __restrict__ partial, unsigned* __restrict__ counter,
                                                 double* __restrict__ sc) {
    if (P.go && *P.go == 0) return;
    __shared__ double sh[8 * 12], res[12];
    __shared__ bool last;
    double acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.0;
    for (long long i = (long long)blockIdx.x * MD_TPB + threadIdx.x; i < n; i += (long long)gridDim.x * MD_TPB) {
#pragma unroll
        for (int j = 0; j < 12; ++j)
            if (j < P.n) acc[j] += P.a[j][i] * P.b[j][i];
    }
    if (!multi_reduce_256<12>(acc, P.n, sh, res, &last, partial, counter)) return;
    if ((int)threadIdx.x < P.n) sc[P.out[threadIdx.x]] = res[threadIdx.x];
}

// compact first loop: xi_i = (s_i . q_i) / (y_i . s_i),  s_i . q_i = -(s_i . g) - sum_{j newer than i} xi_j (s_i . y_j)
__device__ __forceinline__ void lbfgs_xi(const double* __restrict__ sc, const HistList& H, double* xi) {
    for (int i = H.n - 1; i >= 0; --i) {
        const int si = H.slot[i];
        double v = -sc[SC_SG + si];
        for (int j = H.n - 1; j > i; --j) v -= xi[j] * sc[SC_SY + 8 * si + H.slot[j]];
        xi[i] = v / sc[SC_SY + 8 * si + si];
    }
}

__global__ void __launch_bounds__(256) k_lbfgs_q(long long n, double* __restrict__ q, const double* __restrict__ g, HistList H,
                                                 double* __restrict__ sc) {
    if (H.go && *H.go == 0) return;
    __shared__ double xi[LB_MAXH];
    if (threadIdx.x == 0) {
        lbfgs_xi(sc, H, xi);
        if (blockIdx.x == 0)
            for (int i = 0; i < H.n; ++i) sc[SC_XI + H.slot[i]] = xi[i];
    }
    __syncthreads();
    long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= n) return;
    double v = -g[e];
    for (int i = H.n - 1; i >= 0; --i) v -= xi[i] * H.Y[i][e];
    q[e] = v;
}

__global__ void __launch_bounds__(256) k_lbfgs_p(long long n, double* __restrict__ p, HistList H, double* __restrict__ sc) {
    if (H.go && *H.go == 0) return;
    __shared__ double c[LB_MAXH];
    if (threadIdx.x == 0) {
        // compact second loop: beta_i = (y_i . p_i) / (y_i . s_i),  y_i . p_i = y_i . p0 + sum_{j older than i} c_j (y_i . s_j),  c_i = xi_i - beta_i
        double pg = sc[SC_P0G];
        for (int i = 0; i < H.n; ++i) {
            const int si = H.slot[i];
            double v = sc[SC_YP + si];
            for (int j = 0; j < i; ++j) v += c[j] * sc[SC_SY + 8 * H.slot[j] + si];
            c[i] = sc[SC_XI + si] - v / sc[SC_SY + 8 * si + si];
            pg += c[i] * sc[SC_SG + si];
        }
        if (blockIdx.x == 0) sc[SC_PG] = pg;
    }
    __syncthreads();
    long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= n) return;
    double v = p[e];
    for (int i = 0; i < H.n; ++i) v += c[i] * H.S[i][e];
    p[e] = v;
}

// p^T H p with the symmetric matrix stored as CSR-upper: QF_LANES lanes share a row so that the value / column-index reads of a warp
// are 4 contiguous segments instead of 32 scattered ones (a row holds ~23 entries; one thread per row ran at 0.4 of HBM)
constexpr int QF_LANES = 8, QF_ROWS = 256 / QF_LANES;
__global__ void __launch_bounds__(256) k_quadform_alpha(int n, const int* __restrict__ ia, const int* __restrict__ ja,
                                                        const double* __restrict__ a, const double* __restrict__ p,
                                                        double* __restrict__ partial, unsigned* __restrict__ counter, double* __restrict__ sc,
                                                        const int* __restrict__ go, int npass) {
    if (go && *go == 0) return;
    __shared__ double sh[8];
    __shared__ bool last;
    const int sl = threadIdx.x & (QF_LANES - 1), grp = threadIdx.x / QF_LANES;
    double s = 0.0;
#pragma unroll 2
    for (int rr = 0; rr < npass; ++rr) {   // a CTA covers npass * QF_ROWS consecutive rows, QF_ROWS at a time
        const int i = (blockIdx.x * npass + rr) * QF_ROWS + grp;
        int b = 0, e = 0;
        if (i < n) {
            b = ia[i];
            e = ia[i + 1];
        }
        double off = 0.0;
        for (int k = b + 1 + sl; k < e; k += QF_LANES) off += a[k] * p[ja[k]];
#pragma unroll
        for (int o = QF_LANES / 2; o > 0; o >>= 1) off += __shfl_down_sync(0xffffffffu, off, o, QF_LANES);
        if (sl == 0 && i < n) {
            const double pi = p[i];
            s += pi * (a[b] * pi + 2.0 * off);  // first entry of every row is the diagonal
        }
    }
    double r = cta_sum256(s, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = r;
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double v = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 256) v += __ldcg(partial + k);
    __syncthreads();
    double tot = cta_sum256(v, sh);
    if (threadIdx.x == 0) {
        sc[SC_PHP] = tot;
        sc[SC_ALPHA] = fmax(0.1, fmin(1.0, -sc[SC_PG] / tot));
        *counter = 0u;
    }
}

__global__ void k_axpy_dev(long long n, double* __restrict__ out, const double* __restrict__ x0, const double* __restrict__ p,
                           const double* __restrict__ alpha_dev, double alpha_host, const int* __restrict__ go) {
    if (go && *go == 0) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double alpha = alpha_dev ? *alpha_dev : alpha_host;
    if (i < n) out[i] = x0[i] + alpha * p[i];
}

__global__ void __launch_bounds__(MD_TPB) k_pair_dots(long long n, const double* __restrict__ p, const double* __restrict__ gn,
                                                      const double* __restrict__ go, double* __restrict__ Sn, double* __restrict__ Yn, int sl,
                                                      const double* __restrict__ alpha_dev, double alpha_host, HistList H,
                                                      double* __restrict__ partial, unsigned* __restrict__ counter, double* __restrict__ sc,
                                                      PeerSrc src, double* __restrict__ gn_out, int with_energy) {
    // the multi-GPU twin of k_grad_vertex_pair: forms the new pair and takes every inner product the next iteration needs: |g|^2,
    // y.s, s.g_new, and per history pair s_h.y, s.y_h, s_h.g_new.  The gradient is either already reduced over the ranks (gn) or
    // still lies as one slot per rank in peer-written memory (src.world > 0): then this kernel IS the second half of the all-reduce
    // - it waits for the ranks' flags, adds the slots in rank order, stores the reduced gradient to gn_out and the energy to sc[SC_E]
    constexpr int NACC = 3 + 3 * LB_MAXH;
    if (H.go && *H.go == 0) return;
    if (src.world > 0) {
        peer_wait_flags(src);
        if (with_energy && blockIdx.x == 0 && threadIdx.x == 0) sc[SC_E] = peer_sum(src, n);  // [g ; E]: the energy partials ride at index n
    }
    __shared__ double shm[8 * NACC], res[NACC];
    __shared__ bool last;
    const double alpha = alpha_dev ? *alpha_dev : alpha_host;
    double acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
    for (long long i = (long long)blockIdx.x * MD_TPB + threadIdx.x; i < n; i += (long long)gridDim.x * MD_TPB) {
        double gnew;
        if (src.world > 0) {
            gnew = peer_sum(src, i);
            gn_out[i] = gnew;
        } else {
            gnew = gn[i];
        }
        const double s = alpha * p[i], y = gnew - go[i];
        if (Sn) {
            Sn[i] = s;
            Yn[i] = y;
        }
        acc[0] += gnew * gnew;
        acc[1] += y * s;
        acc[2] += s * gnew;
#pragma unroll
        for (int j = 0; j < LB_MAXH; ++j)
            if (j < H.n && Sn) {
                const double si = H.S[j][i];
                acc[3 + 3 * j] += si * y;
                acc[4 + 3 * j] += s * H.Y[j][i];
                acc[5 + 3 * j] += si * gnew;
            }
    }
    const int nacc = Sn ? 3 + 3 * H.n : 3;
    if (!multi_reduce_256<NACC>(acc, nacc, shm, res, &last, partial, counter)) return;
    if ((int)threadIdx.x < nacc) {
        const int j = threadIdx.x;
        const double tot = res[j];
        if (j == 0) sc[SC_GG] = tot;
        else if (j == 1) { sc[SC_YS_NEW] = tot; if (sl >= 0) sc[SC_SY + 8 * sl + sl] = tot; }
        else if (j == 2) { if (sl >= 0) sc[SC_SG + sl] = tot; }
        else {
            const int h = (j - 3) / 3, kind = (j - 3) % 3, sh_ = H.slot[h];
            if (kind == 0) sc[SC_SY + 8 * sh_ + sl] = tot;       // s_h . y_new
            else if (kind == 1) sc[SC_SY + 8 * sl + sh_] = tot;  // s_new . y_h
            else sc[SC_SG + sh_] = tot;                          // s_h . g_new
        }
    }
}

// multi-GPU: p holds the all-reduced sum of the subdomain solutions; divide the interface entries by their duplication count
// (DOTTimeStepper.cpp:447-449) and take the inner products p . P.a[j] in the same pass
// (src.world > 0: the sum is still one slot per rank in peer-written memory - wait for the flags and add the slots in rank order here)
__global__ void __launch_bounds__(MD_TPB) k_divdup_dots(int ndof, const int* __restrict__ dup, double* __restrict__ p, DotPairs P,
                                                        double* __restrict__ partial, unsigned* __restrict__ counter, double* __restrict__ sc,
                                                        PeerSrc src) {
    if (P.go && *P.go == 0) return;
    if (src.world > 0) peer_wait_flags(src);
    __shared__ double shm[8 * 12], res[12];
    __shared__ bool last;
    double acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.0;
    for (int d = blockIdx.x * MD_TPB + threadIdx.x; d < ndof; d += gridDim.x * MD_TPB) {
        double v = src.world > 0 ? peer_sum(src, d) : p[d];
        const int du = dup[d / 3];
        if (du > 1) v /= (double)du;
        if (du > 1 || src.world > 0) p[d] = v;
#pragma unroll
        for (int j = 0; j < 12; ++j)
            if (j < P.n) acc[j] += P.a[j][d] * v;
    }
    if (!multi_reduce_256<12>(acc, P.n, shm, res, &last, partial, counter)) return;
    if ((int)threadIdx.x < P.n) sc[P.out[threadIdx.x]] = res[threadIdx.x];
}

// p = D^-1 sum of the subdomain copies (DOTTimeStepper.cpp:434-450) and, in the same pass, the inner products of p with the
// vectors of P (second multi-dot of the iteration): one launch instead of two.
__global__ void __launch_bounds__(MD_TPB) k_scatter_dots(int ndof, const int* __restrict__ cptr, const int* __restrict__ cidx,
                                                         const double* __restrict__ xs, const int* __restrict__ dup, double* __restrict__ p,
                                                         DotPairs P, double* __restrict__ partial, unsigned* __restrict__ counter,
                                                         double* __restrict__ sc) {
    if (P.go && *P.go == 0) return;
    __shared__ double shm[8 * 12], res[12];
    __shared__ bool last;
    double acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.0;
    for (int d = blockIdx.x * MD_TPB + threadIdx.x; d < ndof; d += gridDim.x * MD_TPB) {
        double v = 0.0;
        for (int k = cptr[d]; k < cptr[d + 1]; ++k) v += xs[cidx[k]];
        const int du = dup[d / 3];
        if (du > 1) v /= (double)du;
        p[d] = v;
#pragma unroll
        for (int j = 0; j < 12; ++j)
            if (j < P.n) acc[j] += P.a[j][d] * v;
    }
    if (!multi_reduce_256<12>(acc, P.n, shm, res, &last, partial, counter)) return;
    if ((int)threadIdx.x < P.n) sc[P.out[threadIdx.x]] = res[threadIdx.x];
}

// multi-GPU: this rank's sum of its subdomain copies goes straight into its slot in every rank's peer buffer (first half of the
// all-reduce of the search direction, peer_reduce.cu); the last CTA publishes the epoch
__global__ void __launch_bounds__(256) k_scatter_push(int ndof, const int* __restrict__ cptr, const int* __restrict__ cidx,
                                                      const double* __restrict__ xs, PeerDst D) {
    for (int d = blockIdx.x * 256 + threadIdx.x; d < ndof; d += gridDim.x * 256) {
        double v = 0.0;
        for (int k = cptr[d]; k < cptr[d + 1]; ++k) v += xs[cidx[k]];
        peer_store(D, d, v);
    }
    peer_publish(D);
}

#define EW_LAUNCH(kernel, n, ...)                                              \
    do {                                                                       \
        if ((n) > 0) {                                                         \
            kernel<<<ceil_div((n), 256), 256, 0, st>>>(__VA_ARGS__);           \
            count_launch();                                                    \
        }                                                                      \
    } while (0)

void launch_axpy(long long n, double* out, const double* x0, const double* p, double alpha, cudaStream_t st) {
    EW_LAUNCH(k_axpy, n, n, out, x0, p, alpha);
}
void launch_warm_start(int nV, double* x, const double* vel, const unsigned char* fixed, double dt, double gx, double gy, double gz,
                       cudaStream_t st) {
    EW_LAUNCH(k_warm_start, nV, nV, x, vel, fixed, dt, gx, gy, gz);
}
void launch_xtilde(int nV, double* xt, const double* xn, const double* vel, const unsigned char* fixed, double dt, double gx, double gy,
                   double gz, cudaStream_t st) {
    EW_LAUNCH(k_xtilde, nV, nV, xt, xn, vel, fixed, dt, gx, gy, gz);
}
void launch_velocity(int nV, double* vel, const double* x, const double* xn, double dt, cudaStream_t st) {
    long long n = 3LL * nV;
    EW_LAUNCH(k_velocity, n, n, vel, x, xn, dt);
}
int multidot_partial_count() { return md_max_blocks() * (4 + 3 * LB_MAXH); }
int multidot_blocks(long long n) { return (int)std::min<long long>(md_max_blocks(), std::max<long long>(1, (n + MD_TPB - 1) / MD_TPB)); }
static int md_blocks(long long n) { return multidot_blocks(n); }

void launch_dots(long long n, const DotPairs& P, double* partial, unsigned* counter, double* sc, cudaStream_t st) {
    if (P.n <= 0) return;
    k_dots<<<md_blocks(n), MD_TPB, 0, st>>>(n, P, partial, counter, sc);
    count_launch();
}
void launch_lbfgs_q(long long n, double* q, const double* g, const HistList& H, double* sc, cudaStream_t st) {
    EW_LAUNCH(k_lbfgs_q, n, n, q, g, H, sc);
}
void launch_lbfgs_p(long long n, double* p, const HistList& H, double* sc, cudaStream_t st) { EW_LAUNCH(k_lbfgs_p, n, n, p, H, sc); }
void launch_quadform_alpha(int n, const int* ia, const int* ja, const double* a, const double* p, double* partial, unsigned* counter,
                           double* sc, cudaStream_t st, const int* go) {
    // big matrices: 8 passes per CTA (few partials for the last CTA to add); small ones are latency-bound: one pass, more CTAs
    const int npass = n >= (1 << 18) ? QF_LANES : 1;
    k_quadform_alpha<<<ceil_div(n, npass * QF_ROWS), 256, 0, st>>>(n, ia, ja, a, p, partial, counter, sc, go, npass);
    count_launch();
}
void launch_axpy_dev(long long n, double* out, const double* x0, const double* p, const double* alpha_dev, double alpha_host,
                     cudaStream_t st, const int* go) {
    EW_LAUNCH(k_axpy_dev, n, n, out, x0, p, alpha_dev, alpha_host, go);
}
void launch_pair_dots(long long n, const double* p, const double* g_new, const double* g_old, double* S_new, double* Y_new, int sl,
                      const double* alpha_dev, double alpha_host, const HistList& H, double* partial, unsigned* counter, double* sc,
                      cudaStream_t st, const PeerSrc* src, double* g_new_out, bool with_energy) {
    PeerSrc S;
    std::memset(&S, 0, sizeof(S));
    if (src) S = *src;
    k_pair_dots<<<md_blocks(n), MD_TPB, 0, st>>>(n, p, g_new, g_old, S_new, Y_new, sl, alpha_dev, alpha_host, H, partial, counter, sc, S,
                                                 g_new_out, with_energy ? 1 : 0);
    count_launch();
}

void launch_scatter_avg_dots(int ndof, const int* cptr, const int* cidx, const double* xs, const int* dup, double* p, const DotPairs& P,
                             double* partial, unsigned* counter, double* sc, cudaStream_t st) {
    k_scatter_dots<<<md_blocks(ndof), MD_TPB, 0, st>>>(ndof, cptr, cidx, xs, dup, p, P, partial, counter, sc);
    count_launch();
}

void launch_divdup_dots(int ndof, const int* dup, double* p, const DotPairs& P, double* partial, unsigned* counter, double* sc, cudaStream_t st,
                        const PeerSrc* src) {
    PeerSrc S;
    std::memset(&S, 0, sizeof(S));
    if (src) S = *src;
    k_divdup_dots<<<md_blocks(ndof), MD_TPB, 0, st>>>(ndof, dup, p, P, partial, counter, sc, S);
    count_launch();
}

void launch_scatter_push(int ndof, const int* cptr, const int* cidx, const double* xs, const PeerDst& D, cudaStream_t st) {
    k_scatter_push<<<std::min(592, ceil_div(ndof, 256)), 256, 0, st>>>(ndof, cptr, cidx, xs, D);
    count_launch();
}

void launch_scatter_avg(int ndof, const int* cptr, const int* cidx, const double* xs, const int* dup, double* p, cudaStream_t st) {
    EW_LAUNCH(k_scatter_avg, ndof, ndof, cptr, cidx, xs, dup, p);
}

}  // namespace dotgpu
