// Numeric multifrontal Cholesky + solves, batched over many matrices, level-scheduled.
//
// Storage per supernode s (front size m, ns pivot columns, nb = m - ns):
//   panel  L[panel + r*ns + c]     r in [0,m), c in [0,ns)   row-major; rows < ns hold the lower-triangular
//                                   diagonal block L_ss, rows >= ns hold L_below
//   cb     CB[cb + i*nb + j]       contribution block (lower triangle used), nb x nb row-major
//   tinv   tinv[tinv + kb*64*64..] inverse of every 64x64 diagonal tile of L_ss (row-major, upper part zero)
// Factorisation of one level: extend-add children CBs -> for every pivot tile kb: potrf(tile) + tile
// inverse, trsm of the rows below (a GEMM with the tile inverse), rank-64 update of the whole trailing
// front (rest of the panel + CB).  GEMM tiles are 64x64 per CTA on DMMA (mma.sync m8n8k4 f64).
// After the numeric factorisation the "solve panels" S_s = [L_ss^-1 ; -L_below L_ss^-1] are built (batched over
// all supernodes), so that each level of a triangular solve is ONE launch of independent dense GEMV slabs:
// forward [y_s; u_s] = S_s (b_s + children updates) with the children gathered deterministically by the parent,
// backward x_s = S_s^T [y_s; x(rows below)].
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "chol_numeric.h"

namespace dotgpu {

namespace {

constexpr int NB = CH_NB;
constexpr int SLD = 36;  // smem leading dimension of a 64 x 32 operand chunk (conflict-free DMMA fragment loads)
constexpr int KC = 32;

struct Task2 { int s, a; };
struct Task3 { int s; short a, b; };

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// sX[r][k] = (r < nrows && k < kw) ? src[(r)*ld + k] : 0   for r < 64, k < KC;  128 threads
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

// acc += A_chunk * B_chunk^T ; warp (wr,wc) owns the 32x32 sub-tile
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

__global__ void k_scatter_a(long long nnz, const long long* __restrict__ amap, const double* __restrict__ a, double* __restrict__ L) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) L[amap[i]] = a[i];
}

// ---- extend-add: parent-driven, one CTA per (parent, 64-row slab, 128-column slab of the parent front), children processed
// in order (deterministic).  The 2-D split keeps the top levels, where a handful of parents receive large contribution blocks,
// spread over enough CTAs. ----
// column slabs of 128 only for fronts above `split_m` rows; the host passes INT_MAX for levels that have enough parents to fill the
// GPU anyway (splitting then only adds overhead)
constexpr int EA_COLS = 128, EA_SPLIT_M = 512, EA_FEW_PARENTS = 96;
__host__ __device__ inline int ea_col_width(int m, int split_m) { return m > split_m ? EA_COLS : m; }
__global__ void __launch_bounds__(256) k_extend_add(const Task3* __restrict__ tasks, const SNDesc* __restrict__ sn,
                                                    const int* __restrict__ rel, const int* __restrict__ child,
                                                    double* __restrict__ L, double* __restrict__ CB, int split_m, int phase) {
    // phase 0: the targets inside the parent's PANEL (columns < ns), before the parent's pivot chain; phase 1: the targets inside
    // the parent's contribution block, AFTER k_update_cb has stored -L_below L_below^T there.  The contribution blocks therefore
    // need no zero-fill and k_update_cb no read (one memset + one read pass over all contribution blocks per factorisation less).
    const Task3 tk = tasks[blockIdx.x];
    const SNDesc p = sn[tk.s];
    const int lo = tk.a * 64, hi = min(lo + 64, p.m);
    const int cw = ea_col_width(p.m, split_m);
    const int clo = tk.b * cw, chi = min(clo + cw, p.m);
    const int nbp = p.m - p.ns;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int s_rng[5];
    for (int ci = p.child_begin; ci < p.child_end; ++ci) {
        const SNDesc c = sn[child[ci]];
        const int nbc = c.m - c.ns;
        const int* __restrict__ crel = rel + c.rows + c.ns;  // [nbc], ascending
        if (threadIdx.x < 5) {  // first index with crel[i] >= {lo, hi, clo, chi, ns}
            const int key = threadIdx.x == 0 ? lo : (threadIdx.x == 1 ? hi : (threadIdx.x == 2 ? clo : (threadIdx.x == 3 ? chi : p.ns)));
            int a = 0, b = nbc;
            while (a < b) { int mid = (a + b) >> 1; if (crel[mid] < key) a = mid + 1; else b = mid; }
            s_rng[threadIdx.x] = a;
        }
        __syncthreads();
        const int i0 = s_rng[0], i1 = s_rng[1], js = s_rng[4];
        // columns of the child's block are ascending in the parent's numbering: [.., js) land in the panel, [js, ..) in the CB
        const int j0 = phase == 0 ? s_rng[2] : max(s_rng[2], js), j1 = phase == 0 ? min(s_rng[3], js) : s_rng[3];
        const double* __restrict__ ccb = CB + c.cb;
        for (int i = i0 + warp; i < i1; i += 8) {
            const int r = crel[i];
            double* __restrict__ prow_l = L + p.panel + (long long)r * p.ns;
            double* __restrict__ prow_c = CB + p.cb + (long long)(r - p.ns) * nbp - p.ns;
            const double* __restrict__ crow = ccb + (long long)i * nbc;
            const int jend = min(j1, i + 1);  // lower triangle of the child's contribution block
            // four independent read-modify-writes in flight per lane (targets of one child row are distinct)
            for (int jb = j0 + lane; jb < jend; jb += 128) {
                double* ptr[4];
                double v[4], o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = jb + 32 * q;
                    ptr[q] = nullptr;
                    v[q] = 0.0;
                    if (j < jend) {
                        const int cc = crel[j];
                        ptr[q] = cc < p.ns ? prow_l + cc : prow_c + cc;
                        v[q] = crow[j];
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = ptr[q] ? *ptr[q] : 0.0;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (ptr[q]) *ptr[q] = o[q] + v[q];
            }
        }
        __syncthreads();
    }
}

// ---- potrf of pivot tile kb (+ its inverse), one CTA per supernode ----
// C[n x n] (ld 65 or 33) = sign * A * B for n in {16,32}, operands in shared memory, 256 threads
template <int N, int LDC>
__device__ __forceinline__ void smem_gemm(double* C, const double* A, int lda, const double* B, int ldb, double sign) {
    for (int e = threadIdx.x; e < N * N; e += 256) {
        const int i = e / N, j = e % N;
        double sum = 0.0;
#pragma unroll 8
        for (int k = 0; k < N; ++k) sum += A[i * lda + k] * B[k * ldb + j];
        C[i * LDC + j] = sign * sum;
    }
}

constexpr int POTRF_SMEM = (2 * 64 * 65 + 32 * 33) * (int)sizeof(double);

// potrf of pivot tile kb (+ its inverse), one CTA per supernode.  Left-looking column sweep with ONE CTA barrier per
// column: 4 threads share a row, every row group also recomputes the pivot (row j . row j) itself, so nobody waits for a
// "pivot thread"; column j is final as soon as its own dot products are subtracted and scaled.
__global__ void __launch_bounds__(256) k_potrf(const int* __restrict__ tasks, int kb, const SNDesc* __restrict__ sn,
                                               double* __restrict__ L, double* __restrict__ tinv, int* __restrict__ status, int dbg) {
    extern __shared__ double smem[];
    double* T = smem;                  // [64][65] tile, then its Cholesky factor
    double* X = smem + 64 * 65;        // [64][65] inverse of the factor
    double* W = smem + 2 * 64 * 65;    // [32][33] scratch
    const int s = tasks[blockIdx.x];
    const SNDesc d = sn[s];
    const int c0 = kb * NB;
    const int w = min(NB, d.ns - c0);
    double* __restrict__ tile = L + d.panel + (long long)c0 * d.ns + c0;
    {
        double v[16];  // all 16 loads of a thread in flight together
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = threadIdx.x + 256 * i, r = e >> 6, c = e & 63;
            v[i] = (r < w && c <= r) ? tile[(long long)r * d.ns + c] : ((r >= w && c == r) ? 1.0 : 0.0);  // identity padding
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int e = threadIdx.x + 256 * i, r = e >> 6, c = e & 63;
            T[r * 65 + c] = v[i];
            X[r * 65 + c] = 0.0;
        }
    }
    __syncthreads();
    if (!(dbg & 1)) {
        // Left-looking sweep in block columns of 8 (3 CTA barriers per block column, 24 in all, instead of one per column):
        //   A  every row (4 lanes per row) subtracts its dot products with the 8 pivot rows over the finished columns k < J0
        //   B  every thread factors the updated 8x8 diagonal block redundantly in registers (no communication)
        //   C  and solves its own row against it: x = u L_D^-T (for a row of the diagonal block this reproduces L_D's row)
        const int row = threadIdx.x >> 2, q = threadIdx.x & 3;
        double* __restrict__ Ti = T + row * 65;
        for (int J0 = 0; J0 < w; J0 += 8) {
            double acc[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = 0.0;
            for (int k = q; k < J0; k += 4) {
                const double a = Ti[k];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] += a * T[(J0 + c) * 65 + k];
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
                acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
            }
            if (q == 0 && row >= J0) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (J0 + c <= row) Ti[J0 + c] -= acc[c];
            }
            __syncthreads();
            double D[8][8], u[8], x[8], inv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) u[c] = Ti[J0 + c];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) D[r][c] = T[(J0 + r) * 65 + J0 + c];
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double dd = D[c][c];
#pragma unroll
                for (int k = 0; k < c; ++k) dd -= D[c][k] * D[c][k];
                if (!(dd > 0.0)) {
                    bad = bad || (J0 + c < w);
                    dd = 1.0;
                }
                const double rs = rsqrt(dd);
                inv[c] = rs;
                D[c][c] = dd * rs;
#pragma unroll
                for (int r = c + 1; r < 8; ++r) {
                    double v = D[r][c];
#pragma unroll
                    for (int k = 0; k < c; ++k) v -= D[r][k] * D[c][k];
                    D[r][c] = v * rs;
                }
            }
            if (bad && threadIdx.x == 0) atomicCAS(status, 0, s + 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double v = u[c];
#pragma unroll
                for (int k = 0; k < c; ++k) v -= x[k] * D[c][k];
                x[c] = v * inv[c];
            }
            __syncthreads();  // everybody holds the diagonal block in registers before its rows are overwritten
            if (q == 0 && row >= J0) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (J0 + c <= row) Ti[J0 + c] = x[c];
            }
            __syncthreads();
        }
    }
    // blocked inverse: four 16x16 diagonal blocks by forward substitution (one thread per column) ...
    if (threadIdx.x < 64 && !(dbg & 2)) {
        const int c = threadIdx.x, b0 = c & ~15;
        X[c * 65 + c] = 1.0 / T[c * 65 + c];
        for (int i = c + 1; i < b0 + 16; ++i) {
            double sum = 0.0;
            for (int k = c; k < i; ++k) sum += T[i * 65 + k] * X[k * 65 + c];
            X[i * 65 + c] = -sum / T[i * 65 + i];
        }
    }
    __syncthreads();
    // ... then inv([A 0; B C]) = [A^-1 0; -C^-1 B A^-1, C^-1] at 32 and at 64
    if (!(dbg & 4)) {
    // (blocks that lie entirely in the identity padding, rows >= w, are skipped: their off-diagonal inverse blocks are zero)
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int o = 32 * p;
        if (o + 16 < w) {
            smem_gemm<16, 33>(W, T + (o + 16) * 65 + o, 65, X + o * 65 + o, 65, 1.0);
            __syncthreads();
            smem_gemm<16, 65>(X + (o + 16) * 65 + o, X + (o + 16) * 65 + (o + 16), 65, W, 33, -1.0);
            __syncthreads();
        }
    }
    if (32 < w) {
        smem_gemm<32, 33>(W, T + 32 * 65, 65, X, 65, 1.0);
        __syncthreads();
        smem_gemm<32, 65>(X + 32 * 65, X + 32 * 65 + 32, 65, W, 33, -1.0);
        __syncthreads();
    }
    }
    double* __restrict__ tinvp = tinv + d.tinv + (long long)kb * NB * NB;
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        int r = e >> 6, c = e & 63;
        const bool in = (r < w && c <= r);
        if (in) tile[(long long)r * d.ns + c] = T[r * 65 + c];
        tinvp[e] = in ? X[r * 65 + c] : 0.0;
    }
}

// ---- trsm: rows below pivot tile kb get multiplied by the tile inverse (transposed) ----
__global__ void __launch_bounds__(128) k_trsm(const Task2* __restrict__ tasks, int kb, const SNDesc* __restrict__ sn,
                                              double* __restrict__ L, const double* __restrict__ tinv) {
    __shared__ double sA[64 * SLD], sB[64 * SLD];
    const Task2 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int c0 = kb * NB;
    const int w = min(NB, d.ns - c0);
    const int r0 = c0 + w + tk.a * 64;  // first row of this slab (w < NB only for the last pivot tile, where c0 + w == ns)
    const int nrows = min(64, d.m - r0);
    double* __restrict__ A = L + d.panel + (long long)r0 * d.ns + c0;
    const double* __restrict__ B = tinv + d.tinv + (long long)kb * NB * NB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp >> 1, wc = warp & 1;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < w; k0 += KC) {
        __syncthreads();
        load_chunk(sA, A + k0, d.ns, nrows, w - k0);
        load_chunk(sB, B + k0, NB, w, w - k0);
        __syncthreads();
        if (wr * 32 < nrows && wc * 32 < w) mma_chunk(acc, sA, sB, wr, wc, lane);  // warps whose 32x32 block is all padding skip the DMMAs
    }
    __syncthreads();  // all reads of A done before anybody overwrites it
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int r = wr * 32 + i * 8 + g, c = wc * 32 + j * 8 + 2 * q + e;
                if (r < nrows && c < w) A[(long long)r * d.ns + c] = acc[i][j][e];
            }
}

// ---- rank-w update of the trailing front with pivot block column kb ----
__global__ void __launch_bounds__(128) k_update(const Task3* __restrict__ tasks, int kb, const SNDesc* __restrict__ sn,
                                                double* __restrict__ L, double* __restrict__ CB) {
    __shared__ double sA[64 * SLD], sB[64 * SLD];
    const Task3 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int c0 = kb * NB;
    const int w = min(NB, d.ns - c0);
    const int base = c0 + w;
    const int r0 = base + tk.a * 64, q0 = base + tk.b * 64;  // front-space origin of the C tile (rows r0.., cols q0..)
    const int nrows = min(64, d.m - r0), ncols = min(64, d.m - q0);
    const double* __restrict__ A = L + d.panel + (long long)r0 * d.ns + c0;
    const double* __restrict__ B = L + d.panel + (long long)q0 * d.ns + c0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp >> 1, wc = warp & 1;
    // warps whose 32x32 block is all padding, lies strictly above the diagonal, or right of the panel columns skip the DMMAs
    const bool warp_active = wr * 32 < nrows && wc * 32 < ncols && r0 + wr * 32 + 31 >= q0 + wc * 32 && q0 + wc * 32 < d.ns;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < w; k0 += KC) {
        __syncthreads();
        load_chunk(sA, A + k0, d.ns, nrows, w - k0);
        load_chunk(sB, B + k0, d.ns, ncols, w - k0);
        __syncthreads();
        if (warp_active) mma_chunk(acc, sA, sB, wr, wc, lane);
    }
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int rr = wr * 32 + i * 8 + g, cc = wc * 32 + j * 8 + 2 * q + e;
                if (rr < nrows && cc < ncols) {
                    int r = r0 + rr, c = q0 + cc;
                    if (r >= c && c < d.ns) L[d.panel + (long long)r * d.ns + c] -= acc[i][j][e];  // CB: see k_update_cb
                }
            }
}

// ---- contribution block: CB = -L_below L_below^T in ONE pass over all pivot columns of the supernode (K = ns), after the
// panel is final.  Every CB tile is accumulated in registers and written once (the per-pivot-step updates above only touch
// the panel columns), which gives the DMMA tiles a long K loop instead of 64-wide read-modify-write passes. ----
__global__ void __launch_bounds__(128) k_update_cb(const Task3* __restrict__ tasks, const SNDesc* __restrict__ sn,
                                                   const double* __restrict__ L, double* __restrict__ CB) {
    __shared__ double sA[64 * SLD], sB[64 * SLD];
    const Task3 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int r0 = d.ns + tk.a * 64, q0 = d.ns + tk.b * 64;
    const int nrows = min(64, d.m - r0), ncols = min(64, d.m - q0);
    const double* __restrict__ A = L + d.panel + (long long)r0 * d.ns;
    const double* __restrict__ B = L + d.panel + (long long)q0 * d.ns;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp >> 1, wc = warp & 1;
    // warps whose 32x32 block is all padding or lies strictly above the diagonal skip the DMMAs (leaf-side fronts have contribution
    // blocks of ~90 rows: 6 of the 12 warp blocks of their 3 tiles hold anything that is stored)
    const bool warp_active = wr * 32 < nrows && wc * 32 < ncols && r0 + wr * 32 + 31 >= q0 + wc * 32;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < d.ns; k0 += KC) {
        __syncthreads();
        load_chunk(sA, A + k0, d.ns, nrows, d.ns - k0);
        load_chunk(sB, B + k0, d.ns, ncols, d.ns - k0);
        __syncthreads();
        if (warp_active) mma_chunk(acc, sA, sB, wr, wc, lane);
    }
    const int g = lane >> 2, q = lane & 3;
    const int nb = d.m - d.ns;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int rr = wr * 32 + i * 8 + g, cc = wc * 32 + j * 8 + 2 * q + e;
                if (rr < nrows && cc < ncols) {
                    int r = r0 + rr, c = q0 + cc;
                    if (r >= c) CB[d.cb + (long long)(r - d.ns) * nb + (c - d.ns)] = -acc[i][j][e];  // the children's blocks are added afterwards
                }
            }
}

// ---- solve panels:  S_s = [ L_ss^-1 ; -L_below L_ss^-1 ]  (m x ns, row-major, same offsets as the L panels) ----
// With S every phase of a triangular solve is a dense GEMV that any number of CTAs can share: there is no serial
// substitution chain left inside a supernode ("hard parts": sparse triangular solves).
constexpr int TLD = 72;  // smem leading dimension of a [KC][64] operand chunk given as B[k][n]

// sB[k][n] = (k < kw && n < ncols) ? src[k*ld + n] : 0   for k < KC, n < 64;  128 threads
__device__ __forceinline__ void load_chunk_kn(double* sB, const double* __restrict__ src, long long ld, int kw, int ncols) {
    const int n = threadIdx.x & 63, k0 = threadIdx.x >> 6;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int k = k0 + 2 * i;
        double v = 0.0;
        if (k < kw && n < ncols) v = src[(long long)k * ld + n];
        sB[k * TLD + n] = v;
    }
}

// acc += A_chunk[64 x KC] * B_chunk[KC x 64] with B stored [k][n]
__device__ __forceinline__ void mma_chunk_kn(double (&acc)[4][4][2], const double* sA, const double* sB, int wr, int wc, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int k0 = 0; k0 < KC; k0 += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sA[(wr * 32 + i * 8 + g) * SLD + k0 + q];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sB[(k0 + q) * TLD + wc * 32 + j * 8 + g];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// diagonal tiles of L_ss^-1 are the tile inverses
__global__ void __launch_bounds__(256) k_sp_diag(const Task2* __restrict__ tasks, const SNDesc* __restrict__ sn,
                                                 const double* __restrict__ tinv, double* __restrict__ Sp) {
    const Task2 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int c0 = tk.a * NB, w = min(NB, d.ns - c0);
    const double* __restrict__ D = tinv + d.tinv + (long long)tk.a * NB * NB;
    for (int e = threadIdx.x; e < 64 * 64; e += 256) {
        int r = e >> 6, c = e & 63;
        if (r < w && c < w) Sp[d.panel + (long long)(c0 + r) * d.ns + c0 + c] = D[e];
    }
}

// block row i of X = L_ss^-1:  X_ij = -D_i * sum_{k=j}^{i-1} L_ik X_kj   (j < i); one CTA per (s, j)
__global__ void __launch_bounds__(128) k_sp_triinv(const Task2* __restrict__ tasks, int i, const SNDesc* __restrict__ sn,
                                                   const double* __restrict__ L, const double* __restrict__ tinv,
                                                   double* __restrict__ Sp) {
    __shared__ double sA[64 * SLD], sB[KC * TLD];
    const Task2 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int j = tk.a;
    const int r0 = i * NB, c0 = j * NB;
    const int wi = min(NB, d.ns - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp >> 1, wc = warp & 1;
    const int g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    // acc = sum_k L_ik X_kj, K runs over columns c0 .. r0-1 of block row i
    const double* __restrict__ A = L + d.panel + (long long)r0 * d.ns;
    const double* __restrict__ X = Sp + d.panel;
    for (int k0 = c0; k0 < r0; k0 += KC) {
        __syncthreads();
        load_chunk(sA, A + k0, d.ns, wi, r0 - k0);
        load_chunk_kn(sB, X + (long long)k0 * d.ns + c0, d.ns, r0 - k0, NB);
        __syncthreads();
        mma_chunk_kn(acc, sA, sB, wr, wc, lane);
    }
    // out = -D_i * acc : feed acc back through shared memory as the [k][n] operand, 32 rows at a time
    double out[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) out[a][b][0] = out[a][b][1] = 0.0;
    const double* __restrict__ D = tinv + d.tinv + (long long)i * NB * NB;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        __syncthreads();
        if (wr == half) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int e = 0; e < 2; ++e) sB[(a * 8 + g) * TLD + wc * 32 + b * 8 + 2 * q + e] = acc[a][b][e];
        }
        load_chunk(sA, D + half * KC, NB, NB, KC);
        __syncthreads();
        mma_chunk_kn(out, sA, sB, wr, wc, lane);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int r = wr * 32 + a * 8 + g, c = wc * 32 + b * 8 + 2 * q + e;
                if (r < wi) Sp[d.panel + (long long)(r0 + r) * d.ns + c0 + c] = -out[a][b][e];
            }
}

// below part: S[ns+r][c] = -sum_{k >= c-tile} L_below[r][k] X[k][c]; one CTA per (s, 64-row slab, column tile)
__global__ void __launch_bounds__(128) k_sp_below(const Task3* __restrict__ tasks, const SNDesc* __restrict__ sn,
                                                  const double* __restrict__ L, double* __restrict__ Sp) {
    __shared__ double sA[64 * SLD], sB[KC * TLD];
    const Task3 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int r0 = d.ns + tk.a * 64, c0 = tk.b * NB;
    const int nrows = min(64, d.m - r0), ncols = min(NB, d.ns - c0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp >> 1, wc = warp & 1;
    const int g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const double* __restrict__ A = L + d.panel + (long long)r0 * d.ns;
    const double* __restrict__ X = Sp + d.panel;
    for (int k0 = c0; k0 < d.ns; k0 += KC) {
        __syncthreads();
        load_chunk(sA, A + k0, d.ns, nrows, d.ns - k0);
        load_chunk_kn(sB, X + (long long)k0 * d.ns + c0, d.ns, d.ns - k0, ncols);
        __syncthreads();
        mma_chunk_kn(acc, sA, sB, wr, wc, lane);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int r = wr * 32 + a * 8 + g, c = wc * 32 + b * 8 + 2 * q + e;
                if (r < nrows && c < ncols) Sp[d.panel + (long long)(r0 + r) * d.ns + c0 + c] = -acc[a][b][e];
            }
}

// ---- forward sweep of one level: [y_s ; u_s] = S_s t (+ children updates), t = b_s + children updates ----
__global__ void __launch_bounds__(256) k_fwd(const Task2* __restrict__ tasks, const SNDesc* __restrict__ sn,
                                             const long long* __restrict__ ea_ptr, const long long* __restrict__ ea_src,
                                             const double* __restrict__ Sp, const double* __restrict__ b,
                                             double* __restrict__ uwork, double* __restrict__ y) {
    extern __shared__ double smem[];
    const Task2 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = tk.a * 64, r1 = min(r0 + 64, d.m);
    // t is needed up to column min(r1, ns) only (rows of the triangular top part stop at their diagonal)
    const int tn = min(d.ns, r1);
    for (int k = threadIdx.x; k < tn; k += 256) {
        double v = b[d.col0 + k];
        for (long long e = ea_ptr[d.rows + k]; e < ea_ptr[d.rows + k + 1]; ++e) v += uwork[ea_src[e]];
        smem[k] = v;
    }
    __syncthreads();
    for (int r = r0 + warp; r < r1; r += 8) {
        const double* __restrict__ row = Sp + d.panel + (long long)r * d.ns;
        const int kn = r < d.ns ? r + 1 : d.ns;
        double sum = 0.0;
        for (int k = lane; k < kn; k += 32) sum += row[k] * smem[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            if (r < d.ns) {
                y[d.col0 + r] = sum;
            } else {
                double v = 0.0;
                for (long long e = ea_ptr[d.rows + r]; e < ea_ptr[d.rows + r + 1]; ++e) v += uwork[ea_src[e]];
                uwork[d.u + (r - d.ns)] = v + sum;
            }
        }
    }
}

// ---- backward sweep of one level: x_s = S_s^T [y_s ; x(rows below)]  (64-column slab per CTA) ----
__global__ void __launch_bounds__(256) k_bwd(const Task2* __restrict__ tasks, const SNDesc* __restrict__ sn,
                                             const int* __restrict__ rows, const double* __restrict__ Sp,
                                             const double* __restrict__ y, double* __restrict__ x) {
    extern __shared__ double smem[];  // z[m]
    __shared__ double red[4][64];
    const Task2 tk = tasks[blockIdx.x];
    const SNDesc d = sn[tk.s];
    const int c0 = tk.a * 64;
    const int cl = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int c = c0 + cl;
    // rows above the slab's first column contribute nothing (upper triangle)
    for (int r = c0 + threadIdx.x; r < d.m; r += 256) smem[r] = r < d.ns ? y[d.col0 + r] : x[rows[d.rows + r]];
    __syncthreads();
    double sum = 0.0;
    if (c < d.ns) {
        const double* __restrict__ col = Sp + d.panel + c;
        // diagonal tile: rows c0 .. c0+63 need the r >= c mask
        const int dend = min(c0 + 64, d.ns);
        for (int r = c0 + grp; r < dend; r += 4)
            if (r >= c) sum += col[(long long)r * d.ns] * smem[r];
        for (int r = dend + grp; r < d.m; r += 4) sum += col[(long long)r * d.ns] * smem[r];
    }
    red[grp][cl] = sum;
    __syncthreads();
    if (grp == 0 && c < d.ns) x[d.col0 + c] = (red[0][cl] + red[1][cl]) + (red[2][cl] + red[3][cl]);
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
void CholBatch::analyze(const std::vector<const int32_t*>& ia, const std::vector<const int32_t*>& ja, const std::vector<int>& n,
                        int leaf_nodes, cudaStream_t st) {
    nmat = (int)n.size();
    sym.assign(nmat, Symbolic());
    std::string first_error;  // exceptions must not leave an OpenMP region
#pragma omp parallel for schedule(dynamic)
    for (int m = 0; m < nmat; ++m) {
        try {
            sym[m].analyze(n[m], ia[m], ja[m], leaf_nodes);
        } catch (const std::exception& e) {
#pragma omp critical
            if (first_error.empty()) first_error = e.what();
        }
    }
    if (!first_error.empty()) throw Error(DOTGPU_ERR_INVALID, first_error);
    col_off.assign(nmat + 1, 0);
    nnz_off.assign(nmat + 1, 0);
    sn_off.assign(nmat + 1, 0);
    std::vector<int64_t> panel_base(nmat + 1, 0), cb_base(nmat + 1, 0), u_base(nmat + 1, 0), rows_base(nmat + 1, 0),
        ea_base(nmat + 1, 0), child_base(nmat + 1, 0);
    nlevels = 0;
    flops_total = 0.0;
    for (int m = 0; m < nmat; ++m) {
        const Symbolic& S = sym[m];
        col_off[m + 1] = col_off[m] + S.n;
        nnz_off[m + 1] = nnz_off[m] + (int64_t)S.amap.size();
        sn_off[m + 1] = sn_off[m] + S.nsuper;
        panel_base[m + 1] = panel_base[m] + S.nnz_l;
        cb_base[m + 1] = cb_base[m] + S.cb_off[S.nsuper];
        u_base[m + 1] = u_base[m] + S.u_off[S.nsuper];
        rows_base[m + 1] = rows_base[m] + S.row_ptr[S.nsuper];
        ea_base[m + 1] = ea_base[m] + (int64_t)S.ea_src.size();
        child_base[m + 1] = child_base[m] + (int64_t)S.child_list.size();
        nlevels = std::max(nlevels, S.nlevels);
        flops_total += S.flops;
    }
    n_total = col_off[nmat];
    nnz_a_total = nnz_off[nmat];
    nnz_l_total = panel_base[nmat];
    cb_total = cb_base[nmat];
    u_total = u_base[nmat];
    nsuper_total = sn_off[nmat];
    DG_REQUIRE(n_total < (1LL << 31) && rows_base[nmat] < (1LL << 40), "batch too large");

    std::vector<SNDesc> sn(nsuper_total);
    std::vector<int> rows(rows_base[nmat]), rel(rows_base[nmat]), child(child_base[nmat]);
    std::vector<long long> amap(nnz_a_total), ea_ptr(rows_base[nmat] + 1), ea_src(ea_base[nmat]);
    tinv_total = 0;
    for (int m = 0; m < nmat; ++m) {
        const Symbolic& S = sym[m];
        for (int s = 0; s < S.nsuper; ++s) {
            SNDesc& d = sn[sn_off[m] + s];
            d.panel = panel_base[m] + S.panel_off[s];
            d.cb = cb_base[m] + S.cb_off[s];
            d.u = u_base[m] + S.u_off[s];
            d.rows = rows_base[m] + S.row_ptr[s];
            d.m = S.front(s);
            d.ns = S.nscol(s);
            d.col0 = (int)(col_off[m] + S.super_ptr[s]);
            d.child_begin = (int)(child_base[m] + S.child_ptr[s]);
            d.child_end = (int)(child_base[m] + S.child_ptr[s + 1]);
            d.parent = S.parent[s] < 0 ? -1 : sn_off[m] + S.parent[s];
            d.tinv = tinv_total;
            tinv_total += (int64_t)((d.ns + NB - 1) / NB) * NB * NB;
        }
        for (int64_t i = 0; i < S.row_ptr[S.nsuper]; ++i) {
            rows[rows_base[m] + i] = (int)(col_off[m] + S.rows[i]);
            rel[rows_base[m] + i] = S.rel[i];
            ea_ptr[rows_base[m] + i] = ea_base[m] + S.ea_ptr[i];
        }
        for (size_t i = 0; i < S.ea_src.size(); ++i) ea_src[ea_base[m] + i] = u_base[m] + S.ea_src[i];
        for (size_t i = 0; i < S.child_list.size(); ++i) child[child_base[m] + i] = sn_off[m] + S.child_list[i];
        for (size_t i = 0; i < S.amap.size(); ++i) amap[nnz_off[m] + i] = panel_base[m] + S.amap[i];
    }
    ea_ptr[rows_base[nmat]] = ea_base[nmat];

    // ---- level plans (levels merged across matrices) ----
    plan.assign(nlevels, LevelPlan());
    level_sns.clear();
    std::vector<int> tasks;
    auto push2 = [&](int s, int a) { tasks.push_back(s); tasks.push_back(a); };
    for (int lv = 0; lv < nlevels; ++lv) {
        LevelPlan& P = plan[lv];
        std::vector<int> sns;
        for (int m = 0; m < nmat; ++m) {
            const Symbolic& S = sym[m];
            if (lv >= S.nlevels) continue;
            for (int i = S.level_ptr[lv]; i < S.level_ptr[lv + 1]; ++i) sns.push_back(sn_off[m] + S.level_list[i]);
        }
        // extend-add
        P.extend.off = (int)tasks.size();
        int nparents = 0;
        for (int s : sns) nparents += sn[s].child_end > sn[s].child_begin;
        P.extend_split_m = nparents < EA_FEW_PARENTS ? EA_SPLIT_M : 0x7fffffff;
        for (int s : sns)
            if (sn[s].child_end > sn[s].child_begin)
                for (int a = 0; a * 64 < sn[s].m; ++a)
                    for (int b = 0; b * ea_col_width(sn[s].m, P.extend_split_m) < std::min(sn[s].m, a * 64 + 64); ++b)
                        push2(s, (a & 0xffff) | (b << 16));  // columns <= rows
        P.extend.cnt = ((int)tasks.size() - P.extend.off) / 2;
        int maxsteps = 0;
        for (int s : sns) maxsteps = std::max(maxsteps, (sn[s].ns + NB - 1) / NB);
        P.potrf.assign(maxsteps, Span());
        P.trsm.assign(maxsteps, Span());
        P.update.assign(maxsteps, Span());
        for (int kb = 0; kb < maxsteps; ++kb) {
            P.potrf[kb].off = (int)tasks.size();
            for (int s : sns)
                if (kb * NB < sn[s].ns) tasks.push_back(s);
            P.potrf[kb].cnt = (int)tasks.size() - P.potrf[kb].off;
            P.trsm[kb].off = (int)tasks.size();
            for (int s : sns)
                if (kb * NB < sn[s].ns)
                    for (int a = 0; std::min((kb + 1) * NB, sn[s].ns) + a * 64 < sn[s].m; ++a) push2(s, a);
            P.trsm[kb].cnt = ((int)tasks.size() - P.trsm[kb].off) / 2;
            P.update[kb].off = (int)tasks.size();
            for (int s : sns)
                if (kb * NB < sn[s].ns) {
                    const int base = std::min((kb + 1) * NB, sn[s].ns);
                    int nt = 0;
                    while (base + nt * 64 < sn[s].m) ++nt;
                    for (int a = 0; a < nt; ++a)
                        for (int b = 0; b <= a; ++b)
                            if (base + b * 64 < sn[s].ns) push2(s, (a & 0xffff) | (b << 16));  // panel columns only
                }
            P.update[kb].cnt = ((int)tasks.size() - P.update[kb].off) / 2;
        }
        P.update_cb.off = (int)tasks.size();
        for (int s : sns) {
            int nt = 0;
            while (sn[s].ns + nt * 64 < sn[s].m) ++nt;
            for (int a = 0; a < nt; ++a)
                for (int b = 0; b <= a; ++b) push2(s, (a & 0xffff) | (b << 16));
        }
        P.update_cb.cnt = ((int)tasks.size() - P.update_cb.off) / 2;
        // solve-panel construction, scheduled behind the pivot chain: after potrf(kb) the diagonal tile kb of L_ss^-1 is its tile
        // inverse and block row kb of L_ss^-1 (needs L rows <= kb, final after trsm(k < kb), and the rows < kb of the inverse)
        P.sp_diag.assign(maxsteps, Span());
        P.sp_triinv.assign(maxsteps, Span());
        for (int kb = 0; kb < maxsteps; ++kb) {
            P.sp_diag[kb].off = (int)tasks.size();
            for (int s : sns)
                if (kb * NB < sn[s].ns) push2(s, kb);
            P.sp_diag[kb].cnt = ((int)tasks.size() - P.sp_diag[kb].off) / 2;
            P.sp_triinv[kb].off = (int)tasks.size();
            for (int s : sns)
                if (kb * NB < sn[s].ns)
                    for (int j = 0; j < kb; ++j) push2(s, j);
            P.sp_triinv[kb].cnt = ((int)tasks.size() - P.sp_triinv[kb].off) / 2;
        }
        P.sp_below.off = (int)tasks.size();
        for (int s : sns)
            for (int a = 0; a * 64 < sn[s].m - sn[s].ns; ++a)
                for (int b = 0; b * NB < sn[s].ns; ++b) push2(s, (a & 0xffff) | (b << 16));
        P.sp_below.cnt = ((int)tasks.size() - P.sp_below.off) / 2;
        level_sns.push_back(sns);
        P.fwd.off = (int)tasks.size();
        for (int s : sns)
            for (int a = 0; a * 64 < sn[s].m; ++a) push2(s, a);
        P.fwd.cnt = ((int)tasks.size() - P.fwd.off) / 2;
        P.bwd.off = (int)tasks.size();
        for (int s : sns)
            for (int a = 0; a * 64 < sn[s].ns; ++a) push2(s, a);
        P.bwd.cnt = ((int)tasks.size() - P.bwd.off) / 2;
    }
    if (tasks.empty()) tasks.push_back(0);

    d_sn.upload(sn, st);
    d_rows.upload(rows, st);
    d_rel.upload(rel, st);
    if (child.empty()) child.push_back(0);
    d_child.upload(child, st);
    d_amap.upload(amap, st);
    d_ea_ptr.upload(ea_ptr, st);
    if (ea_src.empty()) ea_src.push_back(0);
    d_ea_src.upload(ea_src, st);
    d_tasks.upload(tasks, st);
    L.alloc(std::max<int64_t>(nnz_l_total, 1));
    Sp.alloc(std::max<int64_t>(nnz_l_total, 1));
    CB.alloc(std::max<int64_t>(cb_total, 1));
    tinv.alloc(std::max<int64_t>(tinv_total, 1));
    ywork.alloc(n_total);
    xwork.alloc(n_total);
    rwork.alloc(n_total);
    uwork.alloc(std::max<int64_t>(u_total, 1));
    d_status.alloc(1);
    d_status.zero(st);
    max_front_all = 0;
    for (auto& S : sym) max_front_all = std::max(max_front_all, S.max_front);
    const size_t shm_solve = (size_t)(max_front_all + 8) * sizeof(double);
    DG_REQUIRE(shm_solve <= 200 * 1024, "front too large for the solve kernels' shared memory");
    DG_CUDA(cudaFuncSetAttribute(k_potrf, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_SMEM));
    // the opt-in limit is a per-function, process-wide attribute shared by every CholBatch: always raise it to the cap
    // (a smaller batch analysed later must not lower it under a larger one; occupancy follows the launch-time size)
    DG_CUDA(cudaFuncSetAttribute(k_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DG_CUDA(cudaFuncSetAttribute(k_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (graph_exec) {
        cudaGraphExecDestroy(graph_exec);
        graph_exec = nullptr;
    }
    build_solve_plan(sn, st);
    if (!st2) DG_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    if (!ev_join) DG_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    DG_CUDA(cudaStreamSynchronize(st));
    factorized = false;
}

CholBatch::~CholBatch() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (auto& e : evs) cudaEventDestroy(e);
    if (ev_join) cudaEventDestroy(ev_join);
    if (st2) cudaStreamDestroy(st2);
}

int64_t CholBatch::device_bytes() const {
    return (int64_t)(Pf.bytes() + Pb.bytes() + Ubuf.bytes() + d_stasks.bytes() + L.bytes() + Sp.bytes() + CB.bytes() + tinv.bytes() + ywork.bytes() * 3 + uwork.bytes() + d_rows.bytes() * 2 +
                     d_amap.bytes() + d_ea_ptr.bytes() + d_ea_src.bytes() + d_tasks.bytes() + d_sn.bytes());
}

void CholBatch::enqueue_factorize(const double* a_all, cudaStream_t st) {
    // DOTGPU_FACTOR_TIMING=1: per-phase device times of this factorisation (CUDA events on the stream), printed to stderr
    static const bool timing = std::getenv("DOTGPU_FACTOR_TIMING") != nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    auto mark = [&](const char* tag) {
        if (!timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.emplace_back(tag, e);
    };
    mark("start");
    size_t ev_used = 0;
    auto next_event = [&]() {
        if (ev_used == evs.size()) {
            cudaEvent_t e;
            DG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            evs.push_back(e);
        }
        return evs[ev_used++];
    };
    {   // the second stream starts behind everything already queued on the caller's stream
        cudaEvent_t e = next_event();
        DG_CUDA(cudaEventRecord(e, st));
        DG_CUDA(cudaStreamWaitEvent(st2, e, 0));
    }
    static const int potrf_dbg = std::getenv("DOTGPU_POTRF_DBG") ? std::atoi(std::getenv("DOTGPU_POTRF_DBG")) : 0;  // experiments only
    DG_CUDA(cudaMemsetAsync(L.p, 0, L.bytes(), st));
    DG_CUDA(cudaMemsetAsync(d_status.p, 0, sizeof(int), st));
    if (nnz_a_total > 0) {
        k_scatter_a<<<ceil_div(nnz_a_total, 256), 256, 0, st>>>(nnz_a_total, d_amap.p, a_all, L.p);
        count_launch();
    }
    mark("memset+scatter");
    const int* T = d_tasks.p;
    for (int lv = 0; lv < nlevels; ++lv) {
        const LevelPlan& P = plan[lv];
        if (P.extend.cnt) {
            k_extend_add<<<P.extend.cnt, 256, 0, st>>>((const Task3*)(T + P.extend.off), d_sn.p, d_rel.p, d_child.p, L.p, CB.p,
                                                          P.extend_split_m, 0);
            count_launch();
            mark("extend_add");
        }
        for (size_t kb = 0; kb < P.potrf.size(); ++kb) {
            if (P.potrf[kb].cnt) {
                k_potrf<<<P.potrf[kb].cnt, 256, POTRF_SMEM, st>>>(T + P.potrf[kb].off, (int)kb, d_sn.p, L.p, tinv.p,
                                                                                 d_status.p, potrf_dbg);
                count_launch();
                mark("potrf");
                // second stream, one step behind the pivot chain: tiles of the solve panels that are computable now
                cudaEvent_t e = next_event();
                DG_CUDA(cudaEventRecord(e, st));
                DG_CUDA(cudaStreamWaitEvent(st2, e, 0));
                if (P.sp_diag[kb].cnt) {
                    k_sp_diag<<<P.sp_diag[kb].cnt, 256, 0, st2>>>((const Task2*)(T + P.sp_diag[kb].off), d_sn.p, tinv.p, Sp.p);
                    count_launch();
                }
                if (P.sp_triinv[kb].cnt) {
                    k_sp_triinv<<<P.sp_triinv[kb].cnt, 128, 0, st2>>>((const Task2*)(T + P.sp_triinv[kb].off), (int)kb, d_sn.p, L.p, tinv.p, Sp.p);
                    count_launch();
                }
            }
            if (P.trsm[kb].cnt) {
                k_trsm<<<P.trsm[kb].cnt, 128, 0, st>>>((const Task2*)(T + P.trsm[kb].off), (int)kb, d_sn.p, L.p, tinv.p);
                count_launch();
                mark("trsm");
            }
            if (P.update[kb].cnt) {
                k_update<<<P.update[kb].cnt, 128, 0, st>>>((const Task3*)(T + P.update[kb].off), (int)kb, d_sn.p, L.p, CB.p);
                count_launch();
                mark("update_panel");
            }
        }
        {   // the level's panels are final: rows below of its solve panels and the packing, on the second stream
            cudaEvent_t e = next_event();
            DG_CUDA(cudaEventRecord(e, st));
            DG_CUDA(cudaStreamWaitEvent(st2, e, 0));
            if (P.sp_below.cnt) {
                k_sp_below<<<P.sp_below.cnt, 128, 0, st2>>>((const Task3*)(T + P.sp_below.off), d_sn.p, L.p, Sp.p);
                count_launch();
            }
            pack_panels(lv, st2);
        }
        if (P.update_cb.cnt) {
            k_update_cb<<<P.update_cb.cnt, 128, 0, st>>>((const Task3*)(T + P.update_cb.off), d_sn.p, L.p, CB.p);
            count_launch();
            mark("update_cb");
        }
        if (P.extend.cnt && P.update_cb.cnt) {  // the children's contribution blocks into this level's contribution blocks
            k_extend_add<<<P.extend.cnt, 256, 0, st>>>((const Task3*)(T + P.extend.off), d_sn.p, d_rel.p, d_child.p, L.p, CB.p,
                                                          P.extend_split_m, 1);
            count_launch();
            mark("extend_add_cb");
        }
    }
    DG_CUDA(cudaEventRecord(ev_join, st2));
    DG_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
    mark("join_solve_panels");
    factorized = true;
    if (timing && !marks.empty()) {
        cudaStreamSynchronize(st);
        std::vector<std::pair<std::string, std::pair<double, int>>> agg;
        double total = 0.0;
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            total += ms;
            bool found = false;
            for (auto& a : agg)
                if (a.first == marks[i].first) { a.second.first += ms; a.second.second++; found = true; }
            if (!found) agg.push_back({marks[i].first, {ms, 1}});
        }
        std::fprintf(stderr, "[dotgpu factorize] total %.3f ms:", total);
        for (auto& a : agg) std::fprintf(stderr, " %s %.3f (%d)", a.first.c_str(), a.second.first, a.second.second);
        std::fprintf(stderr, "\n");
        for (auto& m : marks) cudaEventDestroy(m.second);
    }
}

// The launch sequence of a numeric factorisation is static (it depends on the symbolic analysis only), so it is captured into
// a CUDA graph once - both streams, ~170 kernel nodes on bar17K_like - and replayed every frame: the dependent launches of
// the pivot chain then follow each other without the stream-launch gaps.  DOTGPU_NO_GRAPH=1 or DOTGPU_FACTOR_TIMING=1 use
// plain stream launches.
void CholBatch::factorize(const double* a_all, cudaStream_t st) {
    static const bool no_graph = std::getenv("DOTGPU_NO_GRAPH") != nullptr || std::getenv("DOTGPU_FACTOR_TIMING") != nullptr;
    if (no_graph) {
        enqueue_factorize(a_all, st);
        return;
    }
    if (!graph_exec || graph_a != a_all || graph_st != st) {
        if (graph_exec) {
            cudaGraphExecDestroy(graph_exec);
            graph_exec = nullptr;
        }
        const int64_t before = g_launch_count;
        cudaGraph_t graph = nullptr;
        DG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try {
            enqueue_factorize(a_all, st);
        } catch (...) {
            cudaStreamEndCapture(st, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        DG_CUDA(cudaStreamEndCapture(st, &graph));
        graph_launches = g_launch_count - before;
        g_launch_count = before;
        const cudaError_t ie = cudaGraphInstantiate(&graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        DG_CUDA(ie);
        graph_a = a_all;
        graph_st = st;
    }
    DG_CUDA(cudaGraphLaunch(graph_exec, st));
    count_launch((int)graph_launches);
    factorized = true;
}

void CholBatch::check_status(cudaStream_t st) {
    int h = 0;
    DG_CUDA(cudaMemcpyAsync(&h, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    DG_CUDA(cudaGetLastError());
    if (h != 0) throw Error(DOTGPU_ERR_NOT_SPD, "matrix not positive definite (supernode " + std::to_string(h - 1) + ")");
}

void CholBatch::solve_levels(const double* b_perm, double* x_perm, cudaStream_t st) {
    if (!factorized) throw Error(DOTGPU_ERR_STATE, "solve before factorize");
    const int* T = d_tasks.p;
    const size_t shm = (size_t)(max_front_all + 8) * sizeof(double);
    for (int lv = 0; lv < nlevels; ++lv) {
        const LevelPlan& P = plan[lv];
        if (P.fwd.cnt) {
            k_fwd<<<P.fwd.cnt, 256, shm, st>>>((const Task2*)(T + P.fwd.off), d_sn.p, d_ea_ptr.p, d_ea_src.p, Sp.p, b_perm, uwork.p, ywork.p);
            count_launch();
        }
    }
    for (int lv = nlevels - 1; lv >= 0; --lv) {
        const LevelPlan& P = plan[lv];
        if (P.bwd.cnt) {
            k_bwd<<<P.bwd.cnt, 256, shm, st>>>((const Task2*)(T + P.bwd.off), d_sn.p, d_rows.p, Sp.p, ywork.p, xwork.p);
            count_launch();
        }
    }
    DG_CUDA(cudaMemcpyAsync(x_perm, xwork.p, n_total * sizeof(double), cudaMemcpyDeviceToDevice, st));
}

}  // namespace dotgpu
