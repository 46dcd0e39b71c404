// Per-tet kernels K1 (energy), K2 (gradient), K3 (elemental projected Hessians) + SVD dump.
// One thread per tet; SoA loads coalesce across the warp; vertex positions are gathered through L2
// (3*nV doubles are L2-resident for every mesh this path targets).  Reductions are deterministic
// (fixed-shape block tree + a single-CTA final pass); the vertex gather sums elemental
// contributions in ascending tet order like Energy.cpp:543-563.
#include <algorithm>

#include "device_mesh.h"
#include "linalg.h"
#include "peer.cuh"
#include "reduce.cuh"
#include "elastic.cuh"

namespace dotgpu {

thread_local int64_t g_launch_count = 0;

namespace {

constexpr int TPB = 128;
constexpr int ETPB = 256;  // K1: fewer, fatter CTAs -> fewer partials for the last block

struct TetIn {
    Mat3 F;
    double B[9];  // Dm^-1 row-major
    double vol, mu, lam;
    int v[4];
};

__device__ __forceinline__ void load_tet(int t, int nT, const int4* __restrict__ tets, const double* __restrict__ DmInv,
                                         const double* __restrict__ vol, const double* __restrict__ mu,
                                         const double* __restrict__ lam, const double* __restrict__ x, TetIn& o) {
    int4 id = tets[t];
    o.v[0] = id.x; o.v[1] = id.y; o.v[2] = id.z; o.v[3] = id.w;
#pragma unroll
    for (int j = 0; j < 9; ++j) o.B[j] = DmInv[(size_t)j * nT + t];
    o.vol = vol[t]; o.mu = mu[t]; o.lam = lam[t];
    double x0[3], d[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) x0[c] = x[3 * (size_t)id.x + c];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        d[0][c] = x[3 * (size_t)id.y + c] - x0[c];
        d[1][c] = x[3 * (size_t)id.z + c] - x0[c];
        d[2][c] = x[3 * (size_t)id.w + c] - x0[c];
    }
    // F = Ds Dm^-1, Ds columns = d[k]:  F(i,j) = sum_k d[k][i] B[k][j]
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.F(i, j) = d[0][i] * o.B[j] + d[1][i] * o.B[3 + j] + d[2][i] * o.B[6 + j];
}

// deterministic block sum (fixed tree); result valid in thread 0
template <int N>
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < N / 32; ++i) r += sh[i];
    }
    return r;
}

// K1: incremental potential in ONE launch.  Blocks [0, nbT) evaluate vol_t*Psi_t of TPB tets each, blocks [nbT, nbT+nbV)
// the inertia term of 256... (TPB) vertices each (sum_v m_v |x_v - xTilde_v|^2 / 2 over ALL vertices, Optimizer.cpp:1204-1211);
// every block stores its partial sum, the last block to finish adds them in a fixed order:
//   out = coef * sum(partial[0..nbT)) + sum(partial[nbT..nbT+nbV))          (deterministic, bit-reproducible)
template <int EN>
__global__ void __launch_bounds__(ETPB) k_energy(int nT, int nV, int nbT, const int4* __restrict__ tets, const double* __restrict__ DmInv,
                                                const double* __restrict__ vol, const double* __restrict__ mu,
                                                const double* __restrict__ lam, const double* __restrict__ x,
                                                const double* __restrict__ xt, const double* __restrict__ mass, double coef,
                                                double* __restrict__ partial, unsigned* __restrict__ counter, double* __restrict__ out,
                                                double* __restrict__ per_elem) {
    __shared__ double sh[ETPB / 32];
    __shared__ bool last;
    double e = 0.0;
    if ((int)blockIdx.x < nbT) {  // grid-stride over the tets: a bounded number of partials keeps the final pass short
        for (int t = blockIdx.x * ETPB + threadIdx.x; t < nT; t += nbT * ETPB) {
            TetIn in;
            load_tet(t, nT, tets, DmInv, vol, mu, lam, x, in);
            const double et = energy_density<EN>(in.F, in.mu, in.lam) * in.vol;
            if (per_elem) per_elem[t] = et;
            e += et;
        }
    } else {
        const int nbV = gridDim.x - nbT;
        for (int v = (blockIdx.x - nbT) * ETPB + threadIdx.x; v < nV; v += nbV * ETPB) {
            double a = x[3 * (size_t)v] - xt[3 * (size_t)v], b = x[3 * (size_t)v + 1] - xt[3 * (size_t)v + 1],
                   c = x[3 * (size_t)v + 2] - xt[3 * (size_t)v + 2];
            e += (a * a + b * b + c * c) * mass[v] / 2.0;
        }
    }
    double s = block_sum<ETPB>(e, sh);
    if (!out) return;  // per-element mode
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int nb = gridDim.x;
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nbT; i += ETPB) a += __ldcg(partial + i);
    for (int i = nbT + threadIdx.x; i < nb; i += ETPB) b += __ldcg(partial + i);
    __syncthreads();
    double sa = block_sum<ETPB>(a, sh);
    __syncthreads();
    double sb = block_sum<ETPB>(b, sh);
    if (threadIdx.x == 0) {
        out[0] = coef * sa + sb;
        *counter = 0u;
    }
}

// K2, stage 1: elemental gradients of TPB tets, summed per vertex INSIDE the CTA before anything goes to memory.
// The mesh is static, so the host precomputed per CTA the sorted list of distinct vertices its tets touch and, per such
// local vertex, the (tet, corner) pairs that land on it (ascending): thread j adds them up from shared memory in that fixed
// order and writes ONE 24-byte partial per (CTA, vertex) instead of 96 bytes per tet.  Deterministic, no atomics.
template <int EN>
__global__ void __launch_bounds__(TPB) k_grad_block(int nT, const int4* __restrict__ tets, const double* __restrict__ DmInv,
                                                    const double* __restrict__ vol, const double* __restrict__ mu,
                                                    const double* __restrict__ lam, const double* __restrict__ x, double coef,
                                                    const int* __restrict__ lv_ptr, const unsigned short* __restrict__ cptr,
                                                    const unsigned short* __restrict__ cidx, double* __restrict__ part,
                                                    double* __restrict__ epartial, const int* __restrict__ go) {
    if (go && *go == 0) return;  // speculatively enqueued iteration whose assumption failed (linalg.h)
    __shared__ double sg[12 * TPB];
    __shared__ double she[TPB / 32];
    const int t = blockIdx.x * TPB + threadIdx.x;
    double e_t = 0.0;  // vol_t * Psi_t: the energy of the same state comes for free (K1 fused into K2 inside the iteration)
    if (t < nT) {
        TetIn in;
        load_tet(t, nT, tets, DmInv, vol, mu, lam, x, in);
        Mat3 P;
        double psi;
        first_piola<EN>(in.F, in.mu, in.lam, P, psi);
        e_t = psi * in.vol;
        const double w = coef * in.vol;
        // g_e[3+3a+b] = w * Dm^-1.row(a) . P.row(b)   (IglUtils.cpp:857-868); corner 0 = -(sum of the others)
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double g0 = w * (in.B[3 * a] * P(0, 0) + in.B[3 * a + 1] * P(0, 1) + in.B[3 * a + 2] * P(0, 2));
            const double g1 = w * (in.B[3 * a] * P(1, 0) + in.B[3 * a + 1] * P(1, 1) + in.B[3 * a + 2] * P(1, 2));
            const double g2 = w * (in.B[3 * a] * P(2, 0) + in.B[3 * a + 1] * P(2, 1) + in.B[3 * a + 2] * P(2, 2));
            sg[(3 + 3 * a) * TPB + threadIdx.x] = g0;
            sg[(4 + 3 * a) * TPB + threadIdx.x] = g1;
            sg[(5 + 3 * a) * TPB + threadIdx.x] = g2;
            s0 += g0; s1 += g1; s2 += g2;
        }
        sg[threadIdx.x] = -s0;
        sg[TPB + threadIdx.x] = -s1;
        sg[2 * TPB + threadIdx.x] = -s2;
    }
    if (epartial) {  // uniform branch; block_sum contains the barrier that also publishes sg
        const double se = block_sum<TPB>(e_t, she);
        if (threadIdx.x == 0) epartial[blockIdx.x] = se;
    }
    __syncthreads();
    const int p0 = lv_ptr[blockIdx.x], nl = lv_ptr[blockIdx.x + 1] - p0;
    const unsigned short* __restrict__ cp = cptr + p0 + blockIdx.x;          // nl + 1 entries per CTA
    const unsigned short* __restrict__ ci = cidx + (size_t)blockIdx.x * 4 * TPB;
    for (int j = threadIdx.x; j < nl; j += TPB) {
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        for (int e = cp[j]; e < cp[j + 1]; ++e) {
            const int code = ci[e], i = code >> 2, k = code & 3;
            g0 += sg[(3 * k) * TPB + i];
            g1 += sg[(3 * k + 1) * TPB + i];
            g2 += sg[(3 * k + 2) * TPB + i];
        }
        double* __restrict__ o = part + 3 * (size_t)(p0 + j);
        o[0] = g0; o[1] = g1; o[2] = g2;
    }
}

// K2, stage 2: per vertex, add the CTA partials (ascending CTA = ascending tet order), the inertia term m (x - xTilde)
// (Optimizer.cpp:1239-1252), zero the Dirichlet vertices (Energy.cpp:561).
// sum of a vertex's per-CTA gradient partials, in list (= ascending CTA = ascending tet) order.  Four list entries at a time: the
// indices are fetched first, then the 12 values, so that a vertex costs two dependent memory latencies per four partials instead of
// eight (the stage-2 kernels are latency-bound: ~5 partials per vertex, one thread per vertex)
__device__ __forceinline__ void gather_partials(const double* __restrict__ part, const int* __restrict__ vp_idx, int b, int e, double (&gv)[3]) {
    for (int i = b; i < e; i += 4) {
        int id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) id[u] = i + u < e ? vp_idx[i + u] : -1;
        double q[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (id[u] >= 0) {
                const double* __restrict__ p = part + 3 * (size_t)id[u];
                q[u][0] = p[0]; q[u][1] = p[1]; q[u][2] = p[2];
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (id[u] >= 0) {
                gv[0] += q[u][0]; gv[1] += q[u][1]; gv[2] += q[u][2];
            }
    }
}

__global__ void __launch_bounds__(256) k_grad_vertex(int nV, const int* __restrict__ vp_ptr, const int* __restrict__ vp_idx,
                                                     const double* __restrict__ part, const unsigned char* __restrict__ fixed,
                                                     const double* __restrict__ x, const double* __restrict__ xt,
                                                     const double* __restrict__ mass, double* __restrict__ g) {
    int v = blockIdx.x * 256 + threadIdx.x;
    if (v >= nV) return;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    if (!fixed[v]) {
        double gs[3] = {0.0, 0.0, 0.0};
        gather_partials(part, vp_idx, vp_ptr[v], vp_ptr[v + 1], gs);
        g0 = gs[0]; g1 = gs[1]; g2 = gs[2];
        if (xt) {
            const double m = mass[v];
            g0 += m * (x[3 * (size_t)v] - xt[3 * (size_t)v]);
            g1 += m * (x[3 * (size_t)v + 1] - xt[3 * (size_t)v + 1]);
            g2 += m * (x[3 * (size_t)v + 2] - xt[3 * (size_t)v + 2]);
        }
    }
    g[3 * (size_t)v] = g0;
    g[3 * (size_t)v + 1] = g1;
    g[3 * (size_t)v + 2] = g2;
}

// K2 stage 2 on several GPUs: this rank's share of [g ; E] (its own tets; the inertia terms on rank 0, which passes xt) goes straight
// into its slot in EVERY rank's peer buffer - the first half of the all-reduce (peer_reduce.cu) - and the last CTA adds the energy
// partials (elastic: one per CTA of k_grad_block, inertia: one per CTA of this kernel; fixed order), stores the energy at index 3 nV
// of the slots and publishes the epoch.  The consumer (k_pair_dots) adds the slots in rank order.
__global__ void __launch_bounds__(256) k_grad_vertex_push(int nV, const int* __restrict__ vp_ptr, const int* __restrict__ vp_idx,
                                                          const double* __restrict__ part, const unsigned char* __restrict__ fixed,
                                                          const double* __restrict__ x, const double* __restrict__ xt,
                                                          const double* __restrict__ mass, PeerDst D, const double* __restrict__ epartial,
                                                          int n_epartial, double coef, double* __restrict__ ipartial) {
    __shared__ double shm[8];
    __shared__ bool last;
    double ein = 0.0;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < nV; v += gridDim.x * 256) {
        double gv[3] = {0.0, 0.0, 0.0}, dx[3] = {0.0, 0.0, 0.0};
        double m = 0.0;
        if (xt) {
            m = mass[v];
#pragma unroll
            for (int c = 0; c < 3; ++c) dx[c] = x[3 * (size_t)v + c] - xt[3 * (size_t)v + c];
            ein += (dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]) * m / 2.0;  // inertia energy: ALL vertices (Optimizer.cpp:1204-1211)
        }
        if (!fixed[v]) {
            gather_partials(part, vp_idx, vp_ptr[v], vp_ptr[v + 1], gv);
#pragma unroll
            for (int c = 0; c < 3; ++c) gv[c] += m * dx[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) peer_store(D, 3 * (long long)v + c, gv[c]);
    }
    // CTA sum of the inertia energy (fixed order), one partial per CTA
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ein += __shfl_down_sync(0xffffffffu, ein, o);
        if (lane == 0) shm[warp] = ein;
        __syncthreads();
        if (threadIdx.x == 0) ipartial[blockIdx.x] = ((shm[0] + shm[1]) + (shm[2] + shm[3])) + ((shm[4] + shm[5]) + (shm[6] + shm[7]));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(D.counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        double ve = 0.0, vi = 0.0;
        for (int i = threadIdx.x; i < n_epartial; i += 32) ve += __ldcg(epartial + i);
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) vi += __ldcg(ipartial + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ve += __shfl_down_sync(0xffffffffu, ve, o);
            vi += __shfl_down_sync(0xffffffffu, vi, o);
        }
        if (threadIdx.x == 0) {
            const double E = coef * ve + vi;
            peer_store(D, 3 * (long long)nV, E);
            __threadfence_system();
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < D.world) {
        __threadfence_system();
        st_release_sys(D.flag[threadIdx.x], D.epoch);
    }
    if (threadIdx.x == 0) *D.counter = 0u;
}

// K2 stage 2 fused with the L-BFGS pair update (DOTTimeStepper.cpp:476-493): besides g_new it forms s = alpha p and
// y = g_new - g_old, and takes every inner product the NEXT iteration needs in the same pass:
//   |g_new|^2, y.s, s_i.y, s.y_i (Gram matrix row/column of the new pair) and s_i.g_new, s.g_new (first multi-dot of the next iteration)
// Deterministic: fixed grid, block partials, last block adds them in order.
__global__ void __launch_bounds__(256) k_grad_vertex_pair(int nV, const int* __restrict__ vp_ptr, const int* __restrict__ vp_idx,
                                                          const double* __restrict__ part, const unsigned char* __restrict__ fixed,
                                                          const double* __restrict__ x, const double* __restrict__ xt,
                                                          const double* __restrict__ mass, double* __restrict__ g,
                                                          const double* __restrict__ pdir, const double* __restrict__ g_old,
                                                          double* __restrict__ Sn, double* __restrict__ Yn, int sl,
                                                          const double* __restrict__ alpha_dev, double alpha_host, HistList H,
                                                          double* __restrict__ partial, unsigned* __restrict__ counter, double* __restrict__ sc,
                                                          const double* __restrict__ epartial, int n_epartial, double coef,
                                                          int* __restrict__ flag_out, double target) {
    if (H.go && *H.go == 0) {  // the speculation behind this iteration failed: pass the "stop" on to whatever was enqueued behind it
        if (flag_out && blockIdx.x == 0 && threadIdx.x == 0) *flag_out = 0;
        return;
    }
    constexpr int NACC = 4 + 3 * LB_MAXH;  // gg, ys, s.g, inertia energy, then per history pair: s_i.y, s.y_i, s_i.g
    __shared__ double shm[8 * NACC], res[NACC];
    __shared__ bool last;
    const double alpha = alpha_dev ? *alpha_dev : alpha_host;
    double acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < nV; v += gridDim.x * 256) {
        double gv[3] = {0.0, 0.0, 0.0};
        const double m = mass[v];
        double dx[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) dx[c] = x[3 * (size_t)v + c] - xt[3 * (size_t)v + c];
        acc[3] += (dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]) * m / 2.0;  // inertia energy: ALL vertices (Optimizer.cpp:1204-1211)
        if (!fixed[v]) {
            gather_partials(part, vp_idx, vp_ptr[v], vp_ptr[v + 1], gv);
#pragma unroll
            for (int c = 0; c < 3; ++c) gv[c] += m * dx[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t e = 3 * (size_t)v + c;
            const double gn = gv[c];
            g[e] = gn;
            const double s = alpha * pdir[e], y = gn - g_old[e];
            if (Sn) {
                Sn[e] = s;
                Yn[e] = y;
            }
            acc[0] += gn * gn;
            acc[1] += y * s;
            acc[2] += s * gn;
#pragma unroll
            for (int j = 0; j < LB_MAXH; ++j)
                if (j < H.n) {
                    const double si = H.S[j][e];
                    acc[4 + 3 * j] += si * y;
                    acc[5 + 3 * j] += s * H.Y[j][e];
                    acc[6 + 3 * j] += si * gn;
                }
        }
    }
    const int nacc = 4 + 3 * H.n;
    if (!multi_reduce_256<NACC>(acc, nacc, shm, res, &last, partial, counter)) return;
    if (epartial && threadIdx.x < 32) {  // E = coef * sum of the per-CTA elastic partials (fixed order) + inertia energy
        double v = 0.0;
        for (int i = threadIdx.x; i < n_epartial; i += 32) v += __ldcg(epartial + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) {
            const double Enew = coef * v + res[3];
            sc[SC_E] = Enew;
            if (flag_out) {
                // what the host will decide when it sees these numbers (stepper.cu): step accepted without halving, pair kept,
                // not converged -> the iteration enqueued behind this one may run
                const int ok = (Enew <= sc[SC_EPREV]) && (sl < 0 || res[1] > 0.0) && (res[0] > target);
                if (ok) sc[SC_EPREV] = Enew;
                *flag_out = ok;
            }
        }
    }
    if ((int)threadIdx.x < nacc) {
        const int j = threadIdx.x;
        const double tot = res[j];
        if (j == 0) sc[SC_GG] = tot;
        else if (j == 1) { sc[SC_YS_NEW] = tot; if (sl >= 0) sc[SC_SY + 8 * sl + sl] = tot; }
        else if (j == 2) { if (sl >= 0) sc[SC_SG + sl] = tot; }
        else if (j == 3) { }
        else {
            const int h = (j - 4) / 3, kind = (j - 4) % 3, sh_ = H.slot[h];
            if (kind == 0) { if (sl >= 0) sc[SC_SY + 8 * sh_ + sl] = tot; }       // s_h . y_new
            else if (kind == 1) { if (sl >= 0) sc[SC_SY + 8 * sl + sh_] = tot; }  // s_new . y_h
            else sc[SC_SG + sh_] = tot;                                             // s_h . g_new
        }
    }
}

__global__ void __launch_bounds__(TPB) k_svd(int nT, const int4* __restrict__ tets, const double* __restrict__ DmInv,
                                             const double* __restrict__ vol, const double* __restrict__ mu,
                                             const double* __restrict__ lam, const double* __restrict__ x, double* Fo, double* Uo,
                                             double* So, double* Vo) {
    int t = blockIdx.x * TPB + threadIdx.x;
    if (t >= nT) return;
    TetIn in;
    load_tet(t, nT, tets, DmInv, vol, mu, lam, x, in);
    Mat3 U, V;
    double S[3];
    svd3(in.F, U, S, V);
    for (int i = 0; i < 9; ++i) {
        if (Fo) Fo[(size_t)9 * t + i] = in.F.m[i];
        if (Uo) Uo[(size_t)9 * t + i] = U.m[i];
        if (Vo) Vo[(size_t)9 * t + i] = V.m[i];
    }
    if (So) for (int i = 0; i < 3; ++i) So[(size_t)3 * t + i] = S[i];
}

// K3: elemental PD-projected Hessians.  Only the 10 unique 3x3 blocks (k <= l) of the symmetric 12x12 matrix are stored
// (He[nT][10][9], 720 B/tet instead of 1152): block (l,k) is the transpose of (k,l).  Every thread stages its 90 values in
// shared memory and the CTA then writes its tets' contiguous 128 x 720 B region with fully coalesced stores (the direct
// per-thread stores had a 1152-byte stride between lanes and ran at ~19 % of the HBM roofline).
constexpr int HE_BLOCKS = 10, HE_DBL = 90, HE_LD = 91;  // 91: odd stride -> conflict-free staging
__host__ __device__ inline int he_block(int k, int l) { return k * 4 - k * (k - 1) / 2 + (l - k); }  // k <= l

template <int EN>
__global__ void __launch_bounds__(TPB) k_hessian(int nT, const int4* __restrict__ tets, const double* __restrict__ DmInv,
                                                 const double* __restrict__ vol, const double* __restrict__ mu,
                                                 const double* __restrict__ lam, const double* __restrict__ x, double coef,
                                                 int project, double* __restrict__ He) {
    extern __shared__ double stage[];  // [TPB][HE_LD]
    const int t0 = blockIdx.x * TPB;
    const int t = t0 + threadIdx.x;
    if (t < nT) {
        TetIn in;
        load_tet(t, nT, tets, DmInv, vol, mu, lam, x, in);
        Mat3 U, V;
        double S[3];
        svd3(in.F, U, S, V);
        HessCoef hc;
        hess_coef<EN>(S, in.mu, in.lam, coef * in.vol, project != 0, hc);
        // beta[k][b] = v_b . w_k,  w_k = row k-1 of Dm^-1 (k=1..3), w_0 = -(sum of rows)
        double beta[4][3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                beta[k + 1][b] = in.B[3 * k] * V(0, b) + in.B[3 * k + 1] * V(1, b) + in.B[3 * k + 2] * V(2, b);
            beta[0][b] = -(beta[1][b] + beta[2][b] + beta[3][b]);
        }
        double* out = stage + threadIdx.x * HE_LD;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int l = k; l < 4; ++l) {
                double blk[9];
                hess_block(hc, U, beta[k], beta[l], blk);
                if (k == l) {  // symmetrise the diagonal block exactly (the reference mirrors the upper part, Energy.cpp:1263-1265)
                    blk[3] = blk[1]; blk[6] = blk[2]; blk[7] = blk[5];
                }
#pragma unroll
                for (int i = 0; i < 9; ++i) out[he_block(k, l) * 9 + i] = blk[i];
            }
        }
    }
    __syncthreads();
    const int ntet = min(TPB, nT - t0);
    double* __restrict__ dst = He + (size_t)HE_DBL * t0;
    for (int e = threadIdx.x; e < ntet * HE_DBL; e += TPB) {
        const int tt = e / HE_DBL, j = e - tt * HE_DBL;
        dst[e] = stage[tt * HE_LD + j];
    }
}

// [nT][10][9] unique blocks -> row-major 12x12 (the reference's elemHessian layout)
__global__ void k_he_to_dense(int nT, const double* __restrict__ He, double* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nT * 144) return;
    int t = (int)(i / 144), e = (int)(i % 144);
    int row = e / 12, col = e % 12;
    int k = row / 3, ii = row % 3, l = col / 3, r = col % 3;
    out[i] = k <= l ? He[(size_t)HE_DBL * t + he_block(k, l) * 9 + 3 * ii + r] : He[(size_t)HE_DBL * t + he_block(l, k) * 9 + 3 * r + ii];
}

}  // namespace

void DeviceMesh::init(int energy_type, int nV_, int nT_, const int32_t* tets_h, const double* DmInv_rm, const double* vol_h,
                      const double* mu_h, const double* lam_h, const double* mass_h, const unsigned char* fixed_h,
                      cudaStream_t st) {
    DG_REQUIRE(energy_type == DOTGPU_ENERGY_FCR || energy_type == DOTGPU_ENERGY_SNH, "unknown energy type");
    DG_REQUIRE(nV_ > 0 && nT_ > 0, "empty mesh");
    energy = energy_type;
    nV = nV_;
    nT = nT_;
    for (size_t i = 0; i < (size_t)4 * nT; ++i) DG_REQUIRE(tets_h[i] >= 0 && tets_h[i] < nV, "tet index out of range");
    tets.upload(tets_h, (size_t)4 * nT, st);
    std::vector<double> soa((size_t)9 * nT);
    for (int t = 0; t < nT; ++t)
        for (int j = 0; j < 9; ++j) soa[(size_t)j * nT + t] = DmInv_rm[(size_t)9 * t + j];
    DmInv.upload(soa, st);
    vol.upload(vol_h, nT, st);
    mu.upload(mu_h, nT, st);
    lam.upload(lam_h, nT, st);
    if (mass_h) mass.upload(mass_h, nV, st);
    std::vector<unsigned char> fx(nV, 0);
    if (fixed_h) fx.assign(fixed_h, fixed_h + nV);
    fixed.upload(fx, st);
    // K2 tables: per CTA of TPB tets the distinct vertices (ascending) and their (tet, corner) lists; per vertex the partials
    {
        const int nb = ceil_div(nT, TPB);
        std::vector<int> lvp(nb + 1, 0), lvv;
        std::vector<unsigned short> cp, ci((size_t)nb * 4 * TPB, 0);
        std::vector<std::pair<int, int>> corners;  // (vertex, code)
        for (int b = 0; b < nb; ++b) {
            corners.clear();
            const int t0 = b * TPB, t1 = std::min(nT, t0 + TPB);
            for (int t = t0; t < t1; ++t)
                for (int k = 0; k < 4; ++k) corners.emplace_back(tets_h[4 * (size_t)t + k], 4 * (t - t0) + k);
            std::sort(corners.begin(), corners.end());  // by vertex, then ascending (tet, corner)
            int nl = 0;
            for (size_t e = 0; e < corners.size(); ++e) {
                if (e == 0 || corners[e].first != corners[e - 1].first) {
                    cp.push_back((unsigned short)e);
                    lvv.push_back(corners[e].first);
                    ++nl;
                }
                ci[(size_t)b * 4 * TPB + e] = (unsigned short)corners[e].second;
            }
            cp.push_back((unsigned short)corners.size());
            lvp[b + 1] = lvp[b] + nl;
        }
        const int npart = lvp[nb];
        std::vector<int> vptr(nV + 1, 0), vidx(npart);
        for (int i = 0; i < npart; ++i) vptr[lvv[i] + 1]++;
        for (int v = 0; v < nV; ++v) vptr[v + 1] += vptr[v];
        std::vector<int> cur(vptr.begin(), vptr.end() - 1);
        for (int i = 0; i < npart; ++i) vidx[cur[lvv[i]]++] = i;  // ascending partial index = ascending CTA
        lv_ptr.upload(lvp, st);
        g_cptr.upload(cp, st);
        g_cidx.upload(ci, st);
        vp_ptr.upload(vptr, st);
        if (vidx.empty()) vidx.push_back(0);
        vp_idx.upload(vidx, st);
        gpart.alloc(3 * (size_t)std::max(npart, 1));
    }
    {   // reduction grids are sized from the SM count of the device this mesh lives on
        int dev = 0;
        DG_CUDA(cudaGetDevice(&dev));
        DG_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
        // the opt-in shared-memory limit is per device and per function: set it for the device of this mesh
        DG_CUDA(cudaFuncSetAttribute(k_hessian<DG_FCR>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPB * HE_LD * (int)sizeof(double)));
        DG_CUDA(cudaFuncSetAttribute(k_hessian<DG_SNH>, cudaFuncAttributeMaxDynamicSharedMemorySize, TPB * HE_LD * (int)sizeof(double)));
    }
    n_partial = nsm * 8 + 8;
    partial.alloc(n_partial);
    counter.alloc(1);
    counter.zero(st);
    DG_CUDA(cudaStreamSynchronize(st));
}

void DeviceMesh::set_fixed(const unsigned char* fixed_h, cudaStream_t st) {
    fixed.upload(fixed_h, nV, st);
    DG_CUDA(cudaStreamSynchronize(st));
}

#define DISPATCH_EN_SMEM(m, KERNEL, grid, block, smem, st, ...)                 \
    do {                                                                       \
        if ((m).energy == DOTGPU_ENERGY_FCR)                                   \
            KERNEL<DG_FCR><<<grid, block, smem, st>>>(__VA_ARGS__);            \
        else                                                                   \
            KERNEL<DG_SNH><<<grid, block, smem, st>>>(__VA_ARGS__);            \
        count_launch();                                                        \
    } while (0)

#define DISPATCH_EN(m, KERNEL, grid, block, st, ...)                           \
    do {                                                                       \
        if ((m).energy == DOTGPU_ENERGY_FCR)                                   \
            KERNEL<DG_FCR><<<grid, block, 0, st>>>(__VA_ARGS__);               \
        else                                                                   \
            KERNEL<DG_SNH><<<grid, block, 0, st>>>(__VA_ARGS__);               \
        count_launch();                                                        \
    } while (0)

void launch_energy(DeviceMesh& m, const double* x, const double* xTilde, double coef, double* E_out, cudaStream_t st) {
    // at most 6 tet CTAs + 1 inertia CTA of 256 threads per SM: <= 7 * #SM partials for the last block
    const int nbT = std::min(ceil_div(m.nT, ETPB), m.nsm * 6), nbV = xTilde ? std::min(ceil_div(m.nV, ETPB), m.nsm) : 0;
    DISPATCH_EN(m, k_energy, nbT + nbV, ETPB, st, m.nT, m.nV, nbT, (const int4*)m.tets.p, m.DmInv.p, m.vol.p, m.mu.p, m.lam.p, x, xTilde,
                m.mass.p, coef, m.partial.p, m.counter.p, E_out, (double*)nullptr);
}

void launch_energy_per_elem(DeviceMesh& m, const double* x, double* out, cudaStream_t st) {
    const int nbT = std::min(ceil_div(m.nT, ETPB), m.nsm * 6);
    DISPATCH_EN(m, k_energy, nbT, ETPB, st, m.nT, m.nV, nbT, (const int4*)m.tets.p, m.DmInv.p, m.vol.p, m.mu.p, m.lam.p, x,
                (const double*)nullptr, (const double*)nullptr, 1.0, (double*)nullptr, (unsigned*)nullptr, (double*)nullptr, out);
}

void launch_gradient(DeviceMesh& m, const double* x, const double* xTilde, double coef, double* g, cudaStream_t st) {
    int nb = ceil_div(m.nT, TPB);
    DISPATCH_EN(m, k_grad_block, nb, TPB, st, m.nT, (const int4*)m.tets.p, m.DmInv.p, m.vol.p, m.mu.p, m.lam.p, x, coef, m.lv_ptr.p,
                m.g_cptr.p, m.g_cidx.p, m.gpart.p, (double*)nullptr, (const int*)nullptr);
    k_grad_vertex<<<ceil_div(m.nV, 256), 256, 0, st>>>(m.nV, m.vp_ptr.p, m.vp_idx.p, m.gpart.p, m.fixed.p, x, xTilde, m.mass.p, g);
    count_launch();
}

void launch_gradient_push(DeviceMesh& m, const double* x, const double* xTilde, double coef, const PeerDst& D, cudaStream_t st) {
    const int nb = ceil_div(m.nT, TPB);
    if (m.epart.n < (size_t)std::max(nb, 1)) m.epart.alloc(std::max(nb, 1));
    if (nb > 0)
        DISPATCH_EN(m, k_grad_block, nb, TPB, st, m.nT, (const int4*)m.tets.p, m.DmInv.p, m.vol.p, m.mu.p, m.lam.p, x, coef, m.lv_ptr.p,
                    m.g_cptr.p, m.g_cidx.p, m.gpart.p, m.epart.p, (const int*)nullptr);
    const int grid = std::min(ceil_div(m.nV, 256), m.nsm * 4);  // <= n_partial inertia partials
    k_grad_vertex_push<<<grid, 256, 0, st>>>(m.nV, m.vp_ptr.p, m.vp_idx.p, m.gpart.p, m.fixed.p, x, xTilde, m.mass.p, D, m.epart.p, nb, coef,
                                             m.partial.p);
    count_launch();
}

void launch_gradient_pair(DeviceMesh& m, const double* x, const double* xTilde, double coef, double* g, const double* pdir,
                          const double* g_old, double* S_new, double* Y_new, int sl, const double* alpha_dev, double alpha_host,
                          const HistList& H, double* partial, unsigned* counter, double* sc, bool with_energy, cudaStream_t st,
                          int* flag_out, double target) {
    int nb = ceil_div(m.nT, TPB);
    if (with_energy && m.epart.n < (size_t)nb) m.epart.alloc(nb);
    DISPATCH_EN(m, k_grad_block, nb, TPB, st, m.nT, (const int4*)m.tets.p, m.DmInv.p, m.vol.p, m.mu.p, m.lam.p, x, coef, m.lv_ptr.p,
                m.g_cptr.p, m.g_cidx.p, m.gpart.p, with_energy ? m.epart.p : (double*)nullptr, H.go);
    k_grad_vertex_pair<<<multidot_blocks(m.nV), 256, 0, st>>>(m.nV, m.vp_ptr.p, m.vp_idx.p, m.gpart.p, m.fixed.p, x, xTilde, m.mass.p, g, pdir,
                                                              g_old, S_new, Y_new, sl, alpha_dev, alpha_host, H, partial, counter, sc,
                                                              with_energy ? m.epart.p : (const double*)nullptr, nb, coef, flag_out, target);
    count_launch();
}

void launch_svd(DeviceMesh& m, const double* x, double* F, double* U, double* S, double* V, cudaStream_t st) {
    k_svd<<<ceil_div(m.nT, TPB), TPB, 0, st>>>(m.nT, (const int4*)m.tets.p, m.DmInv.p, m.vol.p, m.mu.p, m.lam.p, x, F, U, S, V);
    count_launch();
}

void launch_elem_hessians(DeviceMesh& m, const double* x, double coef, bool project, cudaStream_t st) {
    if (m.He.n < (size_t)HE_DBL * m.nT) m.He.alloc((size_t)HE_DBL * m.nT);
    int nb = ceil_div(m.nT, TPB);
    DISPATCH_EN_SMEM(m, k_hessian, nb, TPB, TPB * HE_LD * sizeof(double), st, m.nT, (const int4*)m.tets.p, m.DmInv.p, m.vol.p, m.mu.p, m.lam.p, x, coef,
                project ? 1 : 0, m.He.p);
}

void launch_he_to_dense(DeviceMesh& m, double* out144, cudaStream_t st) {
    size_t n = (size_t)m.nT * 144;
    k_he_to_dense<<<ceil_div(n, 256), 256, 0, st>>>(m.nT, m.He.p, out144);
    count_launch();
}

}  // namespace dotgpu
