// K5: the triangular solves of ALL subdomain factors as ONE persistent, TMA-streamed dataflow kernel.
//
// What it replaces: cholmod_solve(CHOLMOD_A) per subdomain (LinSysSolver/CHOLMODSolver.cpp:149-163 ->
// SuiteSparse/CHOLMOD/Supernodal/t_cholmod_super_solve.c), called k times per L-BFGS iteration from
// DOTTimeStepper::solve_oneStep (TimeStepper/DOTTimeStepper.cpp:406-433).
//
// Design (B200): a solve is a pure stream over the factor - every entry of the "solve panels"
// S_s = [L_ss^-1 ; -L_below L_ss^-1] is used exactly once forward and once backward - so the kernel is built like a
// bandwidth benchmark with a dependency graph on top:
//   * panels are stored PACKED and in BOTH orientations (Pf: rows of S_s, triangle packed; Pb: columns of S_s =
//     rows of S_s^T, triangle packed), so that every task (a range of rows) is one contiguous byte range;
//   * a producer warp walks its CTA's share of the chunk queue (topological order: forward levels up, then backward levels
//     down) and streams the byte ranges into a shared-memory ring with 1-D bulk TMA copies (cp.async.bulk +
//     mbarrier complete_tx).  The factor is read-only during a solve, so the stream runs AHEAD of the dependency
//     chain: when a supernode's inputs become ready its panel rows are already in shared memory;
//   * 4 gatherer warps wait for the chunk's dependency (per-supernode completion counters, acquire loads) and build the
//     right-hand side / update vector of the supernode in shared memory, one supernode ahead of the consumers;
//   * 4 consumer warps wait for the data and the vector (mbarriers), do the row.vector products out of shared memory and
//     store the results; a signaller lane per ring stage publishes the chunk (release reduction on the waiting counter);
//   * the queue slots are dealt to the CTAs round-robin (CTA c: slots c, c+G, ...), so the chunks of one tree level run on
//     different CTAs at the same time and every dependency of a slot sits earlier in some resident CTA's list:
//     no deadlock for any number of resident CTAs (several solver handles may run concurrently).
// All sums have a fixed order (lane-strided partial sums + shuffle tree, children in ascending order): results are
// bit-reproducible from run to run.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "chol_numeric.h"

namespace dotgpu {

namespace {

constexpr int SOLVE_GROUP_MAX = 4;  // chunks per queue slot (they share one gather of the supernode's vector); slots are padded to it
constexpr int CONSUMERS = 128;   // 4 warps: row . vector products
#ifndef DOTGPU_GATHERERS
#define DOTGPU_GATHERERS 128
#endif
constexpr int GATHERERS = DOTGPU_GATHERERS;   // warps x 32: dependency waits + vector gathers, one supernode ahead of the consumers
constexpr int SOLVE_THREADS = CONSUMERS + 32 + GATHERERS + 32;  // consumers, producer warp, gatherers, signaller warp

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: a waiting thread sleeps in hardware until the phase completes (or the hint expires) instead
// of re-polling - the helper warps (producer, gatherers, signallers) share their SM sub-partitions with the consumer warps, and
// their polling loops were taking issue slots from the products (chunk trace r2: the consumer warp that shares its sub-partition
// with the producer took 1.3 us per chunk, the one that does not 0.6 us)
#ifndef DOTGPU_MBAR_SUSPEND_NS
#define DOTGPU_MBAR_SUSPEND_NS 20000
#endif
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((unsigned)DOTGPU_MBAR_SUSPEND_NS)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk TMA copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both sides 16-B aligned)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// DOTGPU_SOLVE_TRACE=1 (experiments, tools/k5_trace.py): %globaltimer stamps of the first 96 chunks of every CTA, 6 per chunk
// {posted, vector ready, data + vector seen by the consumers, products done (after the consumer barrier), published, queue slot,
// warp 0 out of its products, last consumer warp out of its products}; 8 words per chunk
constexpr int TRACE_CHUNKS = 96;
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DG_TRACE(slot, it_)                                                                                       \
    do {                                                                                                          \
        if (trace && (it_) < TRACE_CHUNKS) trace[((size_t)blockIdx.x * TRACE_CHUNKS + (it_)) * 8 + (slot)] = gtime(); \
    } while (0)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }
__device__ __forceinline__ void gatherer_sync() { asm volatile("bar.sync 2, %0;" ::"n"(GATHERERS) : "memory"); }

// counters: [0,nsn) forward chunks done per supernode, [nsn,2nsn) backward chunks done, [2nsn,3nsn) children whose forward
// phase is complete
//
// Warp roles (CTA = 4 consumer warps + producer warp + 4 gatherer warps + signaller warp):
//   producer  (31 lanes)  walks the CTA's queue slots, posts chunk descriptors, issues the TMA copies  -> posted[st], full[st]
//   signaller (1 lane)    publishes finished chunks: fence + completion counters                       <- done[st] -> empty[st]
//   gatherers (128)       per chunk: waits for the chunk's dependency and builds the supernode's vector
//                         (right-hand side + children updates / ancestors' solution) in shared memory  <- posted[st] -> vready[st]
//   consumers (128)       row . vector products out of shared memory, results to global memory         <- full, vready -> done[st]
// The global-memory latencies (queue, descriptors, dependency polls, gathers, fences) all sit in the helper warps and overlap the
// consumers' work on earlier chunks.
template <int NSTAGE>
__global__ void __launch_bounds__(SOLVE_THREADS, SOLVE_THREADS <= 256 ? 4 : 3)
    k_solve_stream(int ngroups, const SolveTask* __restrict__ chunks, const int* __restrict__ rows,
                   const int* __restrict__ rel, const double* __restrict__ Pf, const double* __restrict__ Pb, const double* __restrict__ b,
                   const int* __restrict__ gidx, double* y, double* U, double* x, unsigned* cnt, int stage_dbl,
                   int vec_dbl, int dbg, const int* __restrict__ go, unsigned long long* trace) {
    if (go && *go == 0) return;  // speculatively enqueued iteration whose assumption failed (linalg.h): every CTA leaves at once
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);          // NSTAGE * stage_dbl
    double* vecs = ring + (size_t)NSTAGE * stage_dbl;             // 2 x [vec_dbl] vectors (double-buffered across supernodes)
    int* idxs = reinterpret_cast<int*>(vecs + 2 * (size_t)vec_dbl);  // 2 x [vec_dbl] gather indices / parent positions
    __shared__ __align__(8) unsigned long long full[NSTAGE], empty[NSTAGE], done[NSTAGE], posted[NSTAGE], vready[NSTAGE];
    __shared__ SolveTask s_chunk[NSTAGE];
    __shared__ int s_live[NSTAGE], s_vb[NSTAGE], s_base[NSTAGE];
    __shared__ __align__(16) SolveTask s_desc[2][SOLVE_GROUP_MAX];

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&done[i], 1);
            mbar_init(&posted[i], 1);
            mbar_init(&vready[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= CONSUMERS && tid < CONSUMERS + 31) {
        // ------------------------------ producer (31 lanes): stream the chunks of this CTA's queue slots ------------------------------
        // The chunk descriptors of a slot are requested one slot ahead (cp.async into a double buffer), so the TMA issue rate is
        // not bound by load latency.
        constexpr unsigned PM = 0x7fffffffu;
        const int lane = tid - CONSUMERS;
        auto fetch = [&](int buf, unsigned g) {
            if (g < (unsigned)ngroups) {
                const char* src = reinterpret_cast<const char*>(chunks + (size_t)g * SOLVE_GROUP_MAX);
                char* dst = reinterpret_cast<char*>(&s_desc[buf][0]);
                for (int i = lane; i < SOLVE_GROUP_MAX * (int)sizeof(SolveTask) / 16; i += 31)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 16 * i)), "l"(src + 16 * i) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // STATIC assignment of the queue slots: CTA c takes the slots c, c + G, c + 2G, ...  Consecutive slots - the chunks of one
        // tree level - therefore run on different CTAs at the same time (a dynamic queue with claims prefetched one slot ahead
        // hands consecutive slots to the SAME CTA, which halves the parallelism of the latency-bound top levels), and no
        // atomics are needed.  Every dependency of a slot has a smaller index and sits earlier in some resident CTA's list:
        // no deadlock for any number of resident CTAs.
        const unsigned G = gridDim.x;
        unsigned g0 = blockIdx.x;
        fetch(0, g0);
        int it = 0, buf = 0;
        while (g0 < (unsigned)ngroups) {
            const unsigned g1 = g0 + G;
            fetch(buf ^ 1, g1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp(PM);
            for (int c = 0; c < SOLVE_GROUP_MAX; ++c) {
                const SolveTask& T = s_desc[buf][c];
                const int ndbl = T.ndbl;
                if (ndbl == 0) continue;  // padding of a short group
                const int st = it % NSTAGE;
                if (it >= NSTAGE) {
                    if (lane == 0) mbar_wait(&empty[st], ((it / NSTAGE) - 1) & 1);
                    __syncwarp(PM);
                }
                if (lane < (int)sizeof(SolveTask) / 4) reinterpret_cast<int*>(&s_chunk[st])[lane] = reinterpret_cast<const int*>(&T)[lane];
                if (lane == 0) s_live[st] = 1;
                __syncwarp(PM);
                if (lane == 0) {
                    mbar_arrive(&posted[st]);
                    const unsigned bytes = (unsigned)ndbl * 8u;
                    if (dbg & 2) {
                        mbar_arrive(&full[st]);
                    } else {
                        mbar_expect_tx(&full[st], bytes);
                        tma_load_1d(ring + (size_t)st * stage_dbl, (s_chunk[st].kind ? Pb : Pf) + s_chunk[st].src, bytes, &full[st]);
                    }
                    DG_TRACE(0, it);
                    if (trace && it < TRACE_CHUNKS) trace[((size_t)blockIdx.x * TRACE_CHUNKS + it) * 8 + 5] = g0;
                }
                ++it;
            }
            __syncwarp(PM);
            g0 = g1;
            buf ^= 1;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (lane == 0) {
            // end of stream: a sentinel in every stage (each signaller lane watches one stage)
            for (int i = 0; i < NSTAGE; ++i, ++it) {
                const int st = it % NSTAGE;
                if (it >= NSTAGE) mbar_wait(&empty[st], ((it / NSTAGE) - 1) & 1);
                s_live[st] = 0;
                mbar_arrive(&posted[st]);
                mbar_arrive(&full[st]);
            }
        }
        return;
    }
    if (tid == CONSUMERS + 31) return;
    if (tid >= CONSUMERS + 32 + GATHERERS) {
        // ------------------------------ signallers: one lane per stage publishes that stage's finished chunks ------------------------------
        // (release-reduction on the dependency counter of whoever waits for this chunk; off the consumers' critical path, and the
        //  fences of different stages overlap)
        const int st = tid - (CONSUMERS + 32 + GATHERERS);
        if (st >= NSTAGE) return;
        for (unsigned n = 0;; ++n) {
            const unsigned par = n & 1;
            mbar_wait(&posted[st], par);
            if (!s_live[st]) break;
            mbar_wait(&done[st], par);  // every consumer has stored its results and left the stage
            unsigned* p = cnt + s_chunk[st].sig_idx;
            mbar_arrive(&empty[st]);
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
            if (trace && n * NSTAGE + st < TRACE_CHUNKS) trace[((size_t)blockIdx.x * TRACE_CHUNKS + (n * NSTAGE + st)) * 8 + 4] = gtime();
        }
        return;
    }
    if (tid >= CONSUMERS + 32 && tid < CONSUMERS + 32 + GATHERERS) {
        // ------------------------------ gatherers ------------------------------
        const int gl = tid - (CONSUMERS + 32);
        int cur_s = -1, cur_kind = -1, idx_r1 = 0, vb = 1, base = 0, cur_fresh = -1;
        for (int it = 0;; ++it) {
            const int st = it % NSTAGE;
            mbar_wait(&posted[st], (it / NSTAGE) & 1);
            if (!s_live[st]) break;
            const SolveTask& T = s_chunk[st];
            const int m = T.m, ns = T.ns, r0 = T.r0, kind = T.kind, nchild = T.nchild, col0 = T.col0;
            // a supernode's vector is gathered once and reused by the following chunks of the same phase; the forward phase
            // also needs the parent positions of the update rows, fetched per group
            const bool fresh = !(cur_s == T.s && cur_kind == kind) || (kind == 0 && T.r1 > idx_r1);
            if (fresh) {
                cur_s = T.s;
                cur_kind = kind;
                vb ^= 1;
                base = r0;
                // the buffer's previous occupant served the chunks before the previous gather: wait until the consumers left them
                const int j = cur_fresh - 1;
                if (j >= 0 && it - j < NSTAGE) mbar_wait(&done[j % NSTAGE], (j / NSTAGE) & 1);
                cur_fresh = it;
                double* __restrict__ vec = vecs + (size_t)vb * vec_dbl;
                int* __restrict__ idx = idxs + (size_t)vb * vec_dbl;
                const double* __restrict__ Us = U + T.U;
                // ---- loads that do not depend on other supernodes ----
                if (kind == 0) {
                    if (gidx) {
#pragma unroll 4
                        for (int k = gl; k < ns; k += GATHERERS) vec[k] = b[gidx[col0 + k]];
                    } else {
#pragma unroll 4
                        for (int k = gl; k < ns; k += GATHERERS) vec[k] = b[col0 + k];
                    }
                    idx_r1 = T.g_r1;
                    const int* __restrict__ prel = rel + T.rows;
#pragma unroll 4
                    for (int r = max(r0, ns) + gl; r < idx_r1; r += GATHERERS) idx[r - r0] = prel[r];
                } else {
                    const int* __restrict__ frows = rows + T.rows;
#pragma unroll 4
                    for (int r = max(r0, ns) + gl; r < m; r += GATHERERS) idx[r - r0] = frows[r];
                }
                // ---- dependency: one counter, one expected value ----
                if (gl == 0 && T.dep_need > 0 && !(dbg & 4)) {
                    const unsigned need = (unsigned)T.dep_need;
                    const unsigned* p = cnt + T.dep_idx;
                    while (ld_acquire(p) < need) {
                        __nanosleep(40);   // back off: the poll loop shares an SM sub-partition with a consumer warp
                    }
                }
                gatherer_sync();
                // ---- dependent part: contiguous reads of what the children / ancestors published ----
                if (kind == 0) {
                    if (nchild > 0) {
#pragma unroll 2
                        for (int k = gl; k < ns; k += GATHERERS) {
                            double v = vec[k];
                            for (int w = 0; w < nchild; ++w) v += __ldcg(Us + (long long)w * m + k);
                            vec[k] = v;
                        }
                    }
                    // vec[r], r >= ns: what the children add to update row r (rows of this group only)
#pragma unroll 2
                    for (int r = max(r0, ns) + gl; r < idx_r1; r += GATHERERS) {
                        double v = 0.0;
                        for (int w = 0; w < nchild; ++w) v += __ldcg(Us + (long long)w * m + r);
                        vec[r] = v;
                    }
                } else {
#pragma unroll 4
                    for (int r = r0 + gl; r < m; r += GATHERERS) vec[r - r0] = r < ns ? __ldcg(y + col0 + r) : __ldcg(x + idx[r - r0]);
                }
                gatherer_sync();
            }
            if (gl == 0) {
                s_vb[st] = vb;
                s_base[st] = base;
                mbar_arrive(&vready[st]);
                DG_TRACE(1, it);
            }
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    const int lane = tid & 31;
    for (int it = 0;; ++it) {
        const int st = it % NSTAGE;
        const unsigned par = (it / NSTAGE) & 1;
        mbar_wait(&posted[st], par);
        if (!s_live[st]) break;
        mbar_wait(&vready[st], par);
        mbar_wait(&full[st], par);
        if (tid == 0) DG_TRACE(2, it);
        const SolveTask& T = s_chunk[st];
        const int m = T.m, ns = T.ns, r0 = T.r0, r1 = T.r1, col0 = T.col0;
        const int vbase = s_base[st];
        const double* __restrict__ vec = vecs + (size_t)s_vb[st] * vec_dbl;
        const int* __restrict__ idx = idxs + (size_t)s_vb[st] * vec_dbl;
        const double* __restrict__ chunk = ring + (size_t)st * stage_dbl + T.shift;
        // Row . vector products.  W lanes share a row, W = the largest power of two that still gives every row of the chunk its own
        // lane group (W = 1: one thread per row, no reduction; W = 32: a warp per row).  Short rows dominate the factor (update rows
        // of leaf-side supernodes), so most chunks run with small W and finish in one pass with a short or no shuffle tree.
        const int nrows = r1 - r0;
        // W = 2^lw lanes per row.  Shifts and masks only: a division by the run-time W costs ~200 cycles, and this code runs once per
        // chunk in every consumer warp (chunk trace r2: ~700 cycles of fixed overhead per chunk, a third of a chunk's time at 1M tets)
        int lw = 5;
        while (lw > 0 && (nrows << lw) > CONSUMERS) --lw;
        const int W = 1 << lw;
        const int sub = tid >> lw, sl = tid & (W - 1), nsub = CONSUMERS >> lw;
        const bool fwd = T.kind == 0;
        const int tri0 = r0 < ns ? (ns - r0) * (ns + r0 + 1) / 2 : 0;  // forward: entries of the triangle rows r0..ns-1 of this chunk
        const long long pU = T.pU;
        const long long tl0 = trace ? clock64() : 0;
        if (!(dbg & 1))
        for (int i0 = 0; i0 < nrows; i0 += nsub) {
            const int r = r0 + i0 + sub;
            const bool valid = r < r1;
            int off = 0, len = 0;
            const double* __restrict__ v = vec;
            if (valid) {
                if (fwd) {  // row r of S: triangle rows hold r+1 entries, update rows ns
                    off = r < ns ? (r - r0) * (r + r0 + 1) / 2 : tri0 + (r - max(ns, r0)) * ns;
                    len = r < ns ? r + 1 : ns;
                } else {    // packed row r of S^T holds rows r..m-1: offset (r - r0)*m - (r(r-1) - r0(r0-1))/2
                    off = (r - r0) * m - (r * (r - 1) - r0 * (r0 - 1)) / 2;
                    len = m - r;
                    v = vec + (r - vbase);
                }
            }
            const double* __restrict__ row = chunk + off;
            double s0 = 0.0, s1 = 0.0;
            if (W == 1) {
                // lane stride in shared memory = row length; for even lengths rotate the start by the lane so that the stride
                // becomes odd and the accesses stay bank-conflict free
                int k = (len & 1) == 0 && len > 0 ? lane % len : 0;
                int i = 0;
                for (; i + 1 < len; i += 2) {
                    const int k1 = k + 1 == len ? 0 : k + 1;
                    s0 += row[k] * v[k];
                    s1 += row[k1] * v[k1];
                    k = k1 + 1 == len ? 0 : k1 + 1;
                }
                if (i < len) s0 += row[k] * v[k];
            } else {
                int k = sl;
                for (; k + W < len; k += 2 * W) {
                    s0 += row[k] * v[k];
                    s1 += row[k + W] * v[k + W];
                }
                if (k < len) s0 += row[k] * v[k];
            }
            double sum = s0 + s1;
            for (int o = W >> 1; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, W);
            if (valid && sl == 0) {
                if (!fwd) x[col0 + r] = sum;
                else if (r < ns) y[col0 + r] = sum;
                else U[pU + idx[r - vbase]] = vec[r] + sum;
            }
        }
        if (trace && it < TRACE_CHUNKS && (tid & 31) == 0) {   // per-warp cycles inside the product loop + the chunk's shape
            const long long dtc = clock64() - tl0;
            unsigned long long* q = trace + ((size_t)blockIdx.x * TRACE_CHUNKS + it) * 8;
            if (tid == 0) q[6] = (unsigned long long)dtc | ((unsigned long long)nrows << 32) | ((unsigned long long)W << 48) | ((unsigned long long)T.kind << 56);
            if (tid == CONSUMERS - 32) q[7] = (unsigned long long)dtc;
            if (tid == 32) q[4] = (unsigned long long)dtc;   // (slot 4 is overwritten by the signaller later: read it as warp 1 only if published == 0)
        }
        consumer_sync();  // all results stored, all shared-memory reads of this stage (data, descriptor, vector) done
        if (tid == 0) {
            mbar_arrive(&done[st]);
            DG_TRACE(3, it);
        }
    }
}

// right-hand side in the permuted, concatenated numbering of the factors: b_perm[i] = b[gidx[i]] - the gatherers then read
// contiguous ranges (one latency on the critical path of every supernode instead of two dependent ones)
__global__ void __launch_bounds__(256) k_gather_rhs(long long n, const double* __restrict__ b, const int* __restrict__ gidx,
                                                    double* __restrict__ out, const int* __restrict__ go) {
    if (go && *go == 0) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = b[gidx[i]];
}

// ---- pack the row-major solve panels Sp into the two streamed layouts ----
// one CTA per (supernode, 64-row slab a, 64-column tile b) that intersects the stored region (r >= c or r >= ns)
struct Task3p { int s; short a, b; };
__global__ void __launch_bounds__(256) k_pack_panels(const Task3p* __restrict__ tasks, const SolveSN* __restrict__ sn,
                                                     const SNDesc* __restrict__ snd, const double* __restrict__ Sp,
                                                     double* __restrict__ Pf, double* __restrict__ Pb) {
    __shared__ double tile[64][65];
    const Task3p tk = tasks[blockIdx.x];
    const SolveSN d = sn[tk.s];
    const long long panel = snd[tk.s].panel;
    const int r0 = tk.a * 64, c0 = tk.b * 64;
    const int nr = min(64, d.m - r0), nc = min(64, d.ns - c0);
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const long long tri = (long long)d.ns * (d.ns + 1) / 2;
    for (int i = ty; i < nr; i += 4) {
        const int r = r0 + i, c = c0 + tx;
        double v = 0.0;
        if (tx < nc) {
            v = Sp[panel + (long long)r * d.ns + c];
            if (r >= d.ns)
                Pf[d.pf + tri + (long long)(r - d.ns) * d.ns + c] = v;
            else if (c <= r)
                Pf[d.pf + (long long)r * (r + 1) / 2 + c] = v;
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < nc; j += 4) {
        const int c = c0 + j, r = r0 + tx;
        if (tx < nr && r >= c) Pb[d.pb + (long long)c * d.m - (long long)c * (c - 1) / 2 + (r - c)] = tile[tx][j];
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
void CholBatch::build_solve_plan(const std::vector<SNDesc>& sn, cudaStream_t st) {
    const int nsn = nsuper_total;
    std::vector<SolveSN> ssn(nsn);
    int64_t pf = 0, ub = 0;
    int max_row = 2;
    for (int g = 0; g < nsn; ++g) {
        const SNDesc& d = sn[g];
        SolveSN& o = ssn[g];
        o.m = d.m;
        o.ns = d.ns;
        o.col0 = d.col0;
        o.rows = d.rows;
        o.parent = d.parent;
        o.nchild = d.child_end - d.child_begin;
        o.child_begin = d.child_begin;
        o.pf = o.pb = pf;
        const int64_t sz = (int64_t)d.ns * (d.ns + 1) / 2 + (int64_t)(d.m - d.ns) * d.ns;
        pf += (sz + 15) & ~15LL;
        // contribution buffer: one dense row of length m per child; a child writes only the positions of its own update
        // rows, everything else stays at the zero it was allocated with
        o.U = ub;
        ub += (int64_t)o.nchild * d.m;
        o.pU = -1;
        o.ntask_f = o.ntask_b = 0;
        max_row = std::max(max_row, std::max(d.ns, d.m));
    }
    for (int m = 0; m < nmat; ++m) {
        const Symbolic& S = sym[m];
        for (int s = 0; s < S.nsuper; ++s)
            for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) {
                const int gp = sn_off[m] + s, gc = sn_off[m] + S.child_list[ci];
                ssn[gc].pU = ssn[gp].U + (int64_t)(ci - S.child_ptr[s]) * ssn[gp].m;
            }
    }
    pk_total = pf;
    // ---- chunks (one TMA copy each) and queue slots of <= SOLVE_GROUP_MAX chunks (the vector of a supernode is gathered once per slot) ----
    // tuning knobs (defaults measured on B200): stage size in doubles, ring depth, chunks per slot
    auto env_int = [](const char* name, int dflt) {
        const char* v = std::getenv(name);
        return v && *v ? std::atoi(v) : dflt;
    };
    // Stage size: the largest (multiple of 256 doubles, <= 3072) that still lets THREE CTAs share an SM's shared memory
    // next to the two vector buffers - measured on B200 (profiles/r1/solve_sweep_last.json): the third CTA is worth more than
    // bigger stages (bar136K_like: 0.58 ms with 3 CTAs x 20 KB stages vs 0.73 ms with 2 CTAs x 24 KB), and among 3-CTA
    // configurations bigger stages win (bar1M: 0.78 ms at 24 KB vs 0.85 ms at 20 KB).
    const int vec_bytes = 24 * (((max_front_all + 8 + 15) / 16) * 16);
    int auto_stage = ((73 * 1024 - vec_bytes) / 16) / 256 * 256;
    auto_stage = std::max(1536, std::min(3584, auto_stage));  // r2 sweep on bar1M: 3072 -> 0.714 ms, 3584 -> 0.676 ms, 3840 (2 CTAs/SM) -> 0.890 ms
    const int want_stage = env_int("DOTGPU_SOLVE_STAGE_DBL", auto_stage);
    solve_dbg = env_int("DOTGPU_SOLVE_DBG", 0);  // experiments only: 1 skip the products, 2 skip the TMA copies, 4 skip dependency waits
    solve_nstage = std::min(4, std::max(2, env_int("DOTGPU_SOLVE_NSTAGE", 2)));
    const int max_group = std::min(SOLVE_GROUP_MAX, std::max(1, env_int("DOTGPU_SOLVE_GROUP", SOLVE_GROUP_MAX)));
    stage_dbl = std::max((want_stage + 15) / 16 * 16, ((max_row + 2 + 15) / 16) * 16);
    vec_dbl = ((max_front_all + 8 + 15) / 16) * 16;
    solve_smem = (size_t)solve_nstage * stage_dbl * 8 + 2 * (size_t)vec_dbl * 12;
    DG_REQUIRE(solve_smem <= 200 * 1024, "front too large for the streamed solve");
    std::vector<SolveTask> chunks;
    auto fwd_off = [](const SolveSN& d, int r) -> int64_t {
        return r < d.ns ? (int64_t)r * (r + 1) / 2 : (int64_t)d.ns * (d.ns + 1) / 2 + (int64_t)(r - d.ns) * d.ns;
    };
    auto bwd_off = [](const SolveSN& d, int c) -> int64_t { return (int64_t)c * d.m - (int64_t)c * (c - 1) / 2; };
    const int max_rows_chunk = vec_dbl;  // idxbuf holds one entry per row of a chunk
    int group_chunks = 1;
    size_t group_start = 0;
    SolveTask null_chunk;
    std::memset(&null_chunk, 0, sizeof(null_chunk));
    auto emit = [&](int g, int kind) {
        SolveSN& d = ssn[g];
        const int nrows = kind ? d.ns : d.m;
        auto off_of = [&](int r) { return kind ? bwd_off(d, r) : fwd_off(d, r); };  // offset of row r = one past row r-1
        int cntt = 0, r0 = 0;
        while (r0 < nrows) {
            const int64_t o0 = off_of(r0);
            const int64_t a0 = o0 & ~1LL;  // 16-byte aligned start (panel bases are 128-byte aligned)
            int r1 = r0 + 1;
            while (r1 < nrows && r1 - r0 < max_rows_chunk && ((off_of(r1 + 1) + 1) & ~1LL) - a0 <= stage_dbl) ++r1;
            const int64_t e1 = (off_of(r1) + 1) & ~1LL;
            SolveTask T;
            T.src = (kind ? d.pb : d.pf) + a0;
            T.U = d.U;
            T.pU = d.pU;
            T.rows = d.rows;
            T.s = g;
            T.r0 = r0;
            T.r1 = r1;
            T.ndbl = (int)(e1 - a0);
            T.shift = (int)(o0 - a0);
            T.kind = kind;
            T.m = d.m;
            T.ns = d.ns;
            T.col0 = d.col0;
            T.nchild = d.nchild;
            T.dep_idx = T.dep_need = T.sig_idx = T.sig_total = 0;  // filled once the chunk counts are known
            T.sig_next = -1;
            T.g_r1 = 0;
            DG_REQUIRE(T.ndbl <= stage_dbl, "solve chunk exceeds the stage size");
            if (cntt % group_chunks == 0) {  // open a new queue slot: pad the previous one to SOLVE_GROUP_MAX descriptors
                while (chunks.size() % SOLVE_GROUP_MAX) chunks.push_back(null_chunk);
                group_start = chunks.size();
            }
            chunks.push_back(T);
            for (size_t c = group_start; c < chunks.size(); ++c) chunks[c].g_r1 = r1;  // last row of the group so far
            ++cntt;
            r0 = r1;
        }
        (kind ? d.ntask_b : d.ntask_f) = cntt;
    };
    int dev = 0, nsm = 0;
    DG_CUDA(cudaGetDevice(&dev));
    DG_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    // groups of up to SOLVE_GROUP_CHUNKS chunks where a level has far more chunks than CTAs, single chunks near the root
    auto level_pass = [&](int lv, int kind) {
        int64_t dbl = 0;
        for (int m = 0; m < nmat; ++m) {
            const Symbolic& S = sym[m];
            if (lv >= S.nlevels) continue;
            for (int i = S.level_ptr[lv]; i < S.level_ptr[lv + 1]; ++i) {
                const SolveSN& d = ssn[sn_off[m] + S.level_list[i]];
                dbl += (int64_t)d.ns * (d.ns + 1) / 2 + (int64_t)(d.m - d.ns) * d.ns;
            }
        }
        const int64_t nch = dbl / stage_dbl + 1;
        group_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(max_group, nch / (6LL * nsm)));
        for (int m = 0; m < nmat; ++m) {
            const Symbolic& S = sym[m];
            if (lv >= S.nlevels) continue;
            for (int i = S.level_ptr[lv]; i < S.level_ptr[lv + 1]; ++i) emit(sn_off[m] + S.level_list[i], kind);
        }
    };
    for (int lv = 0; lv < nlevels; ++lv) level_pass(lv, 0);
    for (int lv = nlevels - 1; lv >= 0; --lv) level_pass(lv, 1);
    while (chunks.size() % SOLVE_GROUP_MAX) chunks.push_back(null_chunk);
    n_solve_tasks = (int)(chunks.size() / SOLVE_GROUP_MAX);
    // dependency counters: every finished chunk bumps exactly one counter (fire-and-forget release reduction):
    //   forward chunk of s   -> the parent's "children chunks done" counter [2nsn + parent]   (a root: its own [s])
    //   backward chunk of s  -> its own [nsn + s]
    // and every chunk waits for one counter to reach one value:
    //   forward of s  : [2nsn + s] == sum of the children's forward chunk counts
    //   backward of s : [nsn + parent] == the parent's backward chunk count (which transitively waited for every forward phase
    //                   of the tree);  a root waits for its own forward phase [s] == ntask_f
    std::vector<int> child_chunks(nsn, 0);
    for (int g = 0; g < nsn; ++g)
        if (ssn[g].parent >= 0) child_chunks[ssn[g].parent] += ssn[g].ntask_f;
    for (SolveTask& T : chunks) {
        if (T.ndbl == 0) continue;
        const SolveSN& d = ssn[T.s];
        if (T.kind == 0) {
            T.dep_idx = 2 * nsn + T.s;
            T.dep_need = child_chunks[T.s];
            T.sig_idx = d.parent >= 0 ? 2 * nsn + d.parent : T.s;
        } else {
            if (d.parent >= 0) {
                T.dep_idx = nsn + d.parent;
                T.dep_need = ssn[d.parent].ntask_b;
            } else {
                T.dep_idx = T.s;
                T.dep_need = d.ntask_f;
            }
            T.sig_idx = nsn + T.s;
        }
        T.sig_total = 0;
        T.sig_next = -1;
    }
    // pack tasks, grouped by level (a level is packed as soon as its panels are final)
    std::vector<int> ptasks;
    pack_span.assign(nlevels, Span());
    for (int lv = 0; lv < nlevels; ++lv) {
        pack_span[lv].off = (int)ptasks.size();
        for (int g : level_sns[lv])
            for (int a = 0; a * 64 < sn[g].m; ++a)
                for (int b = 0; b * 64 < sn[g].ns; ++b)
                    if (a * 64 + 63 >= b * 64) {
                        ptasks.push_back(g);
                        ptasks.push_back((a & 0xffff) | (b << 16));
                    }
        pack_span[lv].cnt = ((int)ptasks.size() - pack_span[lv].off) / 2;
    }
    n_pack_tasks = (int)ptasks.size() / 2;
    if (ptasks.empty()) ptasks.push_back(0);
    if (chunks.empty()) chunks.assign(SOLVE_GROUP_MAX, null_chunk);
    d_ssn.upload(ssn, st);
    d_stasks.upload(chunks, st);
    d_ptasks.upload(ptasks, st);
    Pf.alloc(std::max<int64_t>(pk_total, 16));
    Pb.alloc(std::max<int64_t>(pk_total, 16));
    DG_CUDA(cudaMemsetAsync(Pf.p, 0, Pf.bytes(), st));  // alignment padding is streamed too: keep it finite
    DG_CUDA(cudaMemsetAsync(Pb.p, 0, Pb.bytes(), st));
    Ubuf.alloc(std::max<int64_t>(ub, 1));
    Ubuf.zero(st);
    d_cnt.alloc(3 * (size_t)std::max(nsn, 1) + 2);
    if (env_int("DOTGPU_SOLVE_TRACE", 0)) {
        d_trace.alloc((size_t)nsm * 4 * TRACE_CHUNKS * 8);
        d_trace.zero(st);
    }
    DG_CUDA(cudaFuncSetAttribute(k_solve_stream<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DG_CUDA(cudaFuncSetAttribute(k_solve_stream<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DG_CUDA(cudaFuncSetAttribute(k_solve_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = 0;
    if (solve_nstage == 2) DG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_stream<2>, SOLVE_THREADS, solve_smem));
    else if (solve_nstage == 3) DG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_stream<3>, SOLVE_THREADS, solve_smem));
    else DG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_stream<4>, SOLVE_THREADS, solve_smem));
    DG_REQUIRE(per_sm >= 1, "streamed solve kernel does not fit on an SM");
    solve_grid = std::min(nsm * per_sm, std::max(1, n_solve_tasks));
}

void CholBatch::pack_panels(int level, cudaStream_t st) {
    const Span& sp = pack_span[level];
    if (!sp.cnt) return;
    k_pack_panels<<<sp.cnt, 256, 0, st>>>((const Task3p*)(d_ptasks.p + sp.off), d_ssn.p, d_sn.p, Sp.p, Pf.p, Pb.p);
    count_launch();
}

void CholBatch::solve(const double* b, const int* gidx, double* x_perm, cudaStream_t st, const int* go) {
    if (!factorized) throw Error(DOTGPU_ERR_STATE, "solve before factorize");
    if (!n_solve_tasks) return;
    static const bool pre_gather = std::getenv("DOTGPU_SOLVE_NO_PREGATHER") == nullptr;
    if (gidx && pre_gather) {
        k_gather_rhs<<<ceil_div(n_total, 256), 256, 0, st>>>(n_total, b, gidx, rwork.p, go);
        count_launch();
        b = rwork.p;
        gidx = nullptr;
    }
    DG_CUDA(cudaMemsetAsync(d_cnt.p, 0, d_cnt.bytes(), st));
#define DG_SOLVE_LAUNCH(NS)                                                                                                         \
    k_solve_stream<NS><<<solve_grid, SOLVE_THREADS, solve_smem, st>>>(n_solve_tasks, d_stasks.p, d_rows.p, d_rel.p, Pf.p, Pb.p, b, \
                                                                       gidx, ywork.p, Ubuf.p, x_perm, d_cnt.p, stage_dbl, vec_dbl, solve_dbg, go, d_trace.p)
    if (solve_nstage == 2) DG_SOLVE_LAUNCH(2);
    else if (solve_nstage == 3) DG_SOLVE_LAUNCH(3);
    else DG_SOLVE_LAUNCH(4);
#undef DG_SOLVE_LAUNCH
    count_launch();
}

}  // namespace dotgpu
