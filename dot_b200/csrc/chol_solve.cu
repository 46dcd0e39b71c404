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
//   * a producer warp claims tasks from a global queue (topological order: forward levels up, then backward levels
//     down) and streams their byte ranges into a shared-memory ring with 1-D bulk TMA copies (cp.async.bulk +
//     mbarrier complete_tx).  The factor is read-only during a solve, so the stream runs AHEAD of the dependency
//     chain: when a supernode's inputs become ready its panel rows are already in shared memory;
//   * 8 consumer warps wait for the data (mbarrier) and for the task's dependencies (per-supernode completion
//     counters, acquire loads), gather the right-hand side / update vector, do the row.vector products out of shared
//     memory and publish the results (release);
//   * tasks are claimed in queue order, so every dependency of a claimed task was claimed earlier by a resident CTA:
//     no deadlock for any number of resident CTAs (several solver handles may run concurrently).
// All sums have a fixed order (lane-strided partial sums + shuffle tree, children in ascending order): results are
// bit-reproducible from run to run.
#include <algorithm>

#include "chol_numeric.h"

namespace dotgpu {

namespace {

constexpr int NSTAGE = 3;
constexpr int CONSUMERS = 256;
constexpr int SOLVE_THREADS = CONSUMERS + 32;

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
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
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
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }

__global__ void __launch_bounds__(SOLVE_THREADS, 3)
    k_solve_stream(const SolveTask* __restrict__ tasks, int ntasks, const SolveSN* __restrict__ sn, const int* __restrict__ ell,
                   const int* __restrict__ child, const int* __restrict__ rows, const double* __restrict__ Pf, const double* __restrict__ Pb,
                   const double* __restrict__ b, const int* __restrict__ gidx, double* y, double* uwork, double* x,
                   unsigned long long* cnt, unsigned* claim, int nsn, int stage_dbl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);          // NSTAGE * stage_dbl
    double* vec = ring + (size_t)NSTAGE * stage_dbl;              // right-hand side / update vector of the current task
    __shared__ __align__(8) unsigned long long full[NSTAGE], empty[NSTAGE];
    __shared__ SolveTask s_task[NSTAGE];
    __shared__ int s_tid[NSTAGE];

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= CONSUMERS) {
        // ------------------------------ producer warp: claim + stream ------------------------------
        if (tid == CONSUMERS) {
            for (int it = 0;; ++it) {
                const int st = it % NSTAGE;
                if (it >= NSTAGE) mbar_wait(&empty[st], ((it / NSTAGE) - 1) & 1);
                const int t = (int)atomicAdd(claim, 1u);
                s_tid[st] = t;
                if (t >= ntasks) {
                    mbar_arrive(&full[st]);
                    break;
                }
                const SolveTask T = tasks[t];
                s_task[st] = T;
                const unsigned bytes = (unsigned)T.ndbl * 8u;
                mbar_expect_tx(&full[st], bytes);
                tma_load_1d(ring + (size_t)st * stage_dbl, (T.kind ? Pb : Pf) + T.src, bytes, &full[st]);
            }
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    const int warp = tid >> 5, lane = tid & 31;
    int verified_s = -1, verified_kind = -1;
    for (int it = 0;; ++it) {
        const int st = it % NSTAGE;
        mbar_wait(&full[st], (it / NSTAGE) & 1);
        if (s_tid[st] >= ntasks) break;
        const SolveTask T = s_task[st];
        const SolveSN d = sn[T.s];
        const double* __restrict__ chunk = ring + (size_t)st * stage_dbl + T.shift;
        // ---- dependencies ----
        if (!(verified_s == T.s && verified_kind == T.kind)) {
            if (tid == 0) {
                if (T.kind == 0) {
                    for (int ci = 0; ci < d.nchild; ++ci) {
                        const int c = child[d.child_begin + ci];
                        const unsigned long long need = (unsigned long long)sn[c].ntask_f;
                        while (ld_acquire(cnt + c) < need) {
                        }
                    }
                } else if (d.parent >= 0) {
                    const unsigned long long need = (unsigned long long)sn[d.parent].ntask_b;
                    while (ld_acquire(cnt + nsn + d.parent) < need) {
                    }
                } else {
                    const unsigned long long need = (unsigned long long)d.ntask_f;
                    while (ld_acquire(cnt + T.s) < need) {
                    }
                }
            }
            consumer_sync();
            verified_s = T.s;
            verified_kind = T.kind;
        }
        if (T.kind == 0) {
            // ================= forward: [y_s ; u_s](rows r0..r1) = S_s(rows) * t,  t = b_s + children updates =================
            const int tn = min(d.ns, T.r1);
            const int* __restrict__ esrc = ell + d.ell;
            for (int k = tid; k < tn; k += CONSUMERS) {
                double v = gidx ? b[gidx[d.col0 + k]] : b[d.col0 + k];
                for (int w = 0; w < d.nchild; ++w) {
                    const int src = esrc[(long long)k * d.nchild + w];
                    if (src >= 0) v += __ldcg(uwork + src);
                }
                vec[k] = v;
            }
            consumer_sync();
            const long long tri = (long long)d.ns * (d.ns + 1) / 2;
            const long long base0 = T.r0 < d.ns ? (long long)T.r0 * (T.r0 + 1) / 2 : tri + (long long)(T.r0 - d.ns) * d.ns;
            for (int r = T.r0 + warp; r < T.r1; r += CONSUMERS / 32) {
                const long long off = (r < d.ns ? (long long)r * (r + 1) / 2 : tri + (long long)(r - d.ns) * d.ns) - base0;
                const int kn = r < d.ns ? r + 1 : d.ns;
                const double* __restrict__ row = chunk + off;
                double sum = 0.0;
                for (int k = lane; k < kn; k += 32) sum += row[k] * vec[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
                if (lane == 0) {
                    if (r < d.ns) {
                        y[d.col0 + r] = sum;
                    } else {
                        double v = 0.0;
                        for (int w = 0; w < d.nchild; ++w) {
                            const int src = esrc[(long long)r * d.nchild + w];
                            if (src >= 0) v += __ldcg(uwork + src);
                        }
                        uwork[d.u + (r - d.ns)] = v + sum;
                    }
                }
            }
        } else {
            // ================= backward: x_s(cols r0..r1) = S_s^T(cols) * [y_s ; x(rows below)] =================
            const int c0 = T.r0;
            const int* __restrict__ frows = rows + d.rows;
            for (int r = c0 + tid; r < d.m; r += CONSUMERS) vec[r - c0] = r < d.ns ? __ldcg(y + d.col0 + r) : __ldcg(x + frows[r]);
            consumer_sync();
            // packed row c of S^T starts at c*m - c(c-1)/2 and holds rows c..m-1
            const long long base0 = (long long)c0 * d.m - (long long)c0 * (c0 - 1) / 2;
            for (int c = c0 + warp; c < T.r1; c += CONSUMERS / 32) {
                const long long off = (long long)c * d.m - (long long)c * (c - 1) / 2 - base0;
                const double* __restrict__ row = chunk + off;
                const double* __restrict__ z = vec + (c - c0);
                const int len = d.m - c;
                double sum = 0.0;
                for (int k = lane; k < len; k += 32) sum += row[k] * z[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
                if (lane == 0) x[d.col0 + c] = sum;
            }
        }
        consumer_sync();  // all results stored, all shared-memory reads of this stage done
        if (tid == 0) {
            mbar_arrive(&empty[st]);
            __threadfence();
            atomicAdd(cnt + (T.kind ? nsn : 0) + T.s, 1ull);
        }
    }
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
    std::vector<int> ell;
    int64_t pf = 0;
    int max_row = 2;
    // ELL table of update sources: ell[d.ell + i*nchild + w] = uwork index that child w adds to front row i (or -1)
    std::vector<int> child_slot(nsn, 0);  // position of a supernode among its parent's children
    for (int m = 0; m < nmat; ++m) {
        const Symbolic& S = sym[m];
        for (int s = 0; s < S.nsuper; ++s)
            for (int ci = S.child_ptr[s]; ci < S.child_ptr[s + 1]; ++ci) child_slot[sn_off[m] + S.child_list[ci]] = ci - S.child_ptr[s];
    }
    for (int g = 0; g < nsn; ++g) {
        const SNDesc& d = sn[g];
        SolveSN& o = ssn[g];
        o.m = d.m;
        o.ns = d.ns;
        o.col0 = d.col0;
        o.u = d.u;
        o.rows = d.rows;
        o.parent = d.parent;
        o.nchild = d.child_end - d.child_begin;
        o.child_begin = d.child_begin;
        o.pf = o.pb = pf;
        const int64_t sz = (int64_t)d.ns * (d.ns + 1) / 2 + (int64_t)(d.m - d.ns) * d.ns;
        pf += (sz + 15) & ~15LL;
        o.ell = (int64_t)ell.size();
        ell.resize(ell.size() + (size_t)d.m * o.nchild, -1);
        max_row = std::max(max_row, std::max(d.ns, d.m));
    }
    pk_total = pf;
    // fill the ELL table from the symbolic relative indices
    for (int m = 0; m < nmat; ++m) {
        const Symbolic& S = sym[m];
        for (int s = 0; s < S.nsuper; ++s) {
            const int p = S.parent[s];
            if (p < 0) continue;
            const int g = sn_off[m] + s, gp = sn_off[m] + p;
            const int ns = S.nscol(s);
            for (int64_t o = S.row_ptr[s] + ns; o < S.row_ptr[s + 1]; ++o) {
                const int64_t src = sn[g].u + (o - S.row_ptr[s] - ns);
                DG_REQUIRE(src < (1LL << 31), "update workspace too large for 32-bit indices");
                ell[ssn[gp].ell + (int64_t)S.rel[o] * ssn[gp].nchild + child_slot[g]] = (int)src;
            }
        }
    }
    // ---- tasks: forward levels ascending, backward levels descending ----
    stage_dbl = std::max(2048, ((max_row + 2 + 15) / 16) * 16);
    DG_REQUIRE((size_t)NSTAGE * stage_dbl * 8 + (size_t)(max_front_all + 8) * 8 <= 200 * 1024, "front too large for the streamed solve");
    std::vector<SolveTask> tasks;
    auto fwd_off = [](const SolveSN& d, int r) -> int64_t {
        return r < d.ns ? (int64_t)r * (r + 1) / 2 : (int64_t)d.ns * (d.ns + 1) / 2 + (int64_t)(r - d.ns) * d.ns;
    };
    auto bwd_off = [](const SolveSN& d, int c) -> int64_t { return (int64_t)c * d.m - (int64_t)c * (c - 1) / 2; };
    auto emit = [&](int g, int kind) {
        SolveSN& d = ssn[g];
        const int nrows = kind ? d.ns : d.m;
        int cntt = 0;
        int r0 = 0;
        while (r0 < nrows) {
            const int64_t o0 = kind ? bwd_off(d, r0) : fwd_off(d, r0);
            const int64_t a0 = o0 & ~1LL;  // 16-byte aligned start (panel bases are 128-byte aligned)
            int r1 = r0 + 1;
            auto end_of = [&](int r) { return kind ? bwd_off(d, r) : fwd_off(d, r); };  // offset one past row r-1
            while (r1 < nrows && ((end_of(r1 + 1) + 1) & ~1LL) - a0 <= stage_dbl) ++r1;
            const int64_t e1 = (end_of(r1) + 1) & ~1LL;
            SolveTask T;
            T.src = (kind ? d.pb : d.pf) + a0;
            T.s = g;
            T.r0 = r0;
            T.r1 = r1;
            T.ndbl = (int)(e1 - a0);
            T.shift = (int)(o0 - a0);
            T.kind = kind;
            DG_REQUIRE(T.ndbl <= stage_dbl, "solve task exceeds the stage size");
            tasks.push_back(T);
            ++cntt;
            r0 = r1;
        }
        (kind ? d.ntask_b : d.ntask_f) = cntt;
    };
    for (int lv = 0; lv < nlevels; ++lv)
        for (int m = 0; m < nmat; ++m) {
            const Symbolic& S = sym[m];
            if (lv >= S.nlevels) continue;
            for (int i = S.level_ptr[lv]; i < S.level_ptr[lv + 1]; ++i) emit(sn_off[m] + S.level_list[i], 0);
        }
    for (int lv = nlevels - 1; lv >= 0; --lv)
        for (int m = 0; m < nmat; ++m) {
            const Symbolic& S = sym[m];
            if (lv >= S.nlevels) continue;
            for (int i = S.level_ptr[lv]; i < S.level_ptr[lv + 1]; ++i) emit(sn_off[m] + S.level_list[i], 1);
        }
    n_solve_tasks = (int)tasks.size();
    // pack tasks
    std::vector<int> ptasks;
    for (int g = 0; g < nsn; ++g)
        for (int a = 0; a * 64 < sn[g].m; ++a)
            for (int b = 0; b * 64 < sn[g].ns; ++b)
                if (a * 64 + 63 >= b * 64) {
                    ptasks.push_back(g);
                    ptasks.push_back((a & 0xffff) | (b << 16));
                }
    n_pack_tasks = (int)ptasks.size() / 2;
    if (ptasks.empty()) ptasks.push_back(0);
    if (tasks.empty()) tasks.push_back(SolveTask());
    if (ell.empty()) ell.push_back(-1);
    d_ssn.upload(ssn, st);
    d_ell.upload(ell, st);
    d_stasks.upload(tasks, st);
    d_ptasks.upload(ptasks, st);
    Pf.alloc(std::max<int64_t>(pk_total, 16));
    Pb.alloc(std::max<int64_t>(pk_total, 16));
    DG_CUDA(cudaMemsetAsync(Pf.p, 0, Pf.bytes(), st));  // alignment padding is streamed too: keep it finite
    DG_CUDA(cudaMemsetAsync(Pb.p, 0, Pb.bytes(), st));
    d_cnt.alloc(2 * (size_t)std::max(nsn, 1) + 2);
    solve_smem = (size_t)NSTAGE * stage_dbl * 8 + (size_t)(max_front_all + 8) * 8;
    DG_CUDA(cudaFuncSetAttribute(k_solve_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int dev = 0, nsm = 0, per_sm = 0;
    DG_CUDA(cudaGetDevice(&dev));
    DG_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    DG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_stream, SOLVE_THREADS, solve_smem));
    DG_REQUIRE(per_sm >= 1, "streamed solve kernel does not fit on an SM");
    solve_grid = std::min(nsm * per_sm, std::max(1, n_solve_tasks));
}

void CholBatch::pack_panels(cudaStream_t st) {
    if (!n_pack_tasks) return;
    k_pack_panels<<<n_pack_tasks, 256, 0, st>>>((const Task3p*)d_ptasks.p, d_ssn.p, d_sn.p, Sp.p, Pf.p, Pb.p);
    count_launch();
}

void CholBatch::solve(const double* b, const int* gidx, double* x_perm, cudaStream_t st) {
    if (!factorized) throw Error(DOTGPU_ERR_STATE, "solve before factorize");
    if (!n_solve_tasks) return;
    DG_CUDA(cudaMemsetAsync(d_cnt.p, 0, d_cnt.bytes(), st));
    unsigned* claim = reinterpret_cast<unsigned*>(d_cnt.p + 2 * (size_t)std::max(nsuper_total, 1));
    k_solve_stream<<<solve_grid, SOLVE_THREADS, solve_smem, st>>>(d_stasks.p, n_solve_tasks, d_ssn.p, d_ell.p, d_child.p, d_rows.p, Pf.p, Pb.p, b, gidx,
                                                                   ywork.p, uwork.p, x_perm, d_cnt.p, claim, nsuper_total, stage_dbl);
    count_launch();
}

}  // namespace dotgpu
