// Batched supernodal multifrontal Cholesky on the device: numeric factorisation and the
// forward/backward solves for MANY matrices (subdomains) at once, level-scheduled over the merged
// elimination forests.  Replaces cholmod_factorize / cholmod_solve per subdomain
// (LinSysSolver/CHOLMODSolver.cpp:143-163; SuiteSparse/CHOLMOD/Supernodal/t_cholmod_super_numeric.c,
// t_cholmod_super_solve.c).  Dense tile GEMMs run on the fp64 tensor pipe (DMMA).
#pragma once
#include <memory>

#include "chol_symbolic.h"
#include "common.h"

namespace dotgpu {

constexpr int CH_NB = 64;  // pivot tile width

struct SNDesc {
    long long panel, cb, u, rows, tinv;
    int m, ns, col0, child_begin, child_end, parent;
};

// streamed solve (chol_solve.cu): per-supernode descriptor and task (one contiguous byte range of a packed panel)
struct SolveSN {
    long long pf, pb;   // offsets (doubles) of the packed forward / backward panels (same size, 128-byte aligned)
    long long U, pU;    // own contribution buffer [nchild][m] (children write into it); where this supernode writes in its parent's
    long long rows;
    int m, ns, col0, parent, nchild, child_begin, ntask_f, ntask_b;
};
// One TMA copy + everything the consumers need to process it (no further descriptor loads on the critical path).
struct SolveTask {
    long long src;      // 16-byte aligned offset (doubles) into Pf (kind 0) or Pb (kind 1)
    long long U, pU;    // contribution buffers: own (read) / slot in the parent's (written)
    long long rows;     // offset of the supernode's front rows in the row / relative-index arrays
    int s, r0, r1;      // supernode, row range (forward: front rows, backward: columns)
    int ndbl, shift;    // doubles to stream, position of row r0 inside the streamed range
    int kind;           // 0 forward, 1 backward
    int m, ns, col0, nchild;
    int dep_idx, dep_need;             // wait until counter[dep_idx] >= dep_need (dep_need 0: nothing to wait for)
    int sig_idx, sig_total, sig_next;  // bump counter[sig_idx]; the bump that makes it sig_total also bumps counter[sig_next] (if >= 0)
    int g_r1;           // end of the row range covered by the chunk's group
};
static_assert(sizeof(SolveTask) == 96, "descriptor copies assume 96 bytes");

struct CholBatch {
    int nmat = 0;
    std::vector<Symbolic> sym;            // host symbolic per matrix
    std::vector<int64_t> col_off;         // [nmat+1] offsets into the concatenated permuted vectors
    std::vector<int64_t> nnz_off;         // [nmat+1] offsets into the concatenated CSR values
    std::vector<int> sn_off;              // [nmat+1] global supernode ids
    int64_t n_total = 0, nnz_a_total = 0, nnz_l_total = 0, cb_total = 0, u_total = 0, tinv_total = 0;
    int nsuper_total = 0, nlevels = 0;
    double flops_total = 0.0;

    // device
    DevBuf<SNDesc> d_sn;
    DevBuf<int> d_rows, d_rel, d_child;
    DevBuf<long long> d_amap, d_ea_ptr, d_ea_src;
    DevBuf<double> L, Sp, CB, tinv, ywork, xwork, uwork, rwork;
    DevBuf<int> d_status;
    DevBuf<int> d_tasks;

    struct Span { int off = 0, cnt = 0; };
    struct LevelPlan {
        Span extend, fwd, bwd, update_cb;
        int extend_split_m = 0x7fffffff;        // extend-add: fronts above this many rows are split into column slabs
        std::vector<Span> potrf, trsm, update;  // per pivot step
        std::vector<Span> sp_diag, sp_triinv;   // per pivot step: solve-panel tiles that become computable after potrf(kb)
        Span sp_below, pack;                    // after the level's panels are final
    };
    std::vector<LevelPlan> plan;
    int max_front_all = 0;
    bool factorized = false;
    // streamed solve
    DevBuf<SolveSN> d_ssn;
    DevBuf<SolveTask> d_stasks;
    DevBuf<int> d_ptasks;
    DevBuf<double> Pf, Pb, Ubuf;
    DevBuf<unsigned> d_cnt;
    DevBuf<unsigned long long> d_trace;   // DOTGPU_SOLVE_TRACE=1: per-chunk timestamps of the last solve (experiments)
    int64_t pk_total = 0;
    int n_solve_tasks = 0, n_pack_tasks = 0, stage_dbl = 0, vec_dbl = 0, solve_grid = 0, solve_nstage = 3, solve_dbg = 0;
    size_t solve_smem = 0;
    void build_solve_plan(const std::vector<SNDesc>& sn, cudaStream_t st);
    // second stream + events: the solve panels of a supernode are built while the pivot chain of its level / the upper levels runs
    cudaStream_t st2 = nullptr;
    std::vector<cudaEvent_t> evs;
    cudaEvent_t ev_join = nullptr;
    ~CholBatch();
    void pack_panels(int level, cudaStream_t st);
    std::vector<std::vector<int>> level_sns;   // supernodes (batch ids) per level
    std::vector<Span> pack_span;               // per level, into d_ptasks

    // ia/ja per matrix: CSR upper patterns.  Builds symbolic + device structures.
    void analyze(const std::vector<const int32_t*>& ia, const std::vector<const int32_t*>& ja, const std::vector<int>& n,
                 int leaf_nodes, cudaStream_t st);
    // a_all: device pointer to the concatenated CSR values (pattern order) of all matrices of the batch
    void factorize(const double* a_all, cudaStream_t st);
    void enqueue_factorize(const double* a_all, cudaStream_t st);  // the plain launch sequence (captured into a graph by factorize)
    cudaGraphExec_t graph_exec = nullptr;
    const double* graph_a = nullptr;
    cudaStream_t graph_st = nullptr;
    int64_t graph_launches = 0;
    // throws Error(DOTGPU_ERR_NOT_SPD) if the last factorize met a non-positive pivot (syncs the stream)
    void check_status(cudaStream_t st);
    // x_perm: device vector of n_total doubles in the PERMUTED order of each matrix (x_perm[col_off[m] + i] = x_m[perm_m[i]]).
    // Right-hand side: entry i of the permuted concatenation is b[gidx[i]] (gather fused into the solve), or b[i] when
    // gidx == nullptr (then b is already permuted and b == x_perm is allowed).
    // go: optional device flag; *go == 0 turns the whole solve into a no-op (speculatively enqueued L-BFGS iterations, linalg.h)
    void solve(const double* b, const int* gidx, double* x_perm, cudaStream_t st, const int* go = nullptr);
    // level-by-level reference implementation of the same solve (one launch per level and direction); kept as the
    // in-library cross-check of the streamed kernel
    void solve_levels(const double* b_perm, double* x_perm, cudaStream_t st);
    int64_t device_bytes() const;
};

}  // namespace dotgpu
