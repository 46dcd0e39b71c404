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
        Span extend, fwd, bwd;
        std::vector<Span> potrf, trsm, update;  // per pivot step
    };
    std::vector<LevelPlan> plan;
    Span sp_diag, sp_below;
    std::vector<Span> sp_triinv;
    int max_front_all = 0;
    bool factorized = false;

    // ia/ja per matrix: CSR upper patterns.  Builds symbolic + device structures.
    void analyze(const std::vector<const int32_t*>& ia, const std::vector<const int32_t*>& ja, const std::vector<int>& n,
                 int leaf_nodes, cudaStream_t st);
    // a_all: device pointer to the concatenated CSR values (pattern order) of all matrices of the batch
    void factorize(const double* a_all, cudaStream_t st);
    // throws Error(DOTGPU_ERR_NOT_SPD) if the last factorize met a non-positive pivot (syncs the stream)
    void check_status(cudaStream_t st);
    // b_perm / x_perm: device vectors of n_total doubles in the PERMUTED order of each matrix
    // (b_perm[col_off[m] + i] = b_m[perm_m[i]]); in place allowed (b_perm == x_perm)
    void solve(const double* b_perm, double* x_perm, cudaStream_t st);
    int64_t device_bytes() const;
};

}  // namespace dotgpu
