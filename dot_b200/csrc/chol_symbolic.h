// Symbolic analysis for the supernodal multifrontal Cholesky that replaces CHOLMOD
// (cholmod_analyze via CHOLMODSolver::analyze_pattern, LinSysSolver/CHOLMODSolver.cpp:136-141;
// SURVEY.md Appendix C).  CHOLMOD: AMD ordering + etree + relaxed supernodes, tuned for a CPU
// left-looking factorisation.  Here: level-structure nested dissection on the 3x3-block graph,
// every separator / leaf domain is one dense supernode, which gives a short, wide elimination
// tree (few dependent levels, fat fronts) - what the level-scheduled GPU kernels want.
#pragma once
#include <cstdint>
#include <vector>

namespace dotgpu {

struct Symbolic {
    int n = 0;
    std::vector<int32_t> perm, iperm;   // perm[new] = old, iperm[old] = new
    int nsuper = 0;
    std::vector<int32_t> super_ptr;     // [nsuper+1] first column of each supernode (permuted numbering)
    std::vector<int64_t> row_ptr;       // [nsuper+1]
    std::vector<int32_t> rows;          // front row indices, ascending; first nscol = own columns
    std::vector<int32_t> rel;           // same shape as rows: position of a below-row in the PARENT front (-1 for own cols / root)
    std::vector<int32_t> parent;        // supernodal etree
    std::vector<int32_t> level;         // 0 = leaves
    int nlevels = 0;
    std::vector<int32_t> level_ptr, level_list;  // supernodes grouped by level
    std::vector<int32_t> child_ptr, child_list;  // children of each supernode, ascending
    std::vector<int64_t> panel_off;     // [nsuper+1] offsets of the m x nscol row-major panels
    std::vector<int64_t> cb_off;        // [nsuper+1] offsets of the nb x nb contribution blocks
    std::vector<int64_t> u_off;         // [nsuper+1] offsets of the nb update vectors used by the solves
    std::vector<int64_t> amap;          // [nnz(A)] position in panel storage of every CSR-upper entry
    // extend-add gather lists for the forward solve: for front row r of supernode s (global row index
    // row_ptr[s]+r) the children update-vector entries ea_src[ea_ptr[..]..) that land on it
    std::vector<int64_t> ea_ptr;
    std::vector<int64_t> ea_src;
    int64_t nnz_l = 0;
    double flops = 0.0;
    int max_front = 0, max_nscol = 0;

    int nscol(int s) const { return super_ptr[s + 1] - super_ptr[s]; }
    int front(int s) const { return (int)(row_ptr[s + 1] - row_ptr[s]); }

    // ia/ja: CSR upper, 0-based (LinSysSolver::set_pattern layout). leaf_nodes: max graph nodes per leaf domain.
    void analyze(int n, const int32_t* ia, const int32_t* ja, int leaf_nodes = 21);
};

}  // namespace dotgpu
