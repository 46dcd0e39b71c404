// Host-side set-up: mesh features, domain decomposition from element labels, CSR-upper patterns
// and the gather lists the matrix-fill kernel consumes.  Mirrors (without std::map/std::set):
//   Mesh::computeFeatures / computeMassMatrix       Mesh.cpp:589-700, 552-585
//   LinSysSolver::set_pattern                        LinSysSolver/LinSysSolver.hpp:37-135
//   ADMMDDTimeStepper ctor + precompute (DD part)    TimeStepper/ADMMDDTimeStepper.cpp:155-278, 457-496
//   Mesh::constructSubmesh                           Mesh.cpp:855-905
//   DOTTimeStepper::computeHElemAndFillIn / fillInDecomposedHessians (as gather lists)
//                                                    TimeStepper/DOTTimeStepper.cpp:574-616, 619-797
#pragma once
#include <cstdint>
#include <vector>

namespace dotgpu {

void mesh_features(int nV, int nT, const double* V_rest, const int32_t* tets, double YM, double PR, double rho, double* DmInv,
                   double* vol, double* mass, double* mu, double* lam);

// sorted adjacency (vNeighbor) as CSR
void vertex_adjacency(int nV, int nT, const int32_t* tets, std::vector<int>& ptr, std::vector<int>& idx);

struct MatrixPattern {
    int nverts = 0;               // block rows; n = 3*nverts
    std::vector<int32_t> ia, ja;  // scalar CSR upper, 0-based
    std::vector<uint8_t> fixed;   // per block row
    // block view: for block row v, bcol[bptr[v]..bptr[v+1]) = block columns (first = v itself)
    std::vector<int32_t> bptr, bcol;
    int n() const { return 3 * nverts; }
    int64_t nnz() const { return (int64_t)ja.size(); }
};

// adjacency given as CSR (sorted, without self)
void build_pattern(int nverts, const std::vector<int>& adj_ptr, const std::vector<int>& adj_idx,
                   const std::vector<uint8_t>& fixed, MatrixPattern& out);

// Gather lists: for block b (in bptr order) sources src[ptr[b]..ptr[b+1]) are accumulated IN ORDER:
//   code >= 0 : 3x3 block `code` of the elemental Hessian array (code = 16*tet + 4*a + b)
//   code <  0 : consts[-code-1] * I
struct FillList {
    std::vector<int64_t> ptr;
    std::vector<int32_t> src;
    std::vector<double> consts;
};

struct SubdomainHost {
    std::vector<int32_t> elems;        // ascending global tet ids (METIS::getElementList)
    std::vector<int32_t> l2g;          // first-touch order (constructSubmesh)
    std::vector<int32_t> tets_local;   // 4*ne
    std::vector<int32_t> fixed_local;  // ascending local ids
    std::vector<double> mass_local;    // sub-mesh lumped mass
    std::vector<int32_t> iface;        // interface vertices (ascending global id) = keys of globalVIToDual
    MatrixPattern pat;
    FillList fill;
};

struct DDHost {
    int nV = 0, nT = 0, k = 0;
    std::vector<SubdomainHost> subs;
    std::vector<int32_t> dup;
    MatrixPattern gpat;
    FillList gfill;

    // V_rest/rho may be null/0 when only the combinatorial part is wanted (labels, patterns)
    void build(int nV, int nT, const int32_t* tets, const int32_t* epart, int k, const uint8_t* fixed_mask,
               const double* V_rest, double rho, const double* mass_global, bool with_fill, const std::vector<char>* sub_mask = nullptr);
    // LBFGS-JH: node blocks instead of element subdomains (block Jacobi of the global Hessian, LBFGSTimeStepper.cpp:241-262)
    void build_node_blocks(int nV, int nT, const int32_t* tets, const int32_t* npart, int k, const uint8_t* fixed_mask,
                           const double* mass_global);
};

// a14: element labels from the reference's vendored METIS (partition_host.cpp; dlopen of libdotmetis.so)
void metis_partition(int nV, int nT, const int32_t* tets, int k, int32_t* epart_out);
// METIS<3>::partMesh_nodes (METIS_PartMeshNodal with the same option vector, Utils/METIS.hpp:161-212): node labels [nV]
void metis_partition_nodes(int nV, int nT, const int32_t* tets, int k, int32_t* npart_out);
// multilevel vertex separator (METIS_ComputeVertexSeparator) for the fill-reducing ordering; false if libdotmetis.so is absent
bool metis_vertex_separator(int n, const int64_t* xadj, const int64_t* adjncy, int64_t* part);

}  // namespace dotgpu
