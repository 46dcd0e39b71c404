// Device-resident tet mesh (the static part of Mesh<3>, Mesh.hpp:39-60) and the per-tet kernels'
// launchers.  Layout in HBM:
//   tets     int4[nT]            one 16-B load per thread
//   DmInv    double[9][nT]       SoA -> coalesced across a warp of tets
//   vol,mu,lam double[nT]
//   lv_ptr/g_cptr/g_cidx/vp_*    K2 tables: CTA-local vertex lists and per-vertex partial lists (ascending tet order,
//                                the order of Mesh.cpp:606-611 / Energy.cpp:543-563 up to the association into CTA partials)
//   gpart    double[3][#(CTA,vertex)]  per-CTA vertex partial gradients, scratch
//   He       double[nT][10][9]   elemental Hessians: the 10 unique 3x3 blocks (k <= l) of the symmetric 4x4 block matrix,
//                                72 contiguous bytes per block, 720 B per tet
#pragma once
#include "common.h"

namespace dotgpu {

struct DeviceMesh {
    int nV = 0, nT = 0, energy = 0;
    DevBuf<int> tets;       // 4*nT
    DevBuf<double> DmInv;   // 9*nT SoA
    DevBuf<double> vol, mu, lam;
    DevBuf<double> mass;    // nV (may be empty for the bare energy object)
    DevBuf<unsigned char> fixed;  // nV
    // K2 tables (static): per CTA of 128 tets the distinct vertices and their (tet, corner) lists; per vertex its partials
    DevBuf<int> lv_ptr, vp_ptr, vp_idx;
    DevBuf<unsigned short> g_cptr, g_cidx;
    DevBuf<double> gpart;   // 3 doubles per (CTA, local vertex)
    DevBuf<double> epart;   // elastic energy partial per K2 CTA (energy fused into the gradient pass)
    DevBuf<double> He;      // 90*nT, allocated on first use
    DevBuf<double> partial; // block partial sums for reductions
    int n_partial = 0, nsm = 148;
    DevBuf<unsigned> counter;  // last-block detection of the fused energy reduction (self-resetting)

    void init(int energy_type, int nV_, int nT_, const int32_t* tets_h, const double* DmInv_rowmajor, const double* vol_h,
              const double* mu_h, const double* lam_h, const double* mass_h, const unsigned char* fixed_h, cudaStream_t st);
    void set_fixed(const unsigned char* fixed_h, cudaStream_t st);
};

// E_out[0] = coef * sum_t vol_t Psi_t(x)  + (xTilde ? sum_v m_v |x_v - xTilde_v|^2 / 2 : 0)    (device scalar)
void launch_energy(DeviceMesh& m, const double* x, const double* xTilde, double coef, double* E_out, cudaStream_t st);
void launch_energy_per_elem(DeviceMesh& m, const double* x, double* out, cudaStream_t st);
// g = gather(elemental gradients) [+ m (x - xTilde) on free vertices]; fixed entries zero
void launch_gradient(DeviceMesh& m, const double* x, const double* xTilde, double coef, double* g, cudaStream_t st);
// gradient + new L-BFGS pair + every inner product of the next iteration's first loop in one pass (see k_grad_vertex_pair);
// with_energy: also sc[SC_E] = incremental potential at x (K1 fused into K2: the line search needs both at the same point)
struct HistList;
struct PeerDst;
// several GPUs: energy + gradient of this rank's tets in one pass, stored as [g ; E] into this rank's slot of every rank's peer
// buffer (first half of the all-reduce, peer_reduce.cu); xTilde only on the rank that adds the inertia terms
void launch_gradient_push(DeviceMesh& m, const double* x, const double* xTilde, double coef, const PeerDst& D, cudaStream_t st);
void launch_gradient_pair(DeviceMesh& m, const double* x, const double* xTilde, double coef, double* g, const double* pdir,
                          const double* g_old, double* S_new, double* Y_new, int sl, const double* alpha_dev, double alpha_host,
                          const HistList& H, double* partial, unsigned* counter, double* sc, bool with_energy, cudaStream_t st,
                          int* flag_out = nullptr, double target = 0.0);
void launch_svd(DeviceMesh& m, const double* x, double* F, double* U, double* S, double* V, cudaStream_t st);
// fills m.He ([nT][10][9])
void launch_elem_hessians(DeviceMesh& m, const double* x, double coef, bool project, cudaStream_t st);
// converts m.He to the reference's row-major 12x12 layout
void launch_he_to_dense(DeviceMesh& m, double* out144, cudaStream_t st);

}  // namespace dotgpu
