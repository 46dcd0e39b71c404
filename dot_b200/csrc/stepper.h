// Device-resident DOT time stepper (performance boundary): DOTTimeStepper::fullyImplicit /
// solve_oneStep / updateHessianAndFactor (TimeStepper/DOTTimeStepper.cpp:273-504) and the pieces of
// Optimizer it uses (initX, computeXTilta, lineSearch, initStepSize, the BE update of solve():
// TimeStepper/Optimizer.cpp:327-368, 442-610, 752-881, 1076-1093) with every vector on the device.
#pragma once
#include <deque>
#include <memory>

#include "chol_numeric.h"
#include "device_mesh.h"
#include "linalg.h"
#include "mesh_host.h"

namespace dotgpu {

struct Comm;  // NCCL communicator wrapper (comm.cpp)

// static round-robin map of subdomains to ranks (the fallback / bookkeeping map of dotgpu_owned_subdomains; a stepper balances
// its ranks by nnz(L) instead, see balanced_owner)
inline std::vector<int> owned_subdomains(int k, int rank, int world) {
    std::vector<int> o;
    for (int s = 0; s < k; ++s)
        if (s % world == rank) o.push_back(s);
    return o;
}
// whole subdomains are assigned to GPUs balancing sum nnz(L_s): longest-processing-time greedy on the given weights
// (deterministic: ties by lower rank / lower subdomain id, so every rank computes the same map).  Returns owner[s].
std::vector<int> balanced_owner(const std::vector<double>& weight, int world);

struct Stepper {
    dotgpu_stepper_config cfg;
    int nV = 0, nT = 0;
    cudaStream_t st = nullptr;
    std::vector<double> V_rest, mass_h, DmInv_h, vol_h, mu_h, lam_h;
    std::vector<int32_t> tets_h, epart_h;
    std::vector<uint8_t> fixed_h;
    DDHost dd;
    std::vector<int> owned;          // subdomain ids handled by this rank, ascending
    std::vector<int64_t> a_off;      // value offsets: [0]=global, [1+i]=owned[i]
    DeviceMesh mesh;                 // all tets (elemental Hessians of the refresh; energy / gradient on one GPU)
    DeviceMesh mesh_own;             // world > 1: the tets of the owned subdomains only - energy / gradient are sharded by tet
    DeviceMesh& emesh() { return cfg.world > 1 ? mesh_own : mesh; }
    std::vector<int> owner;          // world > 1: owner[s] = rank of subdomain s
    bool owner_locked = false;
    void setup_decomposition();      // everything that depends on the Dirichlet set
    void set_fixed(const uint8_t* fixed_mask, const double* x_eval);
    DevBuf<double> a_all;
    DevBuf<int> g_ia, g_ja;
    DeviceFill fill;
    CholBatch chol;
    DevBuf<int> gidx, cptr, cidx, dup;
    DevBuf<double> x, x0, xn, xt, vel, g, g_old, q, p, xperm, qf_partial, dot_partial, md_partial;
    HistList hist_list() const;
    std::vector<DevBuf<double>> S, Y;
    std::deque<int> hist;            // slots, oldest first
    DevBuf<double> sc;               // device scalars
    DevBuf<unsigned> counter;
    double* h_sc = nullptr;          // pinned mirror
    double* h_sc2 = nullptr;         // two pinned mirrors for the speculative pipeline (iteration parity)
    cudaEvent_t ev_it[2] = {nullptr, nullptr};
    DevBuf<int> spec_flags;          // device flags written by the last kernel of an iteration (linalg.h)
    int spec_enq = 0;
    double* h_x = nullptr;           // pinned staging for positions
    double target = 0.0, target_per_tolsq = 0.0;
    bool newton = false;             // DOTGPU_FLAG_NEWTON
    bool lbfgs_h = false, lbfgs_jh = false, unit_step = false;  // DOTGPU_FLAG_LBFGS_H / _JH; initStepSize = 1 (everything but DOT)
    std::vector<int32_t> npart_h;    // LBFGS-JH node labels
    bool debug_ls_fail = false;      // tests only: every trial energy reads as +inf (exercises the failed-line-search exit)
    double E_last = 0.0;
    std::vector<double> iter_log;    // (alpha, E, |g|^2) rows
    int64_t launches0 = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> pc_ev;  // event pairs around the preconditioner applications of a frame
    std::unique_ptr<Comm> comm;
    // world > 1: energy and gradient of the incremental potential at x_dev, summed over the ranks: G[0..3nV) = gradient,
    // G[3nV] = energy (one all-reduce of 3nV + 1 doubles; the inertia terms are added by rank 0 only)
    void eval_sharded(const double* x_dev, double* G);
    bool eval_sharded_push(const double* x_dev, double* G);  // first half of the peer-memory all-reduce only (stepper.cu)

    ~Stepper();
    void create(const dotgpu_stepper_config& c, int nV, int nT, const double* V_rest, const int32_t* tets, const int32_t* epart,
                const uint8_t* fixed_mask);
    void frame(double* x_inout, dotgpu_frame_stats* stats);
    void frame_resident(const int32_t* idx, const double* pos, int count, dotgpu_frame_stats* stats);
    void frame_core(dotgpu_frame_stats* stats, bool copy_back);
    DevBuf<int> h_idx;
    DevBuf<double> h_pos;
    void set_state(const double* x, const double* velocity);
    void get_state(double* x, double* velocity, double* xTilde);
    // fuse: optional inner products p . fuse->a[j] taken in the scatter pass (single-GPU only; returns false if not fused)
    bool precondition_dev(const double* q_dev, double* p_dev, const DotPairs* fuse = nullptr);
    void refresh();  // elemental Hessians at x, matrix fill, numeric factorisation
    double energy_at(const double* x_dev);
    void gradient_at(const double* x_dev, double* g_dev);
    void fetch_scalars(int first, int count);
    double compute_target() const;
    double time_kernels(int which, int reps);
};


}  // namespace dotgpu
