/* libdotgpu - C ABI of the B200-native DOT hot path.
 *
 * The reference (penn-graphics-research/DOT) has no FFI layer: its substitution points are two
 * C++ class hierarchies chosen at compile time.  Every entry point
 * below states which reference interface it sits under (paths relative to the reference's
 * src/).  INTEGRATION.md shows the C++ subclasses (GpuEnergy : Energy<3>, the shadow
 * CHOLMODSolver.hpp -> GpuCholSolver : LinSysSolver, GpuDOTStepper : Optimizer<3>) that bind them.
 *
 * Conventions: plain pointers and sizes; all host buffers are caller-owned; fp64 values, int32
 * indices; vectors of 3*nV doubles are xyz-interleaved (the layout of Eigen::VectorXd gradient /
 * searchDir in the reference); dense per-tet 3x3 / 12x12 outputs are row-major.  Every function
 * returns DOTGPU_OK (0) or a negative error code, never throws, never exits; dotgpu_last_error()
 * gives a message for the calling thread.  A handle is not re-entrant; different handles may be
 * used from different threads.  There is no CPU fallback: without a CUDA device every create
 * call fails with DOTGPU_ERR_NO_DEVICE.
 */
#ifndef DOTGPU_H
#define DOTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOTGPU_OK 0
#define DOTGPU_ERR_INVALID -1     /* bad argument */
#define DOTGPU_ERR_NO_DEVICE -2   /* no CUDA device / CUDA runtime failure at start-up */
#define DOTGPU_ERR_CUDA -3        /* a CUDA call or kernel failed */
#define DOTGPU_ERR_NOT_SPD -4     /* factorize met a non-positive pivot */
#define DOTGPU_ERR_STATE -5       /* call order violated (e.g. solve before factorize) */
#define DOTGPU_ERR_NCCL -6

#define DOTGPU_ENERGY_FCR 0 /* Fixed CoRotational : Energy/Physics_Elasticity/FixedCoRotEnergy.cpp */
#define DOTGPU_ENERGY_SNH 1 /* Stable Neo-Hookean : Energy/Physics_Elasticity/StableNHEnergy.cpp (SNH_WITHLOG off) */

/* AnimScripter kinds (AnimScripter.cpp:29-453) */
#define DOTGPU_ANIM_NULL 0
#define DOTGPU_ANIM_STRETCH 1
#define DOTGPU_ANIM_SQUASH 2
#define DOTGPU_ANIM_STRETCHNSQUASH 3
#define DOTGPU_ANIM_TWIST 4
#define DOTGPU_ANIM_TWISTNSTRETCH 5
#define DOTGPU_ANIM_TWISTNSNS 6
#define DOTGPU_ANIM_TWISTNSNS_OLD 7
#define DOTGPU_ANIM_RUBBERBANDPULL 8 /* changes the Dirichlet set mid-run (AnimScripter.cpp:219-257, 404-423) */

const char* dotgpu_last_error(void);
int dotgpu_version(void);
/* number of visible CUDA devices (0 if none); never fails */
int dotgpu_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Host-side mesh precomputation (Mesh.cpp:589-700 computeFeatures, :552-585 mass, :741-744).
 * DmInv_out [nT*9] row-major restTriInv, vol_out [nT] signed rest volume (triArea),
 * mass_out [nV] lumped mass (rho*|vol|/4 per corner), mu_out/lambda_out [nT].
 * ------------------------------------------------------------------------------------------ */
int dotgpu_mesh_features(int nV, int nT, const double* V_rest, const int32_t* tets, double YM, double PR, double rho,
                         double* DmInv_out, double* vol_out, double* mass_out, double* mu_out, double* lambda_out);

/* ------------------------------------------------------------------------------------------
 * Energy object: sits under Energy<3> (Energy/Energy.hpp:27-226).
 * ------------------------------------------------------------------------------------------ */
typedef struct dotgpu_energy dotgpu_energy;

/* Uploads the static per-tet data a Mesh<3> holds (Mesh.hpp:39-60): tets [nT*4], restTriInv
 * [nT*9] row-major, triArea [nT], u/lambda [nT], isFixedVert [nV] (0/1). */
int dotgpu_energy_create(dotgpu_energy** out, int device, int energy_type, int nV, int nT, const int32_t* tets,
                         const double* DmInv, const double* vol, const double* mu, const double* lambda,
                         const uint8_t* fixed_mask);
void dotgpu_energy_destroy(dotgpu_energy* e);
int dotgpu_energy_set_fixed(dotgpu_energy* e, const uint8_t* fixed_mask);

/* Energy::computeEnergyVal (Energy.hpp:57-64 -> Energy.cpp:426, 294-423): E = coef * sum_t vol_t Psi_t(x).
 * x [nV*3] host.  redoSVD has no effect on the value (the device path always evaluates at x). */
int dotgpu_energy_value(dotgpu_energy* e, const double* x, double coef, double* E_out);
/* per-element Psi_t*vol_t (Energy::getEnergyValPerElemBySVD, Energy.cpp:294-423), out [nT] */
int dotgpu_energy_per_elem(dotgpu_energy* e, const double* x, double* out);
/* Energy::computeGradient (Energy.hpp:65-72 -> Energy.cpp:441-564): g [nV*3], fixed entries zeroed. */
int dotgpu_energy_gradient(dotgpu_energy* e, const double* x, double coef, double* g_out);
/* deformation gradient + SVD with the reference's conventions (IglUtils::computeSVD_SIMD,
 * IglUtils.cpp:929-1085): F,U,V [nT*9] row-major, Sigma [nT*3]; any output may be NULL. */
int dotgpu_energy_svd(dotgpu_energy* e, const double* x, double* F_out, double* U_out, double* Sigma_out, double* V_out);
/* Energy::computeElemHessianByPK (Energy.hpp:133-140 -> Energy.cpp:673-777, 1129-1270):
 * He_out [nT*144] row-major 12x12 (w = coef*vol), vInds_out [nT*4] = v or -v-1 if fixed (may be NULL). */
int dotgpu_energy_elem_hessians(dotgpu_energy* e, const double* x, double coef, int projectSPD, double* He_out,
                                int32_t* vInds_out);

/* ------------------------------------------------------------------------------------------
 * Sparse SPD solver: sits under LinSysSolver<VectorXi,VectorXd> / replaces CHOLMODSolver
 * (LinSysSolver/LinSysSolver.hpp:22-436, CHOLMODSolver.cpp).  Pattern = upper-triangular CSR
 * (row-major, 0-based, ascending columns, diagonal present in every row) exactly as
 * LinSysSolver::set_pattern builds it (LinSysSolver.hpp:37-135; = CSC lower for CHOLMOD stype=-1).
 * ------------------------------------------------------------------------------------------ */
typedef struct dotgpu_solver dotgpu_solver;

/* set_pattern + analyze_pattern (cholmod_analyze, CHOLMODSolver.cpp:88-106, 136-141): fill-reducing
 * ordering, supernodal symbolic factorisation, device allocation. */
int dotgpu_solver_create(dotgpu_solver** out, int device, int n, const int32_t* ia, const int32_t* ja);
void dotgpu_solver_destroy(dotgpu_solver* s);
/* update_a / setCoeff / addCoeff end state: the nnz values in pattern order */
int dotgpu_solver_set_values(dotgpu_solver* s, const double* a);
/* factorize (CHOLMODSolver.cpp:143-146). Returns DOTGPU_OK on success (NOT the reference's inverted bool). */
int dotgpu_solver_factorize(dotgpu_solver* s);
/* solve (CHOLMODSolver.cpp:149-163): x = A^-1 rhs, both [n] host */
int dotgpu_solver_solve(dotgpu_solver* s, const double* rhs, double* x);
/* multiply (CHOLMODSolver.cpp:185-208, cholmod_sdmult with stype=-1): y = A x with A symmetric */
int dotgpu_solver_multiply(dotgpu_solver* s, const double* x, double* y);
/* introspection (what SURVEY.md App. A.10 reads from cholmod_factor): */
typedef struct dotgpu_solver_info {
    int32_t n, nsuper, nlevels, max_front, max_nscol;
    int64_t nnz_a, nnz_l;     /* nnz of A's stored triangle; nnz of the supernodal factor L (packed triangles + sub-diagonal blocks) */
    double flops;             /* factorisation flops */
    int64_t device_bytes;
} dotgpu_solver_info;
int dotgpu_solver_get_info(dotgpu_solver* s, dotgpu_solver_info* info);
/* symbolic structure for checkers: perm [n] (new->old); super_ptr [nsuper+1] column ranges in the
 * permuted order; row_ptr [nsuper+1] / rows [row_ptr[nsuper]] = front row indices (permuted order,
 * ascending, the first nscol being the supernode's own columns); parent [nsuper]; level [nsuper].
 * Pass NULL for sizes-only query via get_info. */
int dotgpu_solver_get_symbolic(dotgpu_solver* s, int32_t* perm, int32_t* super_ptr, int64_t* row_ptr, int32_t* rows,
                               int32_t* parent, int32_t* level);

/* ------------------------------------------------------------------------------------------
 * Domain-decomposition set-up from element labels: ADMMDDTimeStepper ctor + precompute
 * (ADMMDDTimeStepper.cpp:155-278, 457-496), Mesh::constructSubmesh (Mesh.cpp:855-905),
 * dup (DOTTimeStepper.cpp:38-56).  Labels come from the reference's METIS wrapper (Utils/METIS.hpp)
 * and are an INPUT here (bit-exact labels require the vendored METIS).
 * ------------------------------------------------------------------------------------------ */
typedef struct dotgpu_dd dotgpu_dd;
/* METIS<3>::partMesh (Utils/METIS.hpp:109-160, options :265-321; called at ADMMDDTimeStepper.cpp:88-92): k-way partition of the
 * dual graph (ncommon = 3) with the reference's vendored METIS 5.1.0 (libdotmetis.so, built by dot_b200/build.py from the sources
 * under the reference tree; 64-bit idx_t, 32-bit real_t).  epart_out [nT] = subdomain label per tet, bit-exact with the reference's.
 * Returns DOTGPU_ERR_STATE when libdotmetis.so is not available (then pass labels produced elsewhere to the calls below). */
int dotgpu_partition(int nV, int nT, const int32_t* tets, int k, int32_t* epart_out);
/* METIS<3>::partMesh_nodes (Utils/METIS.hpp:161-212: METIS_PartMeshNodal, same option vector; LBFGSTimeStepper.cpp:71-74): npart_out [nV] */
int dotgpu_partition_nodes(int nV, int nT, const int32_t* tets, int k, int32_t* npart_out);
int dotgpu_dd_create(dotgpu_dd** out, int nV, int nT, const int32_t* tets, const int32_t* epart, int k,
                     const uint8_t* fixed_mask);
void dotgpu_dd_destroy(dotgpu_dd* d);
int dotgpu_dd_num_local_verts(dotgpu_dd* d, int s);
int dotgpu_dd_num_elems(dotgpu_dd* d, int s);
int64_t dotgpu_dd_nnz(dotgpu_dd* d, int s);                   /* s = -1: the global matrix */
int dotgpu_dd_get_l2g(dotgpu_dd* d, int s, int32_t* l2g);      /* localVIToGlobal_subdomain[s] */
int dotgpu_dd_get_fixed_local(dotgpu_dd* d, int s, int32_t* out, int* count);
int dotgpu_dd_get_pattern(dotgpu_dd* d, int s, int32_t* ia, int32_t* ja); /* s = -1: global */
int dotgpu_dd_get_dup(dotgpu_dd* d, int32_t* dup);

/* ------------------------------------------------------------------------------------------
 * Scripted Dirichlet motion: AnimScripter<3> (AnimScripter.cpp:29-453) + handle detection
 * (IglUtils::findBorderVerts, IglUtils.cpp:909-927).  Host-side, tiny.
 * ------------------------------------------------------------------------------------------ */
typedef struct dotgpu_anim dotgpu_anim;
int dotgpu_anim_create(dotgpu_anim** out, int kind, int nV, const double* V_rest, double handle_ratio);
void dotgpu_anim_destroy(dotgpu_anim* a);
int dotgpu_anim_fixed_mask(dotgpu_anim* a, uint8_t* mask_out);    /* [nV] */
int dotgpu_anim_step(dotgpu_anim* a, double* x_inout, double dt); /* stepAnimScript: moves handle rows of x */
/* same, and *dirichlet_set_changed = AnimScripter::stepAnimScript's return value: 1 when the script released / added handles in this step
 * (rubberBandPull).  The caller then fetches the new set with dotgpu_anim_fixed_mask and calls dotgpu_stepper_set_fixed (Optimizer.cpp:334-336). */
int dotgpu_anim_step_ex(dotgpu_anim* a, double* x_inout, double dt, int* dirichlet_set_changed);

/* ------------------------------------------------------------------------------------------
 * Device-resident DOT time stepper: sits under Optimizer<3>'s virtuals precompute / fullyImplicit /
 * solve_oneStep / updatePrecondMtrAndFactorize as DOTTimeStepper overrides them
 * (TimeStepper/DOTTimeStepper.cpp:150-182, 273-504; Optimizer.cpp:327-368, 442-610, 752-881, 1076-1093).
 * ------------------------------------------------------------------------------------------ */
typedef struct dotgpu_stepper dotgpu_stepper;

/* flags of dotgpu_stepper_config */
#define DOTGPU_FLAG_NEWTON 2 /* Projected Newton instead of DOT: Optimizer::solve_oneStep (Optimizer.cpp:703-749) - the global PD-projected
                              * Hessian is re-assembled, factorised and solved in EVERY iteration, step length starts at 1 (initStepSize,
                              * :1076-1093); `timeStepper Newton` scripts, the reference's "1 subdomain" case.  Use num_subdomains = 1,
                              * epart all 0. */

/* the other L-BFGS initialisers of the reference that share the kernels (TimeStepper/LBFGSTimeStepper.cpp:108-265, 286-335,
 * 339-420): same two-loop recursion and history, line search from step 1 (Optimizer::initStepSize uses p.Hp for DOT only), the initial
 * inverse Hessian refreshed once per time step like DOT's. */
#define DOTGPU_FLAG_LBFGS_H 4  /* `timeStepper LBFGSH` (D0T_H): the global PD-projected Hessian, one factorisation; num_subdomains = 1, epart ignored */
#define DOTGPU_FLAG_LBFGS_JH 8 /* `timeStepper LBFGSJH k` (D0T_JH): block Jacobi of that matrix over the NODE partition config.node_part
                                * (METIS<3>::partMesh_nodes = dotgpu_partition_nodes); num_subdomains = k, epart ignored */

typedef struct dotgpu_stepper_config {
    int32_t device;
    int32_t energy_type;      /* DOTGPU_ENERGY_* */
    int32_t num_subdomains;   /* k of `timeStepper DOT k` */
    int32_t history;          /* L-BFGS pairs, reference: 5 (DOTTimeStepper.cpp:47) */
    double dt;                /* 0.025 in every shipped script */
    double gravity[3];        /* (0,-9.80665,0), Optimizer.cpp:107-110 */
    double rel_tol;           /* epsilon of the stopping rule, reference default 1e-5 */
    double YM, PR, rho;       /* stiffness / density script keys */
    int32_t max_iters;        /* 10000, DOTTimeStepper.cpp:302 */
    int32_t rank, world;      /* multi-GPU: subdomains are dealt to ranks; world=1 for one GPU */
    const void* nccl_unique_id; /* ncclUniqueId bytes (128) when world>1, else NULL */
    int32_t target_fixed_count; /* #fixed verts in the tolerance formula; reference uses 1 */
    int32_t flags;            /* DOTGPU_FLAG_* */
    const int32_t* node_part; /* DOTGPU_FLAG_LBFGS_JH: node labels [nV] in [0, num_subdomains); else NULL */
} dotgpu_stepper_config;
void dotgpu_stepper_default_config(dotgpu_stepper_config* c);
/* Round-robin map of subdomains to ranks (bookkeeping / DOTGPU_BALANCE=0; see dotgpu_balanced_owner and
 * dotgpu_stepper_get_owned for what a stepper uses): writes the ascending list of
 * subdomain ids rank `rank` of `world` would factor and solve; returns their number (or a negative error code).  out may be NULL. */
int dotgpu_owned_subdomains(int num_subdomains, int rank, int world, int32_t* out);
/* The map a multi-GPU stepper actually uses: whole subdomains to ranks, balancing the given weights (the stepper passes nnz(L_s) of
 * its own symbolic analysis) by longest-processing-time greedy; deterministic, so every rank computes the same owner_out [k]. */
int dotgpu_balanced_owner(int num_subdomains, const double* weight, int world, int32_t* owner_out);
/* rank 0 calls this and broadcasts the 128 bytes (e.g. torch.distributed) before every rank creates its stepper */
int dotgpu_nccl_unique_id(void* out128);

typedef struct dotgpu_frame_stats {
    int32_t iters;            /* L-BFGS iterations of this frame */
    int32_t halvings;         /* line-search halvings of this frame */
    int32_t energy_evals;
    int32_t converged;        /* 1 if ||g||^2 <= targetGRes */
    double E, grad_sqnorm, target;
    double ms_total, ms_solve, ms_refresh;  /* device times (CUDA events) */
    double ms_precond;        /* device time of the preconditioner applications (K5) of this frame, CUDA events on the stepper's stream */
    int32_t precond_calls;
    int32_t line_search_failed; /* 1 if the step length halved to 0 (Optimizer.cpp:816-824): the time step ended there, converged = 0 */
} dotgpu_frame_stats;

/* Builds everything precompute() builds: mesh features, DD from labels, patterns, symbolic analysis of
 * every subdomain this rank owns, rest-state Hessians + factorisation.  V_rest is the normalised rest
 * shape (main.cpp:709-710 applied), fixed_mask the Dirichlet set (AnimScripter handles). */
int dotgpu_stepper_create(dotgpu_stepper** out, const dotgpu_stepper_config* cfg, int nV, int nT, const double* V_rest,
                          const int32_t* tets, const int32_t* epart, const uint8_t* fixed_mask);
void dotgpu_stepper_destroy(dotgpu_stepper* s);
/* One time step = DOTTimeStepper::fullyImplicit + the BE update of Optimizer::solve (Optimizer.cpp:354-361).
 * x_inout [nV*3] host: on entry x^n with the scripted Dirichlet move applied (what result.V holds after
 * stepAnimScript), on return the converged positions. */
int dotgpu_stepper_frame(dotgpu_stepper* s, double* x_inout, dotgpu_frame_stats* stats);
/* Same time step with the positions already resident on the device (they are: the stepper keeps x^n): only the
 * scripted Dirichlet targets travel, as `count` vertex ids + positions [count*3] (host).  Nothing is copied back;
 * read the result with dotgpu_stepper_get_state. */
int dotgpu_stepper_frame_resident(dotgpu_stepper* s, const int32_t* fixed_idx, const double* fixed_pos, int count,
                                  dotgpu_frame_stats* stats);
/* restart (Optimizer ctor :126-177 reading `status<n>`): positions + velocity; refreshes the Hessians at x. */
int dotgpu_stepper_set_state(dotgpu_stepper* s, const double* x, const double* velocity);
int dotgpu_stepper_get_state(dotgpu_stepper* s, double* x, double* velocity, double* xTilde);
/* DOTTimeStepper::updatePrecondMtrAndFactorize (DOTTimeStepper.cpp:185-270; called from Optimizer::solve when the AnimScripter changes the
 * Dirichlet set, Optimizer.cpp:334-336): new Dirichlet set fixed_mask [nV] -> patterns, fill lists and symbolic analysis of every owned
 * subdomain are rebuilt, then the Hessians and the factorisation at x_eval [nV*3] (the reference uses result.V with this frame's scripted
 * move applied; NULL: the resident x^n).  The subdomain -> rank map is kept. */
int dotgpu_stepper_set_fixed(dotgpu_stepper* s, const uint8_t* fixed_mask, const double* x_eval);
/* per-iteration log of the last frame: rows (alpha, E, |g|^2), row 0 = after initX; returns rows written */
int dotgpu_stepper_get_iter_log(dotgpu_stepper* s, double* out, int max_rows);
/* checkers: matrix values after the last refresh (s=-1 global, else subdomain index; subdomains of other
 * ranks return DOTGPU_ERR_INVALID), one preconditioner application p = D^-1 sum R^T H_s^-1 R q
 * (DOTTimeStepper.cpp:406-450), energy / gradient of the incremental potential at x. */
int dotgpu_stepper_get_matrix(dotgpu_stepper* s, int sub, double* a_out);
int dotgpu_stepper_get_dd(dotgpu_stepper* s, dotgpu_dd** dd_out); /* borrowed pointer */
int dotgpu_stepper_precondition(dotgpu_stepper* s, const double* q, double* p_out);
int dotgpu_stepper_eval(dotgpu_stepper* s, const double* x, double* E_out, double* g_out);
int dotgpu_stepper_get_target(dotgpu_stepper* s, double* target);
/* Optimizer::setRelGL2Tol (Optimizer.cpp:222-228): relative tolerance of the following time steps (scripts give one per frame
 * through `tol`, main.cpp:108-117); recomputes targetGRes.  Host-only, no device work. */
int dotgpu_stepper_set_rel_tol(dotgpu_stepper* s, double rel_tol);
/* device-only timing helpers for bench.py: run `reps` energy+gradient evaluations (K1+K2) / Hessian
 * refreshes (K3+K4+factor) / preconditioner applications (K5) on resident data, return avg ms by CUDA events */
int dotgpu_stepper_time_kernels(dotgpu_stepper* s, int which, int reps, double* ms_out);
int64_t dotgpu_stepper_launch_count(dotgpu_stepper* s); /* kernels launched by this handle so far */
/* diagnostics (DOTGPU_SOLVE_TRACE=1 at create time): %globaltimer stamps of the first 96 chunks of every CTA of the last preconditioner
 * application, 8 x uint64 per (CTA, chunk): posted, vector ready, data landed, products done, published, queue slot, warp 0 / last warp out of the products.  Returns the number
 * of uint64 words (0: tracing off); out may be NULL to query the size. */
int dotgpu_stepper_get_solve_trace(dotgpu_stepper* s, uint64_t* out, int64_t max_words);
/* byte model of the matrix fill K4 (bench.py): stored values of [global | owned subdomain] matrices (8 B each, written once per refresh),
 * 3x3 blocks of those matrices, and elemental 3x3 blocks gathered into them (72 B each, read once per refresh) */
int dotgpu_stepper_get_fill_stats(dotgpu_stepper* s, int64_t* nnz_out, int64_t* blocks_out, int64_t* gathered_blocks_out);
/* the subdomains this rank factors and solves (ascending; balanced by nnz(L) over the ranks); returns their number */
int dotgpu_stepper_get_owned(dotgpu_stepper* s, int32_t* out);
int dotgpu_stepper_get_solver_info(dotgpu_stepper* s, int sub, dotgpu_solver_info* info);

#ifdef __cplusplus
}
#endif
#endif /* DOTGPU_H */
