// extern "C" surface of libdotgpu (include/dotgpu.h).  No exceptions cross this boundary.
#include <cstdlib>
#include <cstring>
#include <memory>

#include "anim_host.h"
#include "chol_numeric.h"
#include "comm.h"
#include "device_mesh.h"
#include "linalg.h"
#include "mesh_host.h"
#include "stepper.h"

using namespace dotgpu;

namespace dotgpu {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
}  // namespace dotgpu

#define API_BEGIN try {
#define API_END                                      \
    return DOTGPU_OK;                                \
    }                                                \
    catch (const dotgpu::Error& e) {                 \
        set_last_error(e.what());                    \
        return e.code;                               \
    }                                                \
    catch (const std::invalid_argument& e) {         \
        set_last_error(e.what());                    \
        return DOTGPU_ERR_INVALID;                   \
    }                                                \
    catch (const std::exception& e) {                \
        set_last_error(e.what());                    \
        return DOTGPU_ERR_CUDA;                      \
    }

static void require_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw Error(DOTGPU_ERR_NO_DEVICE, "no CUDA device visible");
    DG_REQUIRE(device >= 0 && device < n, "device index out of range");
    DG_CUDA(cudaSetDevice(device));
}

struct dotgpu_energy {
    int device = 0;
    cudaStream_t st = nullptr;
    DeviceMesh mesh;
    DevBuf<double> x, out, sc;
    std::vector<int32_t> tets;
    std::vector<uint8_t> fixed;
    ~dotgpu_energy() {
        if (st) cudaStreamDestroy(st);
    }
};

struct dotgpu_solver {
    int device = -1, n = 0;
    cudaStream_t st = nullptr;
    std::vector<int32_t> ia, ja;
    Symbolic sym_only;  // when device < 0
    CholBatch chol;
    DevBuf<double> a, b, x, tmp;
    DevBuf<int> d_ia, d_ja, d_tp, d_tslot, d_trow, d_perm;
    bool has_values = false;
    ~dotgpu_solver() {
        if (st) cudaStreamDestroy(st);
    }
    const Symbolic& sym() const { return device < 0 ? sym_only : chol.sym[0]; }
};

struct dotgpu_dd {
    DDHost dd;
    bool borrowed = false;
};
struct dotgpu_anim {
    AnimHost a;
};
struct dotgpu_stepper {
    Stepper s;
    dotgpu_dd dd_view;
};

namespace {
__global__ void k_permute_in(int n, const int* __restrict__ perm, const double* __restrict__ in, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
__global__ void k_permute_out(int n, const int* __restrict__ perm, const double* __restrict__ in, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[perm[i]] = in[i];
}
}  // namespace

extern "C" {

const char* dotgpu_last_error(void) { return g_last_error.c_str(); }
int dotgpu_version(void) { return 100; }
int dotgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int dotgpu_owned_subdomains(int num_subdomains, int rank, int world, int32_t* out) {
    if (num_subdomains < 1 || world < 1 || rank < 0 || rank >= world) return DOTGPU_ERR_INVALID;
    const std::vector<int> o = owned_subdomains(num_subdomains, rank, world);
    if (out)
        for (size_t i = 0; i < o.size(); ++i) out[i] = o[i];
    return (int)o.size();
}

int dotgpu_balanced_owner(int num_subdomains, const double* weight, int world, int32_t* owner_out) {
    if (num_subdomains < 1 || world < 1 || !weight || !owner_out) return DOTGPU_ERR_INVALID;
    std::vector<double> w(weight, weight + num_subdomains);
    std::vector<int> o = balanced_owner(w, world);
    for (int s = 0; s < num_subdomains; ++s) owner_out[s] = o[s];
    return DOTGPU_OK;
}

int dotgpu_mesh_features(int nV, int nT, const double* V_rest, const int32_t* tets, double YM, double PR, double rho,
                         double* DmInv_out, double* vol_out, double* mass_out, double* mu_out, double* lambda_out) {
    API_BEGIN
    DG_REQUIRE(nV > 0 && nT > 0 && V_rest && tets && DmInv_out && vol_out && mass_out && mu_out && lambda_out, "null argument");
    for (size_t i = 0; i < 4 * (size_t)nT; ++i) DG_REQUIRE(tets[i] >= 0 && tets[i] < nV, "tet index out of range");
    mesh_features(nV, nT, V_rest, tets, YM, PR, rho, DmInv_out, vol_out, mass_out, mu_out, lambda_out);
    API_END
}

// ---------------------------------------------------------------- energy
int dotgpu_energy_create(dotgpu_energy** out, int device, int energy_type, int nV, int nT, const int32_t* tets, const double* DmInv,
                         const double* vol, const double* mu, const double* lambda, const uint8_t* fixed_mask) {
    API_BEGIN
    DG_REQUIRE(out && tets && DmInv && vol && mu && lambda, "null argument");
    require_device(device);
    std::unique_ptr<dotgpu_energy> e(new dotgpu_energy());
    e->device = device;
    DG_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    e->mesh.init(energy_type, nV, nT, tets, DmInv, vol, mu, lambda, nullptr, fixed_mask, e->st);
    e->tets.assign(tets, tets + 4 * (size_t)nT);
    e->fixed.assign(nV, 0);
    if (fixed_mask) e->fixed.assign(fixed_mask, fixed_mask + nV);
    e->x.alloc(3 * (size_t)nV);
    e->sc.alloc(4);
    *out = e.release();
    API_END
}
void dotgpu_energy_destroy(dotgpu_energy* e) { delete e; }

int dotgpu_energy_set_fixed(dotgpu_energy* e, const uint8_t* fixed_mask) {
    API_BEGIN
    DG_REQUIRE(e && fixed_mask, "null argument");
    DG_CUDA(cudaSetDevice(e->device));
    e->mesh.set_fixed(fixed_mask, e->st);
    e->fixed.assign(fixed_mask, fixed_mask + e->mesh.nV);
    API_END
}

int dotgpu_energy_value(dotgpu_energy* e, const double* x, double coef, double* E_out) {
    API_BEGIN
    DG_REQUIRE(e && x && E_out, "null argument");
    DG_CUDA(cudaSetDevice(e->device));
    e->x.upload(x, 3 * (size_t)e->mesh.nV, e->st);
    launch_energy(e->mesh, e->x.p, nullptr, coef, e->sc.p, e->st);
    e->sc.download(E_out, 1, e->st);
    API_END
}

int dotgpu_energy_per_elem(dotgpu_energy* e, const double* x, double* out) {
    API_BEGIN
    DG_REQUIRE(e && x && out, "null argument");
    DG_CUDA(cudaSetDevice(e->device));
    e->x.upload(x, 3 * (size_t)e->mesh.nV, e->st);
    if (e->out.n < (size_t)e->mesh.nT) e->out.alloc(e->mesh.nT);
    launch_energy_per_elem(e->mesh, e->x.p, e->out.p, e->st);
    e->out.download(out, e->mesh.nT, e->st);
    API_END
}

int dotgpu_energy_gradient(dotgpu_energy* e, const double* x, double coef, double* g_out) {
    API_BEGIN
    DG_REQUIRE(e && x && g_out, "null argument");
    DG_CUDA(cudaSetDevice(e->device));
    const size_t n3 = 3 * (size_t)e->mesh.nV;
    e->x.upload(x, n3, e->st);
    if (e->out.n < n3) e->out.alloc(n3);
    launch_gradient(e->mesh, e->x.p, nullptr, coef, e->out.p, e->st);
    e->out.download(g_out, n3, e->st);
    API_END
}

int dotgpu_energy_svd(dotgpu_energy* e, const double* x, double* F_out, double* U_out, double* Sigma_out, double* V_out) {
    API_BEGIN
    DG_REQUIRE(e && x, "null argument");
    DG_CUDA(cudaSetDevice(e->device));
    const size_t nT = e->mesh.nT;
    e->x.upload(x, 3 * (size_t)e->mesh.nV, e->st);
    if (e->out.n < 30 * nT) e->out.alloc(30 * nT);
    double *F = e->out.p, *U = F + 9 * nT, *V = U + 9 * nT, *S = V + 9 * nT;
    launch_svd(e->mesh, e->x.p, F, U, S, V, e->st);
    DG_CUDA(cudaStreamSynchronize(e->st));
    if (F_out) DG_CUDA(cudaMemcpy(F_out, F, 9 * nT * sizeof(double), cudaMemcpyDeviceToHost));
    if (U_out) DG_CUDA(cudaMemcpy(U_out, U, 9 * nT * sizeof(double), cudaMemcpyDeviceToHost));
    if (V_out) DG_CUDA(cudaMemcpy(V_out, V, 9 * nT * sizeof(double), cudaMemcpyDeviceToHost));
    if (Sigma_out) DG_CUDA(cudaMemcpy(Sigma_out, S, 3 * nT * sizeof(double), cudaMemcpyDeviceToHost));
    API_END
}

int dotgpu_energy_elem_hessians(dotgpu_energy* e, const double* x, double coef, int projectSPD, double* He_out, int32_t* vInds_out) {
    API_BEGIN
    DG_REQUIRE(e && x && He_out, "null argument");
    DG_CUDA(cudaSetDevice(e->device));
    const size_t nT = e->mesh.nT;
    e->x.upload(x, 3 * (size_t)e->mesh.nV, e->st);
    launch_elem_hessians(e->mesh, e->x.p, coef, projectSPD != 0, e->st);
    if (e->out.n < 144 * nT) e->out.alloc(144 * nT);
    launch_he_to_dense(e->mesh, e->out.p, e->st);
    e->out.download(He_out, 144 * nT, e->st);
    if (vInds_out)  // Energy.cpp:771-776
        for (size_t i = 0; i < 4 * nT; ++i) vInds_out[i] = e->fixed[e->tets[i]] ? -e->tets[i] - 1 : e->tets[i];
    API_END
}

// ---------------------------------------------------------------- solver
int dotgpu_solver_create(dotgpu_solver** out, int device, int n, const int32_t* ia, const int32_t* ja) {
    API_BEGIN
    DG_REQUIRE(out && ia && ja && n > 0, "null or empty pattern");
    std::unique_ptr<dotgpu_solver> s(new dotgpu_solver());
    s->n = n;
    s->ia.assign(ia, ia + n + 1);
    DG_REQUIRE(s->ia[0] == 0, "ia must be 0-based");
    s->ja.assign(ja, ja + s->ia[n]);
    s->device = device;
    if (device < 0) {  // symbolic analysis only (no device needed)
        s->sym_only.analyze(n, s->ia.data(), s->ja.data());
    } else {
        require_device(device);
        DG_CUDA(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
        s->chol.analyze({s->ia.data()}, {s->ja.data()}, {n}, 21, s->st);
        s->a.alloc(s->ja.size());
        s->b.alloc(n);
        s->x.alloc(n);
        s->tmp.alloc(n);
        s->d_ia.upload(s->ia, s->st);
        s->d_ja.upload(s->ja, s->st);
        std::vector<int> perm(s->chol.sym[0].perm.begin(), s->chol.sym[0].perm.end());
        s->d_perm.upload(perm, s->st);
        // transposed index of the strictly upper part
        std::vector<int> tp(n + 1, 0);
        for (int i = 0; i < n; ++i)
            for (int k = s->ia[i]; k < s->ia[i + 1]; ++k)
                if (s->ja[k] != i) tp[s->ja[k] + 1]++;
        for (int i = 0; i < n; ++i) tp[i + 1] += tp[i];
        std::vector<int> tslot(std::max(tp[n], 1)), trow(std::max(tp[n], 1)), cur(tp.begin(), tp.end() - 1);
        for (int i = 0; i < n; ++i)
            for (int k = s->ia[i]; k < s->ia[i + 1]; ++k)
                if (s->ja[k] != i) {
                    int c = s->ja[k];
                    tslot[cur[c]] = k;
                    trow[cur[c]++] = i;
                }
        s->d_tp.upload(tp, s->st);
        s->d_tslot.upload(tslot, s->st);
        s->d_trow.upload(trow, s->st);
    }
    *out = s.release();
    API_END
}
void dotgpu_solver_destroy(dotgpu_solver* s) { delete s; }

int dotgpu_solver_set_values(dotgpu_solver* s, const double* a) {
    API_BEGIN
    DG_REQUIRE(s && a, "null argument");
    if (s->device < 0) throw Error(DOTGPU_ERR_NO_DEVICE, "symbolic-only solver handle");
    DG_CUDA(cudaSetDevice(s->device));
    s->a.upload(a, s->ja.size(), s->st);
    DG_CUDA(cudaStreamSynchronize(s->st));
    s->has_values = true;
    s->chol.factorized = false;
    API_END
}

int dotgpu_solver_factorize(dotgpu_solver* s) {
    API_BEGIN
    DG_REQUIRE(s, "null argument");
    if (s->device < 0) throw Error(DOTGPU_ERR_NO_DEVICE, "symbolic-only solver handle");
    if (!s->has_values) throw Error(DOTGPU_ERR_STATE, "factorize before set_values");
    DG_CUDA(cudaSetDevice(s->device));
    s->chol.factorize(s->a.p, s->st);
    s->chol.check_status(s->st);
    API_END
}

int dotgpu_solver_solve(dotgpu_solver* s, const double* rhs, double* x) {
    API_BEGIN
    DG_REQUIRE(s && rhs && x, "null argument");
    if (s->device < 0) throw Error(DOTGPU_ERR_NO_DEVICE, "symbolic-only solver handle");
    DG_CUDA(cudaSetDevice(s->device));
    const int n = s->n;
    s->tmp.upload(rhs, n, s->st);
    static const bool by_levels = std::getenv("DOTGPU_SOLVE_LEVELS") != nullptr;  // cross-check path: one launch per level and direction
    if (by_levels) {
        k_permute_in<<<ceil_div(n, 256), 256, 0, s->st>>>(n, s->d_perm.p, s->tmp.p, s->b.p);
        s->chol.solve_levels(s->b.p, s->x.p, s->st);
        count_launch(1);
    } else {
        s->chol.solve(s->tmp.p, s->d_perm.p, s->x.p, s->st);
    }
    k_permute_out<<<ceil_div(n, 256), 256, 0, s->st>>>(n, s->d_perm.p, s->x.p, s->tmp.p);
    count_launch(1);
    s->tmp.download(x, n, s->st);
    API_END
}

int dotgpu_solver_multiply(dotgpu_solver* s, const double* x, double* y) {
    API_BEGIN
    DG_REQUIRE(s && x && y, "null argument");
    if (s->device < 0) throw Error(DOTGPU_ERR_NO_DEVICE, "symbolic-only solver handle");
    if (!s->has_values) throw Error(DOTGPU_ERR_STATE, "multiply before set_values");
    DG_CUDA(cudaSetDevice(s->device));
    s->tmp.upload(x, s->n, s->st);
    launch_spmv_sym(s->n, s->d_ia.p, s->d_ja.p, s->a.p, s->d_tp.p, s->d_tslot.p, s->d_trow.p, s->tmp.p, s->b.p, s->st);
    s->b.download(y, s->n, s->st);
    API_END
}

static void fill_info(const Symbolic& S, int64_t nnz_a, int64_t bytes, dotgpu_solver_info* info) {
    info->n = S.n;
    info->nsuper = S.nsuper;
    info->nlevels = S.nlevels;
    info->max_front = S.max_front;
    info->max_nscol = S.max_nscol;
    info->nnz_a = nnz_a;
    // nnz(L) of the supernodal factor: dense lower-triangular diagonal blocks + dense sub-diagonal blocks (what the solves stream)
    info->nnz_l = 0;
    for (int k = 0; k < S.nsuper; ++k) {
        const int64_t ns = S.nscol(k), m = S.front(k);
        info->nnz_l += ns * (ns + 1) / 2 + (m - ns) * ns;
    }
    info->flops = S.flops;
    info->device_bytes = bytes;
}

int dotgpu_solver_get_info(dotgpu_solver* s, dotgpu_solver_info* info) {
    API_BEGIN
    DG_REQUIRE(s && info, "null argument");
    fill_info(s->sym(), (int64_t)s->ja.size(), s->device < 0 ? 0 : s->chol.device_bytes(), info);
    API_END
}

int dotgpu_solver_get_symbolic(dotgpu_solver* s, int32_t* perm, int32_t* super_ptr, int64_t* row_ptr, int32_t* rows, int32_t* parent,
                               int32_t* level) {
    API_BEGIN
    DG_REQUIRE(s, "null argument");
    const Symbolic& S = s->sym();
    if (perm) std::memcpy(perm, S.perm.data(), S.n * sizeof(int32_t));
    if (super_ptr) std::memcpy(super_ptr, S.super_ptr.data(), (S.nsuper + 1) * sizeof(int32_t));
    if (row_ptr) std::memcpy(row_ptr, S.row_ptr.data(), (S.nsuper + 1) * sizeof(int64_t));
    if (rows) std::memcpy(rows, S.rows.data(), S.rows.size() * sizeof(int32_t));
    if (parent) std::memcpy(parent, S.parent.data(), S.nsuper * sizeof(int32_t));
    if (level) std::memcpy(level, S.level.data(), S.nsuper * sizeof(int32_t));
    API_END
}

// ---------------------------------------------------------------- domain decomposition
int dotgpu_dd_create(dotgpu_dd** out, int nV, int nT, const int32_t* tets, const int32_t* epart, int k, const uint8_t* fixed_mask) {
    API_BEGIN
    DG_REQUIRE(out && tets && epart && nV > 0 && nT > 0, "null or empty argument");
    for (size_t i = 0; i < 4 * (size_t)nT; ++i) DG_REQUIRE(tets[i] >= 0 && tets[i] < nV, "tet index out of range");
    std::unique_ptr<dotgpu_dd> d(new dotgpu_dd());
    d->dd.build(nV, nT, tets, epart, k, fixed_mask, nullptr, 0.0, nullptr, false);
    *out = d.release();
    API_END
}
void dotgpu_dd_destroy(dotgpu_dd* d) {
    if (d && !d->borrowed) delete d;
}
static const DDHost& ddof(dotgpu_dd* d) { return d->dd; }
int dotgpu_dd_num_local_verts(dotgpu_dd* d, int s) {
    if (!d || s < 0 || s >= ddof(d).k) return DOTGPU_ERR_INVALID;
    return (int)ddof(d).subs[s].l2g.size();
}
int dotgpu_dd_num_elems(dotgpu_dd* d, int s) {
    if (!d || s < 0 || s >= ddof(d).k) return DOTGPU_ERR_INVALID;
    return (int)ddof(d).subs[s].elems.size();
}
int64_t dotgpu_dd_nnz(dotgpu_dd* d, int s) {
    if (!d || s < -1 || s >= ddof(d).k) return DOTGPU_ERR_INVALID;
    return s < 0 ? ddof(d).gpat.nnz() : ddof(d).subs[s].pat.nnz();
}
int dotgpu_dd_get_l2g(dotgpu_dd* d, int s, int32_t* l2g) {
    API_BEGIN
    DG_REQUIRE(d && l2g && s >= 0 && s < ddof(d).k, "bad argument");
    std::memcpy(l2g, ddof(d).subs[s].l2g.data(), ddof(d).subs[s].l2g.size() * sizeof(int32_t));
    API_END
}
int dotgpu_dd_get_fixed_local(dotgpu_dd* d, int s, int32_t* out, int* count) {
    API_BEGIN
    DG_REQUIRE(d && count && s >= 0 && s < ddof(d).k, "bad argument");
    const auto& f = ddof(d).subs[s].fixed_local;
    *count = (int)f.size();
    if (out) std::memcpy(out, f.data(), f.size() * sizeof(int32_t));
    API_END
}
int dotgpu_dd_get_pattern(dotgpu_dd* d, int s, int32_t* ia, int32_t* ja) {
    API_BEGIN
    DG_REQUIRE(d && s >= -1 && s < ddof(d).k, "bad argument");
    const MatrixPattern& P = s < 0 ? ddof(d).gpat : ddof(d).subs[s].pat;
    DG_REQUIRE(!P.ia.empty(), "subdomain not built on this rank");
    if (ia) std::memcpy(ia, P.ia.data(), P.ia.size() * sizeof(int32_t));
    if (ja) std::memcpy(ja, P.ja.data(), P.ja.size() * sizeof(int32_t));
    API_END
}
int dotgpu_dd_get_dup(dotgpu_dd* d, int32_t* dup) {
    API_BEGIN
    DG_REQUIRE(d && dup, "null argument");
    std::memcpy(dup, ddof(d).dup.data(), ddof(d).dup.size() * sizeof(int32_t));
    API_END
}

// ---------------------------------------------------------------- anim scripter
int dotgpu_anim_create(dotgpu_anim** out, int kind, int nV, const double* V_rest, double handle_ratio) {
    API_BEGIN
    DG_REQUIRE(out && V_rest && nV > 0, "null or empty argument");
    std::unique_ptr<dotgpu_anim> a(new dotgpu_anim());
    a->a.init(kind, nV, V_rest, handle_ratio);
    *out = a.release();
    API_END
}
void dotgpu_anim_destroy(dotgpu_anim* a) { delete a; }
int dotgpu_anim_fixed_mask(dotgpu_anim* a, uint8_t* mask_out) {
    API_BEGIN
    DG_REQUIRE(a && mask_out, "null argument");
    a->a.fixed_mask(mask_out);
    API_END
}
int dotgpu_anim_step(dotgpu_anim* a, double* x_inout, double dt) {
    API_BEGIN
    DG_REQUIRE(a && x_inout, "null argument");
    a->a.step(x_inout, dt);
    API_END
}
int dotgpu_anim_step_ex(dotgpu_anim* a, double* x_inout, double dt, int* dirichlet_set_changed) {
    API_BEGIN
    DG_REQUIRE(a && x_inout, "null argument");
    const int f = a->a.step(x_inout, dt);
    if (dirichlet_set_changed) *dirichlet_set_changed = f;
    API_END
}

// ---------------------------------------------------------------- stepper
void dotgpu_stepper_default_config(dotgpu_stepper_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->device = 0;
    c->energy_type = DOTGPU_ENERGY_FCR;
    c->num_subdomains = 4;  // Config.cpp:76-80 turns k < 2 into 4
    c->history = 5;
    c->dt = 0.025;
    c->gravity[0] = 0.0;
    c->gravity[1] = -9.80665;
    c->gravity[2] = 0.0;
    c->rel_tol = 1.0e-5;
    c->YM = 1.0e5;
    c->PR = 0.4;
    c->rho = 1000.0;
    c->max_iters = 10000;
    c->rank = 0;
    c->world = 1;
    c->nccl_unique_id = nullptr;
    c->target_fixed_count = 1;
    c->flags = 0;
    c->node_part = nullptr;
}

int dotgpu_nccl_unique_id(void* out128) {
    API_BEGIN
    DG_REQUIRE(out128, "null argument");
    Comm::unique_id(out128);
    API_END
}

int dotgpu_stepper_create(dotgpu_stepper** out, const dotgpu_stepper_config* cfg, int nV, int nT, const double* V_rest,
                          const int32_t* tets, const int32_t* epart, const uint8_t* fixed_mask) {
    API_BEGIN
    DG_REQUIRE(out && cfg, "null argument");
    for (size_t i = 0; tets && i < 4 * (size_t)nT; ++i) DG_REQUIRE(tets[i] >= 0 && tets[i] < nV, "tet index out of range");
    std::unique_ptr<dotgpu_stepper> s(new dotgpu_stepper());
    s->s.create(*cfg, nV, nT, V_rest, tets, epart, fixed_mask);
    *out = s.release();
    API_END
}
void dotgpu_stepper_destroy(dotgpu_stepper* s) {
    if (s) {
        cudaSetDevice(s->s.cfg.device);
        delete s;
    }
}
int dotgpu_stepper_frame(dotgpu_stepper* s, double* x_inout, dotgpu_frame_stats* stats) {
    API_BEGIN
    DG_REQUIRE(s && x_inout, "null argument");
    DG_CUDA(cudaSetDevice(s->s.cfg.device));
    s->s.frame(x_inout, stats);
    API_END
}
int dotgpu_stepper_frame_resident(dotgpu_stepper* s, const int32_t* fixed_idx, const double* fixed_pos, int count,
                                  dotgpu_frame_stats* stats) {
    API_BEGIN
    DG_REQUIRE(s, "null argument");
    DG_CUDA(cudaSetDevice(s->s.cfg.device));
    s->s.frame_resident(fixed_idx, fixed_pos, count, stats);
    API_END
}
int dotgpu_stepper_set_state(dotgpu_stepper* s, const double* x, const double* velocity) {
    API_BEGIN
    DG_REQUIRE(s && x, "null argument");
    DG_CUDA(cudaSetDevice(s->s.cfg.device));
    s->s.set_state(x, velocity);
    API_END
}
int dotgpu_stepper_set_fixed(dotgpu_stepper* s, const uint8_t* fixed_mask, const double* x_eval) {
    API_BEGIN
    DG_REQUIRE(s && fixed_mask, "null argument");
    DG_CUDA(cudaSetDevice(s->s.cfg.device));
    s->s.set_fixed(fixed_mask, x_eval);
    API_END
}
int dotgpu_stepper_get_state(dotgpu_stepper* s, double* x, double* velocity, double* xTilde) {
    API_BEGIN
    DG_REQUIRE(s, "null argument");
    DG_CUDA(cudaSetDevice(s->s.cfg.device));
    s->s.get_state(x, velocity, xTilde);
    API_END
}
int dotgpu_stepper_get_iter_log(dotgpu_stepper* s, double* out, int max_rows) {
    if (!s || !out) return DOTGPU_ERR_INVALID;
    int rows = std::min<int>(max_rows, (int)s->s.iter_log.size() / 3);
    std::memcpy(out, s->s.iter_log.data(), (size_t)rows * 3 * sizeof(double));
    return rows;
}
int dotgpu_stepper_get_matrix(dotgpu_stepper* s, int sub, double* a_out) {
    API_BEGIN
    DG_REQUIRE(s && a_out, "null argument");
    Stepper& S = s->s;
    DG_CUDA(cudaSetDevice(S.cfg.device));
    int slot = -1;
    if (sub < 0) slot = 0;
    else
        for (size_t i = 0; i < S.owned.size(); ++i)
            if (S.owned[i] == sub) slot = 1 + (int)i;
    DG_REQUIRE(slot >= 0, "subdomain not owned by this rank");
    int64_t cnt = S.a_off[slot + 1] - S.a_off[slot];
    DG_CUDA(cudaMemcpyAsync(a_out, S.a_all.p + S.a_off[slot], cnt * sizeof(double), cudaMemcpyDeviceToHost, S.st));
    DG_CUDA(cudaStreamSynchronize(S.st));
    API_END
}
int dotgpu_stepper_get_dd(dotgpu_stepper* s, dotgpu_dd** dd_out) {
    API_BEGIN
    DG_REQUIRE(s && dd_out, "null argument");
    // expose a borrowed view: copy is cheap relative to set-up and keeps ownership simple
    dotgpu_dd* d = new dotgpu_dd();
    d->dd = s->s.dd;
    *dd_out = d;
    API_END
}
int dotgpu_stepper_precondition(dotgpu_stepper* s, const double* q, double* p_out) {
    API_BEGIN
    DG_REQUIRE(s && q && p_out, "null argument");
    Stepper& S = s->s;
    DG_CUDA(cudaSetDevice(S.cfg.device));
    const size_t n3 = 3 * (size_t)S.nV;
    S.q.upload(q, n3, S.st);
    S.precondition_dev(S.q.p, S.p.p);
    S.p.download(p_out, n3, S.st);
    API_END
}
int dotgpu_stepper_eval(dotgpu_stepper* s, const double* x, double* E_out, double* g_out) {
    API_BEGIN
    DG_REQUIRE(s && x, "null argument");
    Stepper& S = s->s;
    DG_CUDA(cudaSetDevice(S.cfg.device));
    const size_t n3 = 3 * (size_t)S.nV;
    S.x0.upload(x, n3, S.st);
    if (E_out) *E_out = S.energy_at(S.x0.p);
    if (g_out) {
        S.gradient_at(S.x0.p, S.q.p);
        S.q.download(g_out, n3, S.st);
    }
    API_END
}
int dotgpu_stepper_get_target(dotgpu_stepper* s, double* target) {
    if (!s || !target) return DOTGPU_ERR_INVALID;
    *target = s->s.target;
    return DOTGPU_OK;
}
int dotgpu_stepper_set_rel_tol(dotgpu_stepper* s, double rel_tol) {
    if (!s || !(rel_tol > 0.0)) return DOTGPU_ERR_INVALID;
    s->s.cfg.rel_tol = rel_tol;
    s->s.target = s->s.target_per_tolsq * rel_tol * rel_tol;  // O(1): the O(nT) geometry factor was computed at create time
    return DOTGPU_OK;
}
int dotgpu_stepper_time_kernels(dotgpu_stepper* s, int which, int reps, double* ms_out) {
    API_BEGIN
    DG_REQUIRE(s && ms_out, "null argument");
    DG_CUDA(cudaSetDevice(s->s.cfg.device));
    *ms_out = s->s.time_kernels(which, reps);
    API_END
}
int dotgpu_stepper_get_fill_stats(dotgpu_stepper* s, int64_t* nnz_out, int64_t* blocks_out, int64_t* gathered_blocks_out) {
    if (!s) return DOTGPU_ERR_INVALID;
    const Stepper& S = s->s;
    int64_t gathered = 0, blocks = 0;
    auto count = [&](const FillList& f) {
        blocks += (int64_t)f.ptr.size() - 1;
        for (int32_t c : f.src) gathered += c >= 0;
    };
    count(S.dd.gfill);
    for (int sd : S.owned) count(S.dd.subs[sd].fill);
    if (nnz_out) *nnz_out = S.a_off.empty() ? 0 : S.a_off.back();
    if (blocks_out) *blocks_out = blocks;
    if (gathered_blocks_out) *gathered_blocks_out = gathered;
    return DOTGPU_OK;
}
int dotgpu_stepper_get_owned(dotgpu_stepper* s, int32_t* out) {
    if (!s) return DOTGPU_ERR_INVALID;
    if (out)
        for (size_t i = 0; i < s->s.owned.size(); ++i) out[i] = s->s.owned[i];
    return (int)s->s.owned.size();
}
int dotgpu_stepper_get_solve_trace(dotgpu_stepper* s, uint64_t* out, int64_t max_words) {
    API_BEGIN
    DG_REQUIRE(s, "null argument");
    CholBatch& C = s->s.chol;
    const int64_t nw = (int64_t)C.d_trace.n;
    if (out && nw > 0) {
        DG_REQUIRE(max_words >= nw, "trace buffer too small");
        DG_CUDA(cudaSetDevice(s->s.cfg.device));
        DG_CUDA(cudaStreamSynchronize(s->s.st));
        DG_CUDA(cudaMemcpy(out, C.d_trace.p, (size_t)nw * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    return (int)std::min<int64_t>(nw, 1 << 30);
    API_END
}
int dotgpu_partition_nodes(int nV, int nT, const int32_t* tets, int k, int32_t* npart_out) {
    API_BEGIN
    metis_partition_nodes(nV, nT, tets, k, npart_out);
    API_END
}
int dotgpu_partition(int nV, int nT, const int32_t* tets, int k, int32_t* epart_out) {
    API_BEGIN
    metis_partition(nV, nT, tets, k, epart_out);
    API_END
}
int64_t dotgpu_stepper_launch_count(dotgpu_stepper* s) { return s ? g_launch_count - s->s.launches0 : 0; }
int dotgpu_stepper_get_solver_info(dotgpu_stepper* s, int sub, dotgpu_solver_info* info) {
    API_BEGIN
    DG_REQUIRE(s && info, "null argument");
    Stepper& S = s->s;
    int slot = -1;
    for (size_t i = 0; i < S.owned.size(); ++i)
        if (S.owned[i] == sub) slot = (int)i;
    DG_REQUIRE(slot >= 0, "subdomain not owned by this rank");
    fill_info(S.chol.sym[slot], S.dd.subs[sub].pat.nnz(), S.chol.device_bytes(), info);
    API_END
}

}  // extern "C"
