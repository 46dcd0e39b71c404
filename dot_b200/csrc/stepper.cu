#include "stepper.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "comm.h"

namespace dotgpu {

namespace {
__global__ void k_div_dup(int ndof, const int* __restrict__ dup, double* __restrict__ p) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= ndof) return;
    int du = dup[d / 3];
    if (du > 1) p[d] /= (double)du;
}
__global__ void k_set_rows(int count, const int* __restrict__ idx, const double* __restrict__ pos, double* __restrict__ x) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int v = idx[i];
    x[3 * (size_t)v] = pos[3 * i];
    x[3 * (size_t)v + 1] = pos[3 * i + 1];
    x[3 * (size_t)v + 2] = pos[3 * i + 2];
}
}  // namespace

std::vector<int> balanced_owner(const std::vector<double>& weight, int world) {
    const int k = (int)weight.size();
    std::vector<int> order(k), own(k, 0);
    for (int s = 0; s < k; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight[a] > weight[b]; });
    std::vector<double> load(world, 0.0);
    for (int s : order) {
        int best = 0;
        for (int r = 1; r < world; ++r)
            if (load[r] < load[best]) best = r;
        own[s] = best;
        load[best] += weight[s];
    }
    return own;
}

Stepper::~Stepper() {
    if (h_sc) cudaFreeHost(h_sc);
    if (h_sc2) cudaFreeHost(h_sc2);
    for (auto& e : ev_it)
        if (e) cudaEventDestroy(e);
    if (h_x) cudaFreeHost(h_x);
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : pc_ev) cudaEventDestroy(e);
    comm.reset();
    if (st) cudaStreamDestroy(st);
}

// ||d2Psi/dF2 (I)||_F^2 without projection (Optimizer.cpp:613-628 via compute_dP_div_dF at the identity)
static double rest_hessian_sqnorm(int energy, double mu, double lam) {
    double Ad, Ao, l, r;
    if (energy == DOTGPU_ENERGY_FCR) {
        Ad = 2.0 * mu + lam;
        Ao = lam * (1.0 * (1.0 - 1.0) + 1.0);
        l = mu - lam / 2.0 * 1.0 * (1.0 - 1.0);
        double dE = 2.0 * mu * 0.0 + 1.0 * lam * 0.0;
        r = (dE + dE) / (2.0 * 2.0);
    } else {
        double alpha = 1.0 + mu / lam;
        Ad = mu + lam;
        Ao = 1.0 * lam * (2.0 - alpha);
        double t0 = lam * (1.0 - alpha);
        l = (mu - t0 * 1.0) / 2.0;
        double dE = 1.0 * mu + t0 * 1.0;
        r = (dE + dE) / (2.0 * 2.0);
    }
    double s = 3.0 * Ad * Ad + 6.0 * Ao * Ao;
    s += 3.0 * (2.0 * (l + r) * (l + r) + 2.0 * (l - r) * (l - r));
    return s;
}

double Stepper::compute_target() const {
    // Optimizer::computeCharNormSq (Optimizer.cpp:613-651), evaluated on data0 whose fixedVert is {0}
    const double mu = cfg.YM / 2.0 / (1.0 + cfg.PR), lam = cfg.YM * cfg.PR / (1.0 + cfg.PR) / (1.0 - 2.0 * cfg.PR);
    std::vector<double> ls(nV, 0.0);
    for (int t = 0; t < nT; ++t) {
        const int32_t* T = tets_h.data() + 4 * (size_t)t;
        for (int j = 0; j < 4; ++j) {  // face opposite vertex j (igl::face_areas)
            int o[3], c = 0;
            for (int i = 0; i < 4; ++i)
                if (i != j) o[c++] = T[i];
            const double *a = &V_rest[3 * (size_t)o[0]], *b = &V_rest[3 * (size_t)o[1]], *cc = &V_rest[3 * (size_t)o[2]];
            double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {cc[0] - a[0], cc[1] - a[1], cc[2] - a[2]};
            double cr[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            ls[T[j]] += 0.5 * std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
        }
    }
    double sq = 0.0;
    for (double v : ls) sq += v * v;
    const double dt = cfg.dt;
    double t = cfg.rel_tol * cfg.rel_tol * rest_hessian_sqnorm(cfg.energy_type, mu, lam) * sq * (nV - cfg.target_fixed_count) / nV;
    return t * (dt * dt) * (dt * dt);
}

// Everything that depends on the Dirichlet set: domain decomposition patterns (fixed rows are identity rows, LinSysSolver.hpp:114-132),
// matrix fill lists, symbolic analysis of the owned subdomain matrices, the gather / scatter maps of the preconditioner.
// Runs at creation and again from set_fixed (DOTTimeStepper::updatePrecondMtrAndFactorize, DOTTimeStepper.cpp:185-231).
void Stepper::setup_decomposition() {
    // ---- domain decomposition, owned subdomains ----
    const int k = cfg.num_subdomains;
    std::vector<char> mask(k, 0);
    if (!owner_locked) owner.assign(k, 0);
    if (cfg.world > 1 && !owner_locked) {
        // balance the ranks by sum nnz(L_s): patterns + symbolic analysis of EVERY subdomain (host, OpenMP),
        // then the same deterministic greedy map on every rank.  DOTGPU_BALANCE=0: plain round-robin.
        const char* e = std::getenv("DOTGPU_BALANCE");
        if (e && *e == '0') {
            for (int s = 0; s < k; ++s) owner[s] = s % cfg.world;
        } else {
            DDHost probe;
            probe.build(nV, nT, tets_h.data(), epart_h.data(), k, fixed_h.data(), nullptr, 0.0, nullptr, false, nullptr);
            std::vector<double> wgt(k, 0.0);
#pragma omp parallel for schedule(dynamic)
            for (int s = 0; s < k; ++s) {
                Symbolic S;
                S.analyze(probe.subs[s].pat.n(), probe.subs[s].pat.ia.data(), probe.subs[s].pat.ja.data(), 21);
                wgt[s] = (double)S.nnz_l;
            }
            owner = balanced_owner(wgt, cfg.world);
        }
    }
    owned.clear();
    for (int s = 0; s < k; ++s)
        if (owner[s] == cfg.rank) owned.push_back(s);
    for (int s : owned) mask[s] = 1;
    if (lbfgs_jh) dd.build_node_blocks(nV, nT, tets_h.data(), npart_h.data(), k, fixed_h.data(), mass_h.data());
    else dd.build(nV, nT, tets_h.data(), epart_h.data(), k, fixed_h.data(), V_rest.data(), cfg.rho, mass_h.data(), true, &mask);
    if (cfg.world > 1 && mesh_own.nT == 0) {
        // energy / gradient are sharded by tet: this rank evaluates the tets of its own subdomains (the element partition is
        // disjoint, DOTTimeStepper.cpp:406-450 / SURVEY 8(e)); the sums over the ranks are one all-reduce per evaluation
        std::vector<int32_t> to;
        std::vector<double> Do, vo, muo, lo;
        for (int t = 0; t < nT; ++t)
            if (mask[epart_h[t]]) {
                to.insert(to.end(), tets_h.begin() + 4 * (size_t)t, tets_h.begin() + 4 * (size_t)t + 4);
                Do.insert(Do.end(), DmInv_h.begin() + 9 * (size_t)t, DmInv_h.begin() + 9 * (size_t)t + 9);
                vo.push_back(vol_h[t]); muo.push_back(mu_h[t]); lo.push_back(lam_h[t]);
            }
        DG_REQUIRE(!vo.empty(), "a rank without subdomains (more GPUs than subdomains)");
        mesh_own.init(cfg.energy_type, nV, (int)vo.size(), to.data(), Do.data(), vo.data(), muo.data(), lo.data(), mass_h.data(), fixed_h.data(), st);
    }

    // ---- concatenated matrices: [global | owned subdomains] ----
    const int nm = 1 + (int)owned.size();
    a_off.assign(nm + 1, 0);
    a_off[1] = dd.gpat.nnz();
    for (size_t i = 0; i < owned.size(); ++i) a_off[2 + i] = a_off[1 + i] + dd.subs[owned[i]].pat.nnz();
    DG_REQUIRE(a_off[nm] < (1LL << 31), "too many matrix entries for int32 slots");
    a_all.alloc(a_off[nm]);
    g_ia.upload(dd.gpat.ia, st);
    g_ja.upload(dd.gpat.ja, st);
    {
        std::vector<long long> ptr(1, 0);
        std::vector<int> src, row, bj, ia;
        std::vector<double> consts;
        auto add = [&](const MatrixPattern& P, const FillList& F, int64_t aoff) {
            const int ia_base = (int)ia.size();
            for (int32_t v : P.ia) ia.push_back((int)(v + aoff));
            const int cbase = (int)consts.size();
            consts.insert(consts.end(), F.consts.begin(), F.consts.end());
            const long long sbase = (long long)src.size();
            for (int32_t code : F.src) src.push_back(code >= 0 ? code : code - cbase);
            for (size_t b = 1; b < F.ptr.size(); ++b) ptr.push_back(sbase + F.ptr[b]);
            for (int v = 0; v < P.nverts; ++v)
                for (int b = P.bptr[v]; b < P.bptr[v + 1]; ++b) {
                    row.push_back(ia_base + 3 * v);
                    bj.push_back(P.fixed[v] ? -1 : b - P.bptr[v]);
                }
        };
        add(dd.gpat, dd.gfill, a_off[0]);
        for (size_t i = 0; i < owned.size(); ++i) add(dd.subs[owned[i]].pat, dd.subs[owned[i]].fill, a_off[1 + i]);
        fill.nblk = (long long)row.size();
        DG_REQUIRE((long long)ptr.size() == fill.nblk + 1, "fill list size mismatch");
        fill.ptr.upload(ptr, st);
        fill.src.upload(src, st);
        fill.row.upload(row, st);
        fill.j.upload(bj, st);
        fill.consts.upload(consts, st);
        fill.ia.upload(ia, st);
    }

    // ---- symbolic analysis of the owned subdomain matrices ----
    {
        std::vector<const int32_t*> ia, ja;
        std::vector<int> n;
        for (int s : owned) {
            ia.push_back(dd.subs[s].pat.ia.data());
            ja.push_back(dd.subs[s].pat.ja.data());
            n.push_back(dd.subs[s].pat.n());
        }
        chol.analyze(ia, ja, n, 21, st);
    }
    // ---- preconditioner gather / scatter maps ----
    {
        std::vector<int> gi(chol.n_total);
        std::vector<std::vector<int>> copies(3 * (size_t)nV);
        for (size_t i = 0; i < owned.size(); ++i) {
            const SubdomainHost& sd = dd.subs[owned[i]];
            const Symbolic& S = chol.sym[i];
            for (int pnew = 0; pnew < S.n; ++pnew) {
                int old = S.perm[pnew];
                int gd = 3 * sd.l2g[old / 3] + old % 3;
                gi[chol.col_off[i] + pnew] = gd;
                copies[gd].push_back((int)(chol.col_off[i] + pnew));
            }
        }
        std::vector<int> cp(3 * (size_t)nV + 1, 0), ci;
        ci.reserve(chol.n_total);
        for (size_t d = 0; d < copies.size(); ++d) {
            ci.insert(ci.end(), copies[d].begin(), copies[d].end());  // ascending subdomain order (DOTTimeStepper.cpp:436-445)
            cp[d + 1] = (int)ci.size();
        }
        if (ci.empty()) ci.push_back(0);
        gidx.upload(gi, st);
        cptr.upload(cp, st);
        cidx.upload(ci, st);
        std::vector<int> du(dd.dup.begin(), dd.dup.end());
        dup.upload(du, st);
    }
}

// DOTTimeStepper::updatePrecondMtrAndFactorize (DOTTimeStepper.cpp:185-270): the scripted Dirichlet set changed (AnimScripter returns 1,
// Optimizer.cpp:334-336) -> new patterns + symbolic analysis per subdomain, then Hessians and factorisation at the current positions.
void Stepper::set_fixed(const uint8_t* fixed_mask, const double* x_eval) {
    DG_REQUIRE(fixed_mask != nullptr, "null fixed mask");
    fixed_h.assign(fixed_mask, fixed_mask + nV);
    mesh.set_fixed(fixed_h.data(), st);
    owner_locked = true;   // the subdomain -> rank map stays what it was: factors move only inside a rank
    setup_decomposition();
    if (cfg.world > 1) mesh_own.set_fixed(fixed_h.data(), st);
    const double dtsq = cfg.dt * cfg.dt;
    launch_xtilde(nV, xt.p, xn.p, vel.p, mesh.fixed.p, cfg.dt, cfg.gravity[0] * dtsq, cfg.gravity[1] * dtsq, cfg.gravity[2] * dtsq, st);
    // the reference evaluates the new preconditioner at result.V, i.e. with this frame's scripted move already applied
    if (x_eval) x.upload(x_eval, 3 * (size_t)nV, st);
    else DG_CUDA(cudaMemcpyAsync(x.p, xn.p, 3 * (size_t)nV * sizeof(double), cudaMemcpyDeviceToDevice, st));
    refresh();
    chol.check_status(st);
}

void Stepper::create(const dotgpu_stepper_config& c, int nV_, int nT_, const double* Vr, const int32_t* T, const int32_t* ep,
                     const uint8_t* fixed_mask) {
    cfg = c;
    nV = nV_;
    nT = nT_;
    DG_REQUIRE(nV > 0 && nT > 0 && Vr && T, "null or empty mesh");
    DG_REQUIRE(ep || (c.flags & (DOTGPU_FLAG_LBFGS_H | DOTGPU_FLAG_LBFGS_JH | DOTGPU_FLAG_NEWTON)), "element labels missing");
    // history + 1 (S, Y) buffers are in use (the candidate pair lives in a spare slot), and the device scalar table has
    // LB_MAXH slots per row (linalg.h) -> at most LB_MAXH - 1 pairs
    static_assert(SC_YP - SC_SG == LB_MAXH && SC_XI - SC_YP == LB_MAXH && SC_SY - SC_XI == LB_MAXH && SC_COUNT == SC_SY + LB_MAXH * LB_MAXH,
                  "scalar table layout assumes LB_MAXH slots");
    DG_REQUIRE(cfg.num_subdomains >= 1 && cfg.history >= 0 && cfg.history <= LB_MAXH - 1, "bad subdomain count / history size (max 7 pairs)");
    DG_REQUIRE(cfg.world >= 1 && cfg.rank >= 0 && cfg.rank < cfg.world, "bad rank/world");
    DG_REQUIRE(cfg.dt > 0, "dt must be positive");
    newton = (cfg.flags & DOTGPU_FLAG_NEWTON) != 0;
    if (const char* e = std::getenv("DOTGPU_DEBUG_LS_FAIL")) debug_ls_fail = *e == '1';
    if (newton) {
        DG_REQUIRE(cfg.num_subdomains == 1 && cfg.world == 1, "Projected Newton runs on one subdomain (the whole mesh) and one GPU");
        cfg.history = 0;  // no quasi-Newton pairs: p = -H(x)^-1 g with the Hessian at the current iterate
    }
    lbfgs_h = (cfg.flags & DOTGPU_FLAG_LBFGS_H) != 0;
    lbfgs_jh = (cfg.flags & DOTGPU_FLAG_LBFGS_JH) != 0;
    DG_REQUIRE((int)newton + (int)lbfgs_h + (int)lbfgs_jh <= 1, "DOTGPU_FLAG_NEWTON / LBFGS_H / LBFGS_JH exclude each other");
    if (lbfgs_h) DG_REQUIRE(cfg.num_subdomains == 1 && cfg.world == 1, "LBFGS-H factorises ONE global matrix: num_subdomains = 1, one GPU");
    if (lbfgs_jh) DG_REQUIRE(cfg.node_part != nullptr && cfg.num_subdomains >= 2 && cfg.world == 1, "LBFGS-JH needs node labels, k >= 2, one GPU");
    unit_step = newton || lbfgs_h || lbfgs_jh;  // Optimizer::initStepSize: only TST_DOT starts from -p.g / p.Hp (Optimizer.cpp:1076-1093)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw Error(DOTGPU_ERR_NO_DEVICE, "no CUDA device");
    DG_REQUIRE(cfg.device >= 0 && cfg.device < ndev, "device index out of range");
    DG_CUDA(cudaSetDevice(cfg.device));
    DG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto& e : ev) DG_CUDA(cudaEventCreate(&e));
    launches0 = g_launch_count;
    V_rest.assign(Vr, Vr + 3 * (size_t)nV);
    tets_h.assign(T, T + 4 * (size_t)nT);
    if (ep && !(c.flags & (DOTGPU_FLAG_LBFGS_H | DOTGPU_FLAG_LBFGS_JH))) epart_h.assign(ep, ep + nT);
    else epart_h.assign(nT, 0);   // one "subdomain" = the whole mesh (LBFGS-H, Newton); unused by LBFGS-JH
    if (c.flags & DOTGPU_FLAG_LBFGS_JH) {
        DG_REQUIRE(c.node_part != nullptr, "LBFGS-JH needs node labels");
        npart_h.assign(c.node_part, c.node_part + nV);
    }
    fixed_h.assign(nV, 0);
    if (fixed_mask) fixed_h.assign(fixed_mask, fixed_mask + nV);

    // ---- mesh features + upload ----
    DmInv_h.assign(9 * (size_t)nT, 0.0);
    vol_h.assign(nT, 0.0);
    mu_h.assign(nT, 0.0);
    lam_h.assign(nT, 0.0);
    mass_h.assign(nV, 0.0);
    mesh_features(nV, nT, V_rest.data(), tets_h.data(), cfg.YM, cfg.PR, cfg.rho, DmInv_h.data(), vol_h.data(), mass_h.data(), mu_h.data(),
                  lam_h.data());
    for (int t = 0; t < nT; ++t) DG_REQUIRE(vol_h[t] > 0.0, "inverted or degenerate rest tet");
    mesh.init(cfg.energy_type, nV, nT, tets_h.data(), DmInv_h.data(), vol_h.data(), mu_h.data(), lam_h.data(), mass_h.data(), fixed_h.data(), st);

    setup_decomposition();
    // ---- vectors ----
    const size_t n3 = 3 * (size_t)nV;
    for (DevBuf<double>* b : {&x, &x0, &xn, &xt, &vel, &g, &g_old, &q, &p}) {
        b->alloc(n3 + 1);   // + 1: the energy rides behind the gradient in the multi-GPU all-reduce
        b->zero(st);
    }
    xperm.alloc(std::max<int64_t>(chol.n_total, 1));
    qf_partial.alloc(ceil_div((long long)n3, 32) + 1);  // one partial per CTA of k_quadform_alpha (32 rows each)
    dot_partial.alloc(dot_partial_count((long long)n3));
    md_partial.alloc(multidot_partial_count());
    S.resize(cfg.history + 1);
    Y.resize(cfg.history + 1);
    for (int i = 0; i <= cfg.history; ++i) {
        S[i].alloc(n3);
        Y[i].alloc(n3);
    }
    sc.alloc(SC_COUNT);
    sc.zero(st);
    counter.alloc(1);
    counter.zero(st);
    DG_CUDA(cudaMallocHost((void**)&h_sc, SC_COUNT * sizeof(double)));
    DG_CUDA(cudaMallocHost((void**)&h_x, n3 * sizeof(double)));
    if (cfg.world > 1) {
        comm.reset(new Comm());
        comm->init(cfg.nccl_unique_id, cfg.rank, cfg.world);
        comm->enable_peer((long long)n3 + 1, st);  // [g ; E] and p go through NVLink peer memory (peer_reduce.cu) when every rank can map it
    }
    target = compute_target();
    target_per_tolsq = target / (cfg.rel_tol * cfg.rel_tol);  // the geometry / material factor: set_rel_tol only rescales it
    // ---- precompute(): rest-state energy, Hessians, factorisation (DOTTimeStepper.cpp:150-182) ----
    x.upload(V_rest.data(), n3, st);
    DG_CUDA(cudaMemcpyAsync(xn.p, x.p, n3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    launch_xtilde(nV, xt.p, xn.p, vel.p, mesh.fixed.p, cfg.dt, cfg.gravity[0] * cfg.dt * cfg.dt, cfg.gravity[1] * cfg.dt * cfg.dt,
                  cfg.gravity[2] * cfg.dt * cfg.dt, st);
    refresh();
    chol.check_status(st);
}

void Stepper::fetch_scalars(int first, int count) {
    DG_CUDA(cudaMemcpyAsync(h_sc + first, sc.p + first, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaStreamSynchronize(st));
    DG_CUDA(cudaGetLastError());
}

double Stepper::energy_at(const double* x_dev) {
    if (cfg.world > 1) {  // partial energy of the owned tets (+ the inertia term on rank 0), summed over the ranks
        launch_energy(mesh_own, x_dev, cfg.rank == 0 ? xt.p : nullptr, cfg.dt * cfg.dt, sc.p + SC_E, st);
        comm->all_reduce_sum(sc.p + SC_E, 1, st);
    } else {
        launch_energy(mesh, x_dev, xt.p, cfg.dt * cfg.dt, sc.p + SC_E, st);
    }
    fetch_scalars(SC_E, 1);
    return h_sc[SC_E];
}

void Stepper::gradient_at(const double* x_dev, double* g_dev) {
    if (cfg.world > 1) {
        launch_gradient(mesh_own, x_dev, cfg.rank == 0 ? xt.p : nullptr, cfg.dt * cfg.dt, g_dev, st);
        comm->all_reduce_sum(g_dev, 3LL * nV, st);
    } else {
        launch_gradient(mesh, x_dev, xt.p, cfg.dt * cfg.dt, g_dev, st);
    }
}

void Stepper::eval_sharded(const double* x_dev, double* G) {
    const long long n3 = 3LL * nV;
    const double* xtp = cfg.rank == 0 ? xt.p : nullptr;
    launch_energy(mesh_own, x_dev, xtp, cfg.dt * cfg.dt, G + n3, st);
    launch_gradient(mesh_own, x_dev, xtp, cfg.dt * cfg.dt, G, st);
    comm->all_reduce_sum(G, n3 + 1, st);
}

// first half only: this rank's [g ; E] goes into its slot on every rank; the consumer (k_pair_dots) adds the slots.  Returns false
// when the peer path is not available (then G holds the reduced vector, as after eval_sharded)
bool Stepper::eval_sharded_push(const double* x_dev, double* G) {
    if (!comm->peer) {
        eval_sharded(x_dev, G);
        return false;
    }
    const long long n3 = 3LL * nV;
    const double* xtp = cfg.rank == 0 ? xt.p : nullptr;
    (void)n3;
    (void)G;
    launch_gradient_push(mesh_own, x_dev, xtp, cfg.dt * cfg.dt, comm->peer->begin(), st);  // K1 + K2 + push in two launches
    comm->peer->after_push(st);
    return true;
}

void Stepper::refresh() {
    launch_elem_hessians(mesh, x.p, cfg.dt * cfg.dt, true, st);
    launch_fill(fill, mesh.He.p, a_all.p, st);
    if (!owned.empty()) chol.factorize(a_all.p + a_off[1], st);
}

bool Stepper::precondition_dev(const double* q_dev, double* p_dev, const DotPairs* fuse) {
    const int ndof = 3 * nV;
    if (chol.n_total > 0) chol.solve(q_dev, gidx.p, xperm.p, st, fuse ? fuse->go : nullptr);
    if (cfg.world == 1) {
        if (fuse) {
            launch_scatter_avg_dots(ndof, cptr.p, cidx.p, xperm.p, dup.p, p_dev, *fuse, md_partial.p, counter.p, sc.p, st);
            return true;
        }
        launch_scatter_avg(ndof, cptr.p, cidx.p, xperm.p, dup.p, p_dev, st);
    } else {
        if (comm->peer && fuse) {
            // all-reduce over peer memory, both halves fused: the scatter kernel stores this rank's sums into every rank's slots
            // and publishes the epoch; the multi-dot kernel waits for the flags, adds the slots in rank order, divides by dup
            launch_scatter_push(ndof, cptr.p, cidx.p, xperm.p, comm->peer->begin(), st);
            comm->peer->after_push(st);
            const PeerSrc src = comm->peer->src();
            launch_divdup_dots(ndof, dup.p, p_dev, *fuse, md_partial.p, counter.p, sc.p, st, &src);
            return true;
        }
        launch_scatter_avg(ndof, cptr.p, cidx.p, xperm.p, nullptr, p_dev, st);
        comm->all_reduce_sum(p_dev, ndof, st);
        if (fuse) {  // division by dup fused with the second multi-dot of the iteration
            launch_divdup_dots(ndof, dup.p, p_dev, *fuse, md_partial.p, counter.p, sc.p, st);
            return true;
        }
        k_div_dup<<<ceil_div(ndof, 256), 256, 0, st>>>(ndof, dup.p, p_dev);
        count_launch();
    }
    return false;
}

void Stepper::frame(double* x_inout, dotgpu_frame_stats* stats) {
    const size_t n3 = 3 * (size_t)nV;
    DG_CUDA(cudaEventRecord(ev[0], st));
    std::memcpy(h_x, x_inout, n3 * sizeof(double));
    DG_CUDA(cudaMemcpyAsync(x.p, h_x, n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    frame_core(stats, true);
    std::memcpy(x_inout, h_x, n3 * sizeof(double));
}

void Stepper::frame_resident(const int32_t* idx, const double* pos, int count, dotgpu_frame_stats* stats) {
    const size_t n3 = 3 * (size_t)nV;
    DG_REQUIRE(count >= 0 && (count == 0 || (idx && pos)), "bad Dirichlet target list");
    for (int i = 0; i < count; ++i) DG_REQUIRE(idx[i] >= 0 && idx[i] < nV, "Dirichlet vertex out of range");
    DG_CUDA(cudaEventRecord(ev[0], st));
    // x = x^n (resident) with the scripted rows overwritten
    DG_CUDA(cudaMemcpyAsync(x.p, xn.p, n3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (count > 0) {
        if (h_idx.n < (size_t)count) { h_idx.alloc(count); h_pos.alloc(3 * (size_t)count); }
        DG_CUDA(cudaMemcpyAsync(h_idx.p, idx, count * sizeof(int), cudaMemcpyHostToDevice, st));
        DG_CUDA(cudaMemcpyAsync(h_pos.p, pos, 3 * (size_t)count * sizeof(double), cudaMemcpyHostToDevice, st));
        k_set_rows<<<ceil_div(count, 128), 128, 0, st>>>(count, h_idx.p, h_pos.p, x.p);
        count_launch();
    }
    frame_core(stats, false);
}

HistList Stepper::hist_list() const {
    HistList H;
    H.go = nullptr;
    H.n = (int)hist.size();
    for (int i = 0; i < LB_MAXH; ++i) {
        H.slot[i] = 0;
        H.S[i] = H.Y[i] = nullptr;
    }
    for (int i = 0; i < H.n; ++i) {
        H.slot[i] = hist[i];
        H.S[i] = S[hist[i]].p;
        H.Y[i] = Y[hist[i]].p;
    }
    return H;
}

void Stepper::frame_core(dotgpu_frame_stats* stats, bool copy_back) {
    const size_t n3 = 3 * (size_t)nV;
    const long long n = (long long)n3;
    const double dt = cfg.dt, dtsq = dt * dt;
    hist.clear();
    iter_log.clear();
    int halvings = 0, evals = 0;
    // initX(warmStart = 2)
    launch_warm_start(nV, x.p, vel.p, mesh.fixed.p, dt, cfg.gravity[0] * dtsq, cfg.gravity[1] * dtsq, cfg.gravity[2] * dtsq, st);
    if (cfg.world > 1) {
        eval_sharded(x.p, g.p);
        DG_CUDA(cudaMemcpyAsync(sc.p + SC_E, g.p + n3, sizeof(double), cudaMemcpyDeviceToDevice, st));
    } else {
        launch_energy(mesh, x.p, xt.p, dtsq, sc.p + SC_E, st);
        gradient_at(x.p, g.p);
    }
    ++evals;
    launch_dot(n, g.p, g.p, dot_partial.p, counter.p, sc.p + SC_GG, st);
    fetch_scalars(SC_E, 2);
    double E = h_sc[SC_E], gg = h_sc[SC_GG];
    iter_log.insert(iter_log.end(), {0.0, E, gg});
    int iters = 0;
    bool sg_valid = false, stopped = false;
    std::vector<int> free_slots;
    for (int i = 0; i <= cfg.history; ++i) free_slots.push_back(i);
    // ------------------------------------------------------------------------------------------------------------------------
    // Speculative iteration pipeline (one GPU, DOT / L-BFGS variants): iteration i+1 is enqueued BEFORE the host has seen the result of
    // iteration i, under the assumption that holds for almost every iteration - step accepted without halving, new pair kept, not
    // converged.  The last kernel of iteration i evaluates exactly that predicate on the device and writes a flag; every kernel of
    // iteration i+1 returns at once if it is 0.  The host then only CONFIRMS iteration i while the GPU already runs i+1: no host round
    // trip and no launch gaps inside the loop, and the arithmetic (hence every iteration log) is unchanged.  When the assumption fails
    // (a halving, a rejected pair, convergence - the latter once per time step) the host state is rolled back and the exceptional
    // case is handled on the synchronous path below.
    // ------------------------------------------------------------------------------------------------------------------------
    static const bool spec_env = !(std::getenv("DOTGPU_SPECULATE") && std::getenv("DOTGPU_SPECULATE")[0] == '0');
    const bool speculate = spec_env && cfg.world == 1 && !newton && cfg.history > 0 && !debug_ls_fail;
    if (speculate) {
        if (!h_sc2) {
            DG_CUDA(cudaMallocHost((void**)&h_sc2, 2 * SC_COUNT * sizeof(double)));
            DG_CUDA(cudaEventCreateWithFlags(&ev_it[0], cudaEventDisableTiming));
            DG_CUDA(cudaEventCreateWithFlags(&ev_it[1], cudaEventDisableTiming));
            spec_flags.alloc(2);
        }
        // energy of the accepted point lives on the device too
        DG_CUDA(cudaMemcpyAsync(sc.p + SC_EPREV, sc.p + SC_E, sizeof(double), cudaMemcpyDeviceToDevice, st));
        struct HostState {
            std::deque<int> hist;
            std::vector<int> free_slots;
            double *x, *x0, *g, *g_old;
            bool sg_valid;
        };
        auto save = [&]() { return HostState{hist, free_slots, x.p, x0.p, g.p, g_old.p, sg_valid}; };
        auto restore = [&](const HostState& h) {
            hist = h.hist; free_slots = h.free_slots; x.p = h.x; x0.p = h.x0; g.p = h.g; g_old.p = h.g_old; sg_valid = h.sg_valid;
        };
        int enq = 0;  // iterations enqueued so far in this time step
        // enqueue one iteration on the current host state; returns the candidate slot of its new pair
        auto enqueue = [&](const int* go, int* flag_out) -> int {
            HistList H = hist_list();
            H.go = go;
            if (H.n > 0 && !sg_valid) {
                DotPairs P;
                P.go = go;
                P.n = H.n;
                for (int i = 0; i < H.n; ++i) { P.a[i] = H.S[i]; P.b[i] = g.p; P.out[i] = SC_SG + H.slot[i]; }
                launch_dots(n, P, md_partial.p, counter.p, sc.p, st);
            }
            launch_lbfgs_q(n, q.p, g.p, H, sc.p, st);
            if (pc_ev.size() < 2 * (size_t)(enq + 1)) {
                cudaEvent_t a, b;
                DG_CUDA(cudaEventCreate(&a));
                DG_CUDA(cudaEventCreate(&b));
                pc_ev.push_back(a);
                pc_ev.push_back(b);
            }
            DotPairs P;
            P.go = go;
            P.n = H.n + 1;
            for (int i = 0; i < H.n; ++i) { P.a[i] = H.Y[i]; P.b[i] = p.p; P.out[i] = SC_YP + H.slot[i]; }
            P.a[H.n] = g.p; P.b[H.n] = p.p; P.out[H.n] = SC_P0G;
            DG_CUDA(cudaEventRecord(pc_ev[2 * enq], st));
            precondition_dev(q.p, p.p, &P);
            DG_CUDA(cudaEventRecord(pc_ev[2 * enq + 1], st));
            launch_lbfgs_p(n, p.p, H, sc.p, st);
            const double* alpha_dev = unit_step ? nullptr : sc.p + SC_ALPHA;
            if (!unit_step) launch_quadform_alpha(3 * nV, g_ia.p, g_ja.p, a_all.p, p.p, qf_partial.p, counter.p, sc.p, st, go);
            std::swap(x.p, x0.p);
            launch_axpy_dev(n, x.p, x0.p, p.p, alpha_dev, 1.0, st, go);
            const int sl = free_slots.back();
            launch_gradient_pair(mesh, x.p, xt.p, dtsq, g_old.p, p.p, g.p, S[sl].p, Y[sl].p, sl, alpha_dev, 1.0, H, md_partial.p, counter.p, sc.p,
                                 true, st, flag_out, target);
            DG_CUDA(cudaMemcpyAsync(h_sc2 + (enq & 1) * SC_COUNT, sc.p, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st));
            DG_CUDA(cudaEventRecord(ev_it[enq & 1], st));
            ++enq;
            return sl;
        };
        // the host-side effect of "step accepted, pair kept"
        auto accept = [&](int sl) {
            std::swap(g.p, g_old.p);
            sg_valid = true;
            free_slots.pop_back();
            hist.push_back(sl);
            if ((int)hist.size() > cfg.history) {
                free_slots.push_back(hist.front());
                hist.pop_front();
            }
        };
        int pending_sl = enqueue(nullptr, spec_flags.p + 0);  // iteration 0 of the time step: nothing to speculate on
        ++evals;
        while (true) {
            const int cur = enq - 1;               // index of the iteration waiting for confirmation
            const HostState before = save();       // host state as iteration `cur` left it (pointers already swapped by enqueue)
            int next_sl = -1;
            const bool ahead = iters + 1 < cfg.max_iters;
            if (ahead) {
                accept(pending_sl);
                next_sl = enqueue(spec_flags.p + (cur & 1), spec_flags.p + ((cur + 1) & 1));
            }
            DG_CUDA(cudaEventSynchronize(ev_it[cur & 1]));
            const double* hs = h_sc2 + (cur & 1) * SC_COUNT;
            double alpha = unit_step ? 1.0 : hs[SC_ALPHA], Et = hs[SC_E];
            if (debug_ls_fail) Et = INFINITY;
            const bool ok = !debug_ls_fail && (Et <= E) && (hs[SC_YS_NEW] > 0.0) && (hs[SC_GG] > target);
            if (ok && ahead) {                     // the common case: iteration `cur` is confirmed, cur + 1 is already running
                E = Et;
                gg = hs[SC_GG];
                ++iters;
                ++evals;
                iter_log.insert(iter_log.end(), {alpha, E, gg});
                pending_sl = next_sl;
                continue;
            }
            // ---- exceptional: roll the host back to the state after iteration `cur` was enqueued; whatever was enqueued behind it
            //      sees flag 0 and does nothing (the device evaluated the same predicate on the same numbers) ----
            restore(before);
            DG_CUDA(cudaStreamSynchronize(st));
            DG_CUDA(cudaGetLastError());
            std::memcpy(h_sc, hs, SC_COUNT * sizeof(double));
            const int sl = pending_sl;
            const HistList H = hist_list();
            if (Et > E && alpha > 0.0) {  // back-tracking (Optimizer.cpp:803-833)
                while (true) {
                    alpha /= 2.0;
                    ++halvings;
                    if (alpha == 0.0) {
                        stopped = true;
                        break;
                    }
                    launch_axpy_dev(n, x.p, x0.p, p.p, nullptr, alpha, st);
                    Et = energy_at(x.p);
                    if (debug_ls_fail) Et = INFINITY;
                    ++evals;
                    if (!(Et > E)) break;
                }
                launch_gradient_pair(mesh, x.p, xt.p, dtsq, g_old.p, p.p, g.p, S[sl].p, Y[sl].p, sl, nullptr, alpha, H, md_partial.p, counter.p,
                                     sc.p, false, st);
                fetch_scalars(0, SC_COUNT);
            }
            E = Et;
            std::swap(g.p, g_old.p);
            gg = h_sc[SC_GG];
            sg_valid = true;
            if (h_sc[SC_YS_NEW] > 0.0) {
                free_slots.pop_back();
                hist.push_back(sl);
                if ((int)hist.size() > cfg.history) {
                    free_slots.push_back(hist.front());
                    hist.pop_front();
                }
            }
            if (stopped) break;
            ++iters;
            iter_log.insert(iter_log.end(), {alpha, E, gg});
            if (!(gg > target) || iters >= cfg.max_iters) break;
            // continue from the corrected state: the accepted energy goes back to the device, the next iteration is not speculative
            h_sc[SC_EPREV] = E;
            DG_CUDA(cudaMemcpyAsync(sc.p + SC_EPREV, h_sc + SC_EPREV, sizeof(double), cudaMemcpyHostToDevice, st));
            pending_sl = enqueue(nullptr, spec_flags.p + ((enq) & 1));
            ++evals;
        }
        spec_enq = enq;
    } else
    do {
        // ---- L-BFGS two-loop with the decomposed Hessian as initialiser (DOTTimeStepper.cpp:384-466), compact form: the inner
        //      products against the history are taken in two multi-dot passes, the recursions run on scalars ----
        if (newton) refresh();  // Optimizer::solve_oneStep: computePrecondMtr + factorize at the current iterate (Optimizer.cpp:703-730)
        const HistList H = hist_list();
        if (H.n > 0 && !sg_valid) {  // normally produced by the previous iteration's fused gradient kernel
            DotPairs P;
            P.go = nullptr;
            P.n = H.n;
            for (int i = 0; i < H.n; ++i) { P.a[i] = H.S[i]; P.b[i] = g.p; P.out[i] = SC_SG + H.slot[i]; }
            launch_dots(n, P, md_partial.p, counter.p, sc.p, st);
        }
        launch_lbfgs_q(n, q.p, g.p, H, sc.p, st);
        if (pc_ev.size() < 2 * (size_t)(iters + 1)) {
            cudaEvent_t a, b;
            DG_CUDA(cudaEventCreate(&a));
            DG_CUDA(cudaEventCreate(&b));
            pc_ev.push_back(a);
            pc_ev.push_back(b);
        }
        {
            DotPairs P;  // second multi-dot: y_i . p0 and g . p0, taken inside the scatter pass on one GPU
            P.go = nullptr;
            P.n = H.n + 1;
            for (int i = 0; i < H.n; ++i) { P.a[i] = H.Y[i]; P.b[i] = p.p; P.out[i] = SC_YP + H.slot[i]; }
            P.a[H.n] = g.p; P.b[H.n] = p.p; P.out[H.n] = SC_P0G;
            DG_CUDA(cudaEventRecord(pc_ev[2 * iters], st));
            const bool fused = precondition_dev(q.p, p.p, &P);
            DG_CUDA(cudaEventRecord(pc_ev[2 * iters + 1], st));
            if (!fused) launch_dots(n, P, md_partial.p, counter.p, sc.p, st);
        }
        launch_lbfgs_p(n, p.p, H, sc.p, st);
        // ---- initial step length (Optimizer.cpp:1076-1093), computed and consumed on the device ----
        const double* alpha_dev = unit_step ? nullptr : sc.p + SC_ALPHA;  // everything but DOT: initStepSize = 1 (Optimizer.cpp:1088)
        if (!unit_step) launch_quadform_alpha(3 * nV, g_ia.p, g_ja.p, a_all.p, p.p, qf_partial.p, counter.p, sc.p, st);
        // ---- back-tracking line search (Optimizer.cpp:752-881): the first trial and everything that follows an accepted
        //      step are enqueued without waiting; the host looks at the result once per iteration ----
        std::swap(x.p, x0.p);  // x0 = current positions
        launch_axpy_dev(n, x.p, x0.p, p.p, alpha_dev, 1.0, st);
        ++evals;
        // energy AND gradient at the trial point in one pass over the tets (g_old is the spare buffer until the step is accepted)
        // + new pair + the next iteration's dots
        const int sl = cfg.history > 0 ? free_slots.back() : -1;  // history+1 buffers: a candidate slot is always free
        const bool multi = cfg.world > 1;
        // one GPU: gradient + energy + pair + dots fused (k_grad_vertex_pair).  Several GPUs: each rank evaluates its own tets,
        // one all-reduce of [g ; E], then the pair and its dots from the reduced gradient (s_i . g follows at the next iteration's start)
        auto grad_pair = [&](const double* a_dev, double a_host, bool with_energy) {
            if (!multi) {
                launch_gradient_pair(mesh, x.p, xt.p, dtsq, g_old.p, p.p, g.p, sl >= 0 ? S[sl].p : nullptr, sl >= 0 ? Y[sl].p : nullptr, sl,
                                     a_dev, a_host, H, md_partial.p, counter.p, sc.p, with_energy, st);
                return;
            }
            if (eval_sharded_push(x.p, g_old.p)) {  // peer memory: the pair kernel is the second half of the all-reduce
                const PeerSrc src = comm->peer->src();
                launch_pair_dots(n, p.p, nullptr, g.p, sl >= 0 ? S[sl].p : nullptr, sl >= 0 ? Y[sl].p : nullptr, sl, a_dev, a_host, H,
                                 md_partial.p, counter.p, sc.p, st, &src, g_old.p, with_energy);
                return;
            }
            if (with_energy) DG_CUDA(cudaMemcpyAsync(sc.p + SC_E, g_old.p + n3, sizeof(double), cudaMemcpyDeviceToDevice, st));
            launch_pair_dots(n, p.p, g_old.p, g.p, sl >= 0 ? S[sl].p : nullptr, sl >= 0 ? Y[sl].p : nullptr, sl, a_dev, a_host, H, md_partial.p,
                             counter.p, sc.p, st);
        };
        grad_pair(alpha_dev, 1.0, true);
        fetch_scalars(0, SC_COUNT);  // the one host round trip of an iteration (measured: ~14 us of 240 on bar17K_like)
        double alpha = unit_step ? 1.0 : h_sc[SC_ALPHA], Et = h_sc[SC_E];
        // tests only (DOTGPU_DEBUG_LS_FAIL=1): every trial energy reads as +inf, so the step halves until it underflows.  (With real
        // energies that branch is nearly unreachable: x0 + alpha p rounds to x0 long before alpha reaches 0, and E(x0) > E(x0) is false.)
        if (debug_ls_fail) Et = INFINITY;
        if (Et > E && alpha > 0.0) {
            // rare: halve until the energy does not increase, then redo the gradient / pair at the accepted point
            while (true) {
                alpha /= 2.0;
                ++halvings;
                if (alpha == 0.0) {  // Optimizer.cpp:816-824: the step underflowed, the line search has failed
                    stopped = true;
                    break;
                }
                launch_axpy_dev(n, x.p, x0.p, p.p, nullptr, alpha, st);
                Et = energy_at(x.p);
                if (debug_ls_fail) Et = INFINITY;
                ++evals;
                if (!(Et > E)) break;
            }
            grad_pair(nullptr, alpha, false);
            fetch_scalars(0, SC_COUNT);
        }
        E = Et;
        std::swap(g.p, g_old.p);
        gg = h_sc[SC_GG];
        sg_valid = true;  // sc[SC_SG + slot] now holds s_i . g for every pair that can be in the next history (both paths)
        // ---- history update (DOTTimeStepper.cpp:476-493): keep the pair iff y.s > 0, then drop the oldest beyond `history` ----
        if (sl >= 0 && h_sc[SC_YS_NEW] > 0.0) {
            free_slots.pop_back();
            hist.push_back(sl);
            if ((int)hist.size() > cfg.history) {
                free_slots.push_back(hist.front());
                hist.pop_front();
            }
        }
        // a failed line search ends the time step at once, uncounted and without the Hessian refresh
        // (DOTTimeStepper.cpp:313-317 returns before innerIterAmt++ and before updateHessianAndFactor)
        if (stopped) break;
        ++iters;
        iter_log.insert(iter_log.end(), {alpha, E, gg});
    } while (gg > target && iters < cfg.max_iters);
    DG_CUDA(cudaEventRecord(ev[1], st));
    // ---- Hessian refresh at the end of the step (DOTTimeStepper.cpp:343, 349-380); Newton refactorises per iteration instead ----
    if (!newton && !stopped) refresh();
    DG_CUDA(cudaEventRecord(ev[2], st));
    // ---- BE update (Optimizer.cpp:354-361) ----
    launch_velocity(nV, vel.p, x.p, xn.p, dt, st);
    DG_CUDA(cudaMemcpyAsync(xn.p, x.p, n3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    launch_xtilde(nV, xt.p, xn.p, vel.p, mesh.fixed.p, dt, cfg.gravity[0] * dtsq, cfg.gravity[1] * dtsq, cfg.gravity[2] * dtsq, st);
    if (copy_back) DG_CUDA(cudaMemcpyAsync(h_x, x.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaEventRecord(ev[3], st));
    DG_CUDA(cudaStreamSynchronize(st));
    chol.check_status(st);
    E_last = E;
    if (stats) {
        stats->iters = iters;
        stats->halvings = halvings;
        stats->energy_evals = evals;
        stats->converged = !stopped && gg <= target;
        stats->E = E;
        stats->grad_sqnorm = gg;
        stats->target = target;
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, ev[0], ev[3]);
        cudaEventElapsedTime(&b, ev[0], ev[1]);
        cudaEventElapsedTime(&c, ev[1], ev[2]);
        stats->ms_total = a;
        stats->ms_solve = b;
        stats->ms_refresh = c;
        double pc = 0.0;
        for (int i = 0; i < (speculate ? spec_enq : iters); ++i) {
            float t = 0;
            cudaEventElapsedTime(&t, pc_ev[2 * i], pc_ev[2 * i + 1]);
            pc += t;
        }
        stats->ms_precond = pc;
        stats->precond_calls = iters;
        stats->line_search_failed = stopped ? 1 : 0;
    }
}

void Stepper::set_state(const double* xs, const double* velocity) {
    const size_t n3 = 3 * (size_t)nV;
    x.upload(xs, n3, st);
    DG_CUDA(cudaMemcpyAsync(xn.p, x.p, n3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (velocity) vel.upload(velocity, n3, st);
    else vel.zero(st);
    const double dtsq = cfg.dt * cfg.dt;
    launch_xtilde(nV, xt.p, xn.p, vel.p, mesh.fixed.p, cfg.dt, cfg.gravity[0] * dtsq, cfg.gravity[1] * dtsq, cfg.gravity[2] * dtsq, st);
    refresh();
    chol.check_status(st);
}

void Stepper::get_state(double* xs, double* velocity, double* xTilde) {
    const size_t n3 = 3 * (size_t)nV;
    if (xs) xn.download(xs, n3, st);
    if (velocity) vel.download(velocity, n3, st);
    if (xTilde) xt.download(xTilde, n3, st);
}

double Stepper::time_kernels(int which, int reps) {
    DG_REQUIRE(reps > 0, "reps must be positive");
    const long long n = 3LL * nV;
    cudaEvent_t a = ev[0], b = ev[1];
    // one untimed run first
    for (int r = -1; r < reps; ++r) {
        if (r == 0) DG_CUDA(cudaEventRecord(a, st));
        switch (which) {
            case 0: launch_energy(emesh(), x.p, xt.p, cfg.dt * cfg.dt, sc.p + SC_E, st); break;
            case 1: launch_gradient(emesh(), x.p, xt.p, cfg.dt * cfg.dt, g_old.p, st); break;
            case 2: launch_elem_hessians(mesh, x.p, cfg.dt * cfg.dt, true, st); break;
            case 3: launch_fill(fill, mesh.He.p, a_all.p, st); break;
            case 4: if (!owned.empty()) chol.factorize(a_all.p + a_off[1], st); break;
            case 5: precondition_dev(g.p, q.p); break;
            case 6: launch_dot(n, g.p, g.p, dot_partial.p, counter.p, sc.p + SC_DOT, st); break;
            default: throw Error(DOTGPU_ERR_INVALID, "unknown kernel id");
        }
    }
    DG_CUDA(cudaEventRecord(b, st));
    DG_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    DG_CUDA(cudaEventElapsedTime(&ms, a, b));
    return (double)ms / reps;
}

}  // namespace dotgpu
