// GpuEnergy<BaseEnergy>: the reference's Energy<3> virtual interface (src/Energy/Energy.hpp:27-226) with the three
// entry points the DOT stepper calls on the hot path answered by libdotgpu:
//   computeEnergyVal(data, redoSVD, svd, F, U, V, Sigma, coef, E)         Energy.hpp:57-64   -> dotgpu_energy_value
//   computeGradient (data, redoSVD, svd, F, U, V, Sigma, coef, g)         Energy.hpp:65-72   -> dotgpu_energy_gradient
//   computeElemHessianByPK(data, redoSVD, svd, F, coef, mask, He, vInds)   Energy.hpp:133-140 -> dotgpu_energy_elem_hessians
// BaseEnergy is the reference's own FixedCoRotEnergy<3> or StableNHEnergy<3>; everything else (the sigma-space
// virtuals used by Optimizer::computeCharNormSq, the unit tests, the per-element GSDD paths) stays the reference's.
// Usage (where main.cpp:892-900 constructs the energy):
//     energyTerms.emplace_back(new DOT::GpuEnergy<DOT::StableNHEnergy<3>>(DOTGPU_ENERGY_SNH));
//
// The device path always evaluates at data.V, so the answers never depend on the caches.  The caches the stepper owns are
// nevertheless kept in the state the reference leaves them in (Energy.cpp:349-382, AutoFlipSVD.hpp:75-81), so that
// ANY reference path that reads them after a GPU call (per-element GSDD paths, computeHessian with redoSVD = false, ...) sees
// what it would see after the CPU call:
//   computeEnergyVal  redoSVD = 0: caches untouched; 1: F, U, V, Sigma AND svd[t].set(U, Sigma, V); 2: F, U, V, Sigma only
//   computeGradient   redoSVD = true: F and svd[t] (the reference recomputes F and the per-element SVD objects, Energy.cpp:457-500)
// filled from dotgpu_energy_svd (the batched 3x3 SVD with the reference's conventions).  coef already contains dt^2.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "Energy.hpp"
#include "dotgpu.h"

namespace DOT {

template <class BaseEnergy>
class GpuEnergy : public BaseEnergy {
    struct Slot {
        dotgpu_energy* h = nullptr;
        std::vector<uint8_t> fixed;
        std::vector<double> x, g, Fb, Ub, Sb, Vb;
    };
    const int energy_type, device;
    mutable std::map<const Mesh<3>*, Slot> slots;  // the stepper passes its global mesh; sub-meshes get their own handle
    mutable std::mutex mtx;

    static void check(int rc, const char* what) {
        if (rc != DOTGPU_OK) {
            std::fprintf(stderr, "dotgpu: %s failed (%d): %s\n", what, rc, dotgpu_last_error());
            std::exit(1);
        }
    }
    Slot& slot(const Mesh<3>& data) const {
        Slot& s = slots[&data];
        const int nV = (int)data.V.rows(), nT = (int)data.F.rows();
        std::vector<uint8_t> fm(nV, 0);
        for (int v : data.fixedVert) fm[v] = 1;
        if (!s.h) {
            std::vector<int32_t> tets((size_t)4 * nT);
            std::vector<double> B((size_t)9 * nT), vol(nT), mu(nT), lam(nT);
            for (int t = 0; t < nT; ++t) {
                for (int k = 0; k < 4; ++k) tets[4 * (size_t)t + k] = data.F(t, k);
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) B[9 * (size_t)t + 3 * i + j] = data.restTriInv[t](i, j);
                vol[t] = data.triArea[t] * data.triWeight[t];  // gradient / Hessian weight (Energy.cpp:965, 763); triWeight == 1 under DOT
                mu[t] = data.u[t];
                lam[t] = data.lambda[t];
            }
            check(dotgpu_energy_create(&s.h, device, energy_type, nV, nT, tets.data(), B.data(), vol.data(), mu.data(), lam.data(), fm.data()),
                  "energy_create");
            s.fixed = fm;
            s.x.resize((size_t)3 * nV);
            s.g.resize((size_t)3 * nV);
        } else if (fm != s.fixed) {
            check(dotgpu_energy_set_fixed(s.h, fm.data()), "energy_set_fixed");
            s.fixed = fm;
        }
        for (int v = 0; v < nV; ++v)  // Eigen::MatrixXd is column-major: interleave to xyz
            for (int c = 0; c < 3; ++c) s.x[3 * (size_t)v + c] = data.V(v, c);
        return s;
    }

    // fills the stepper-owned caches from the device SVD: F always, U/V/Sigma if given, svd[t].set(...) if set_svd
    void fill_caches(Slot& s, const Mesh<3>& data, std::vector<AutoFlipSVD<Eigen::Matrix3d>>* svd, std::vector<Eigen::Matrix3d>& F,
                     std::vector<Eigen::Matrix3d>* U, std::vector<Eigen::Matrix3d>* V, std::vector<Eigen::Vector3d>* Sigma) const {
        const size_t nT = (size_t)data.F.rows();
        s.Fb.resize(9 * nT); s.Ub.resize(9 * nT); s.Vb.resize(9 * nT); s.Sb.resize(3 * nT);
        check(dotgpu_energy_svd(s.h, s.x.data(), s.Fb.data(), s.Ub.data(), s.Sb.data(), s.Vb.data()), "energy_svd");
        if (F.size() < nT) F.resize(nT);
        if (svd && svd->size() < nT) svd->resize(nT);
        for (size_t t = 0; t < nT; ++t) {
            Eigen::Matrix3d Ut, Vt;
            Eigen::Vector3d St(s.Sb[3 * t], s.Sb[3 * t + 1], s.Sb[3 * t + 2]);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    F[t](i, j) = s.Fb[9 * t + 3 * i + j];
                    Ut(i, j) = s.Ub[9 * t + 3 * i + j];
                    Vt(i, j) = s.Vb[9 * t + 3 * i + j];
                }
            if (U && t < U->size()) (*U)[t] = Ut;       // U / V / Sigma are pre-sized to ceil4(nT) by the stepper (Optimizer.cpp:60-65)
            if (V && t < V->size()) (*V)[t] = Vt;
            if (Sigma && t < Sigma->size()) (*Sigma)[t] = St;
            if (svd) (*svd)[t].set(Ut, St, Vt);
        }
    }

public:
    explicit GpuEnergy(int p_energy_type, int p_device = 0) : energy_type(p_energy_type), device(p_device) {}
    ~GpuEnergy() {
        for (auto& kv : slots)
            if (kv.second.h) dotgpu_energy_destroy(kv.second.h);
    }

    // DOTGPU_DROPIN_FILL_CACHES=0 skips the cache upkeep (the DOT stepper itself never reads the caches after a GPU call)
    static bool fill_enabled() {
        static const bool on = !(std::getenv("DOTGPU_DROPIN_FILL_CACHES") && std::getenv("DOTGPU_DROPIN_FILL_CACHES")[0] == '0');
        return on;
    }
    virtual void computeEnergyVal(const Mesh<3>& data, int redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix3d>>& svd, std::vector<Eigen::Matrix3d>& F,
                                  std::vector<Eigen::Matrix3d>& U, std::vector<Eigen::Matrix3d>& V, std::vector<Eigen::Vector3d>& Sigma, double coef,
                                  double& energyVal) const {
        std::lock_guard<std::mutex> lock(mtx);
        Slot& s = slot(data);
        check(dotgpu_energy_value(s.h, s.x.data(), coef, &energyVal), "energy_value");
        if (redoSVD != 0 && fill_enabled()) fill_caches(s, data, redoSVD == 1 ? &svd : nullptr, F, &U, &V, &Sigma);
    }
    virtual void computeGradient(const Mesh<3>& data, bool redoSVD, std::vector<AutoFlipSVD<Eigen::Matrix3d>>& svd, std::vector<Eigen::Matrix3d>& F,
                                 std::vector<Eigen::Matrix3d>&, std::vector<Eigen::Matrix3d>&, std::vector<Eigen::Vector3d>&, double coef,
                                 Eigen::VectorXd& gradient) const {
        std::lock_guard<std::mutex> lock(mtx);
        Slot& s = slot(data);
        gradient.conservativeResize(data.V.rows() * 3);
        check(dotgpu_energy_gradient(s.h, s.x.data(), coef, gradient.data()), "energy_gradient");
        if (redoSVD && fill_enabled()) fill_caches(s, data, &svd, F, nullptr, nullptr, nullptr);
    }
    virtual void computeElemHessianByPK(const Mesh<3>& data, bool /*redoSVD*/, std::vector<AutoFlipSVD<Eigen::Matrix3d>>&,
                                        std::vector<Eigen::Matrix3d>&, double coef, const std::vector<bool>& computeElem,
                                        std::vector<Eigen::Matrix<double, 12, 12>>& elemHessian, std::vector<Eigen::Matrix<int, 1, 4>>& vInds,
                                        bool projectSPD = true) const {
        std::lock_guard<std::mutex> lock(mtx);
        Slot& s = slot(data);
        const size_t nT = (size_t)data.F.rows();
        std::vector<double> He(nT * 144);
        std::vector<int32_t> vi(nT * 4);
        check(dotgpu_energy_elem_hessians(s.h, s.x.data(), coef, projectSPD ? 1 : 0, He.data(), vi.data()), "energy_elem_hessians");
        elemHessian.resize(nT);
        vInds.resize(nT);
        for (size_t t = 0; t < nT; ++t) {
            if (!computeElem[t]) continue;
            for (int i = 0; i < 12; ++i)
                for (int j = 0; j < 12; ++j) elemHessian[t](i, j) = He[t * 144 + 12 * i + j];
            for (int k = 0; k < 4; ++k) vInds[t][k] = vi[4 * t + k];
        }
    }
};

}  // namespace DOT
