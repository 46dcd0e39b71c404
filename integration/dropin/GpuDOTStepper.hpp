// GpuDOTStepper : DOT::Optimizer<3> - the PERFORMANCE boundary of the drop-in: the reference's time-stepper
// virtuals (src/TimeStepper/Optimizer.hpp:87-126) answered by the device-resident stepper of libdotgpu, so that main.cpp's
// `optimizer->precompute(); while (...) optimizer->solve(1);` loop (main.cpp:92-132, 934-960) runs unchanged with every vector
// resident in HBM.  What it replaces: DOTTimeStepper (src/TimeStepper/DOTTimeStepper.cpp:150-182 precompute, 185-270
// updatePrecondMtrAndFactorize, 273-346 fullyImplicit, 384-504 solve_oneStep) + the ADMMDDTimeStepper constructor's domain
// decomposition (ADMMDDTimeStepper.cpp:88-278) + Optimizer::solve's BE update (Optimizer.cpp:327-368).
//
// Construct it where main.cpp:934-937 constructs DOTTimeStepper:
//     optimizer = new DOT::GpuDOTStepper(*temp, energyTerms, energyParams, false, config);
// The base-class constructor still runs (gravity, AnimScripter handles -> result.fixedVert, restart files); everything after it is
// libdotgpu.  State the reference keeps on the host (result.V, velocity, resultV_n, xTilta, dx_Elastic) is refreshed after every
// time step, so saveStatus / the viewer / restart keep working.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Optimizer.hpp"
#include "dotgpu.h"

namespace DOT {

class GpuDOTStepper : public Optimizer<3> {
    typedef Optimizer<3> Base;
    dotgpu_stepper* h = nullptr;
    std::vector<double> xbuf, vbuf, tbuf;
    std::vector<uint8_t> fixedMask;
    std::vector<int32_t> epart;
    int lineSearchHalvings = 0;

    static void check(int rc, const char* what) {
        if (rc != DOTGPU_OK) {
            std::fprintf(stderr, "dotgpu: %s failed (%d): %s\n", what, rc, dotgpu_last_error());
            std::exit(1);
        }
    }
    void maskFromResult() {
        fixedMask.assign(Base::result.V.rows(), 0);
        for (int v : Base::result.fixedVert) fixedMask[v] = 1;
    }
    void packV(const Eigen::MatrixXd& V) {
        const long nV = V.rows();
        xbuf.resize(3 * (size_t)nV);
        for (long v = 0; v < nV; ++v)
            for (int c = 0; c < 3; ++c) xbuf[3 * (size_t)v + c] = V(v, c);
    }
    // host mirrors of the dynamic state (Optimizer.cpp:354-361 keeps them on the host)
    void pullState() {
        const long nV = Base::result.V.rows();
        vbuf.resize(3 * (size_t)nV);
        tbuf.resize(3 * (size_t)nV);
        std::vector<double> xn(3 * (size_t)nV);
        check(dotgpu_stepper_get_state(h, xn.data(), vbuf.data(), tbuf.data()), "stepper_get_state");
        for (long v = 0; v < nV; ++v)
            for (int c = 0; c < 3; ++c) {
                Base::dx_Elastic(v, c) = xn[3 * (size_t)v + c] - Base::xTilta(v, c);   // x - xTilta of THIS step (before computeXTilta)
                Base::result.V(v, c) = xn[3 * (size_t)v + c];
                Base::resultV_n(v, c) = xn[3 * (size_t)v + c];
                Base::velocity[3 * v + c] = vbuf[3 * (size_t)v + c];
            }
        for (long v = 0; v < nV; ++v)
            for (int c = 0; c < 3; ++c) Base::xTilta(v, c) = tbuf[3 * (size_t)v + c];
    }

public:
    GpuDOTStepper(const Mesh<3>& p_data0, const std::vector<Energy<3>*>& p_energyTerms, const std::vector<double>& p_energyParams,
                  bool p_mute = false, const Config& animConfig = Config())
        : Base(p_data0, p_energyTerms, p_energyParams, p_mute, animConfig) {}
    ~GpuDOTStepper() {
        if (h) dotgpu_stepper_destroy(h);
    }
    int getLineSearchHalvings() const { return lineSearchHalvings; }
    const std::vector<int32_t>& getElementLabels() const { return epart; }

    // DOTTimeStepper::precompute (+ the ADMMDDTimeStepper constructor): METIS labels, DD set-up, rest-state Hessians, factorisation
    virtual void precompute(void) {
        const Mesh<3>& M = Base::result;
        const int nV = (int)M.V_rest.rows(), nT = (int)M.F.rows();
        std::vector<int32_t> tets(4 * (size_t)nT);
        for (int t = 0; t < nT; ++t)
            for (int k = 0; k < 4; ++k) tets[4 * (size_t)t + k] = M.F(t, k);
        dotgpu_stepper_config c;
        dotgpu_stepper_default_config(&c);
        const bool newton = Base::animConfig.timeStepperType == TST_NEWTON;
        const bool lbfgsh = Base::animConfig.timeStepperType == TST_LBFGSH, lbfgsjh = Base::animConfig.timeStepperType == TST_LBFGSJH;
        c.energy_type = Base::animConfig.energyType == ET_SNH ? DOTGPU_ENERGY_SNH : DOTGPU_ENERGY_FCR;
        c.num_subdomains = (newton || lbfgsh) ? 1 : Base::animConfig.partitionAmt;
        c.dt = Base::dt;
        for (int i = 0; i < 3; ++i) c.gravity[i] = Base::gravity[i];
        c.YM = Base::animConfig.YM;
        c.PR = Base::animConfig.PR;
        c.rho = Base::animConfig.rho;
        if (newton) c.flags |= DOTGPU_FLAG_NEWTON;
        if (lbfgsh) c.flags |= DOTGPU_FLAG_LBFGS_H;      // LBFGSTimeStepper D0T_H / D0T_JH on the same kernels
        epart.assign(nT, 0);
        std::vector<int32_t> npart;
        if (lbfgsjh) {
            c.flags |= DOTGPU_FLAG_LBFGS_JH;
            npart.resize(nV);
            check(dotgpu_partition_nodes(nV, nT, tets.data(), c.num_subdomains, npart.data()), "partition_nodes");
            c.node_part = npart.data();
        } else if (!newton && !lbfgsh) {
            // METIS<3>::partMesh with the reference's vendored METIS and option vector (Utils/METIS.hpp:109-160, 265-321)
            check(dotgpu_partition(nV, nT, tets.data(), c.num_subdomains, epart.data()), "partition");
        }
        maskFromResult();
        packV(M.V_rest);
        check(dotgpu_stepper_create(&h, &c, nV, nT, xbuf.data(), tets.data(), epart.data(), fixedMask.data()), "stepper_create");
        // a restart file (Optimizer ctor :126-177) or a pre-deformed start: positions + velocity of the base class
        bool moved = (M.V - M.V_rest).cwiseAbs().maxCoeff() > 0.0 || Base::velocity.cwiseAbs().maxCoeff() > 0.0;
        if (moved) {
            packV(M.V);
            std::vector<double> vel(Base::velocity.data(), Base::velocity.data() + Base::velocity.size());
            check(dotgpu_stepper_set_state(h, xbuf.data(), vel.data()), "stepper_set_state");
        }
        double tg = 0.0;
        check(dotgpu_stepper_get_target(h, &tg), "stepper_get_target");
        Base::targetGRes = tg;
    }

    // Optimizer::setRelGL2Tol (Optimizer.cpp:222-228)
    virtual void setRelGL2Tol(double p_relTol = 1.0e-5) {
        Base::relGL2Tol = p_relTol * p_relTol;
        if (h) {
            check(dotgpu_stepper_set_rel_tol(h, p_relTol), "stepper_set_rel_tol");
            double tg = 0.0;
            check(dotgpu_stepper_get_target(h, &tg), "stepper_get_target");
            Base::targetGRes = tg;
        }
    }

    // DOTTimeStepper::updatePrecondMtrAndFactorize (DOTTimeStepper.cpp:185-270): the Dirichlet set changed
    virtual void updatePrecondMtrAndFactorize(void) {
        maskFromResult();
        packV(Base::result.V);
        check(dotgpu_stepper_set_fixed(h, fixedMask.data(), xbuf.data()), "stepper_set_fixed");
    }

    // Optimizer::solve (Optimizer.cpp:327-368): scripted Dirichlet motion, fullyImplicit, BE update - per time step
    virtual int solve(int maxIter = 100) {
        int returnFlag = 0;
        for (int iterI = 0; iterI < maxIter; iterI++) {
            if (Base::animScripter.stepAnimScript(Base::result, Base::dt, Base::energyTerms)) updatePrecondMtrAndFactorize();
            if (Base::globalIterNum >= Base::frameAmt) {
                Base::lastEDec = 0.0;
                Base::globalIterNum++;
                return 1;
            }
            packV(Base::result.V);
            dotgpu_frame_stats st;
            check(dotgpu_stepper_frame(h, xbuf.data(), &st), "stepper_frame");
            if (!st.converged) returnFlag = 2;
            Base::innerIterAmt += st.iters;
            Base::numOfLineSearch += st.halvings;
            lineSearchHalvings += st.halvings;
            Base::lastEnergyVal = st.E;
            // iterStats.txt rows of DOTTimeStepper::fullyImplicit (DOTTimeStepper.cpp:298, 329)
            std::vector<double> log(3 * (size_t)(st.iters + 2));
            const int rows = dotgpu_stepper_get_iter_log(h, log.data(), st.iters + 2);
            for (int r = 0; r < rows; ++r)
                Base::file_iterStats << Base::globalIterNum << " " << log[3 * r] << " " << log[3 * r + 1] << " " << log[3 * r + 2] << std::endl;
            pullState();
            Base::globalIterNum++;
        }
        return returnFlag;
    }
};

}  // namespace DOT
