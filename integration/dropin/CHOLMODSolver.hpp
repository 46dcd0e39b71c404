// Drop-in replacement for the reference's LinSysSolver/CHOLMODSolver.hpp.
//
// The reference spells `new CHOLMODSolver<Eigen::VectorXi, Eigen::VectorXd>()` at every construction site
// (Optimizer.cpp:114-120, ADMMDDTimeStepper.cpp:355-371, DOTTimeStepper.cpp:58-64, 88-94) and finds the class through
// `#include "CHOLMODSolver.hpp"` on the include path.  Putting THIS directory before src/LinSysSolver on the include
// path (and not compiling CHOLMODSolver.cpp) swaps every solver of the unmodified time steppers for the libdotgpu
// one: same class name, same virtuals, same observable behaviour (0-based ia/ja after set_pattern, `factorize`
// returning false on success like CHOLMODSolver.cpp:143-146), no CHOLMOD / BLAS needed any more.
//
// Only the C ABI of include/dotgpu.h is used.  Values live in the base class's host vector `a` exactly as before
// (the steppers fill it through addCoeff/setCoeff from TBB tasks); factorize() ships them to the device.
#ifndef CHOLMODSolver_hpp
#define CHOLMODSolver_hpp

#include "LinSysSolver.hpp"

#include <Eigen/Eigen>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <set>
#include <vector>

#include "dotgpu.h"

namespace DOT {

template <typename vectorTypeI, typename vectorTypeS>
class CHOLMODSolver : public LinSysSolver<vectorTypeI, vectorTypeS> {
    typedef LinSysSolver<vectorTypeI, vectorTypeS> Base;

protected:
    dotgpu_solver* h = nullptr;
    bool values_fresh = false;  // device copy of `a` matches the host copy used by the last factorize
    std::mutex mtx;             // solve_threadSafe is called concurrently on ONE solver by the ADMM consensus step

    static void check(int rc, const char* what) {
        if (rc != DOTGPU_OK) {
            std::fprintf(stderr, "dotgpu: %s failed (%d): %s\n", what, rc, dotgpu_last_error());
            std::exit(1);  // the reference has no error channel here either (asserts / exit)
        }
    }
    void drop() {
        if (h) dotgpu_solver_destroy(h);
        h = nullptr;
        values_fresh = false;
    }
    // CHOLMOD is handed the base class's arrays as a CSC matrix with stype=-1, i.e. it reads ONLY entries with row >= column
    // (CHOLMODSolver.cpp:95-105) - callers of set_pattern(SparseMatrix) may store more (ADMM's consensus matrix).  The C ABI wants
    // exactly that triangle as CSR-upper with ascending columns and the diagonal first, so filter once per pattern.
    std::vector<int32_t> cia, cja;
    std::vector<int> keep;  // position in Base::a of every kept entry
    std::vector<double> vals;
    void ensure() {
        if (h) return;
        const int n = Base::numRows;
        cia.assign(n + 1, 0);
        cja.clear();
        keep.clear();
        std::vector<std::pair<int, int>> row;
        for (int i = 0; i < n; ++i) {
            row.clear();
            for (int k = Base::ia[i]; k < Base::ia[i + 1]; ++k)
                if (Base::ja[k] >= i) row.emplace_back(Base::ja[k], k);
            std::sort(row.begin(), row.end());
            for (const auto& e : row) {
                cja.push_back(e.first);
                keep.push_back(e.second);
            }
            cia[i + 1] = (int32_t)cja.size();
        }
        vals.resize(keep.size());
        check(dotgpu_solver_create(&h, dropin_device(), n, cia.data(), cja.data()), "solver_create");
    }
    void push_values() {
        for (size_t k = 0; k < keep.size(); ++k) vals[k] = Base::a[keep[k]];
        check(dotgpu_solver_set_values(h, vals.data()), "solver_set_values");
    }

public:
    static int& dropin_device() {
        static int d = 0;
        return d;
    }
    CHOLMODSolver(void) {}
    ~CHOLMODSolver(void) { drop(); }

    void set_type(int, int, bool = false) {}

    void set_pattern(const std::vector<std::set<int>>& vNeighbor, const std::set<int>& fixedVert) {
        Base::set_pattern(vNeighbor, fixedVert);
        Base::ia.array() -= 1;  // the base class builds 1-based CSR; CHOLMODSolver.cpp:101 makes it 0-based and so do we
        Base::ja.array() -= 1;
        drop();
    }
    void set_pattern(const Eigen::SparseMatrix<double>& mtr) {  // NOTE: mtr must be SPD, upper/lower triangle as the caller stores it
        Base::set_pattern(mtr);
        drop();
    }
    void update_a(const Eigen::SparseMatrix<double>& mtr) {
        Base::update_a(mtr);
        values_fresh = false;
    }

    void analyze_pattern(void) {  // cholmod_analyze: ordering + supernodal symbolic factorisation
        drop();
        ensure();
    }
    bool factorize(void) {  // cholmod_factorize on the current values of `a`
        ensure();
        push_values();
        const int rc = dotgpu_solver_factorize(h);
        if (rc != DOTGPU_OK && rc != DOTGPU_ERR_NOT_SPD) check(rc, "solver_factorize");
        values_fresh = true;
        return rc != DOTGPU_OK;  // the reference returns !cholmod_factorize(...): false means success
    }
    void solve(Eigen::VectorXd& rhs, Eigen::VectorXd& result) {
        result.conservativeResize(rhs.size());
        check(dotgpu_solver_solve(h, rhs.data(), result.data()), "solver_solve");
    }
    void solve_threadSafe(Eigen::VectorXd& rhs, Eigen::VectorXd& result, int) {
        std::lock_guard<std::mutex> lock(mtx);
        solve(rhs, result);
    }
    virtual void multiply(const Eigen::VectorXd& x, Eigen::VectorXd& Ax) {  // cholmod_sdmult with stype=-1: symmetric SpMV
        ensure();
        // callers modify `a` through get_a()/addCoeff between calls (it is never factorised under DOT)
        push_values();
        Ax.conservativeResize(Base::numRows);
        check(dotgpu_solver_multiply(h, x.data(), Ax.data()), "solver_multiply");
    }
    virtual void outputFactorization(const std::string&) {}
};

}  // namespace DOT

#endif /* CHOLMODSolver_hpp */
