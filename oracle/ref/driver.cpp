// TEST INFRASTRUCTURE — oracle/_ref driver.  Not part of the product path.
//
// Headless replacement for the reference's viewer-coupled main.cpp: it compiles against
// the UNMODIFIED reference sources under /root/reference (never copied into this repo)
// and follows main.cpp:652-960 minus the viewer, i.e. load script -> load+normalise mesh
// (main.cpp:709-710) -> Mesh -> initSIMD scratch (main.cpp:521-597) -> Energy ->
// DOTTimeStepper/Optimizer -> precompute -> per frame { setRelGL2Tol; solve(1) }
// (main.cpp:92-132).  On top of that it can dump kernel-level quantities as .npy files
// so that tests/ can pin the numpy restatement (oracle/dot_oracle.py) and the CUDA path
// against the reference's own numbers.
//
// `#define protected public` below is confined to this translation unit and only serves
// read access to the stepper's state (SURVEY.md App. A.10 uses the same trick).
#include <sstream>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <deque>
#include <vector>
#include <string>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <sys/stat.h>
#include <unistd.h>
#include <omp.h>
#include <mutex>

#define protected public
#define private public
#include "Types.hpp"
#include "IglUtils.hpp"
#include "Config.hpp"
#include "Optimizer.hpp"
#include "ADMMDDTimeStepper.hpp"
#include "DOTTimeStepper.hpp"
#include "LBFGSTimeStepper.hpp"
#include "FixedCoRotEnergy.hpp"
#include "StableNHEnergy.hpp"
#include "METIS.hpp"
#include "CHOLMODSolver.hpp"
#include "Timer.hpp"
#undef protected
#undef private
#ifdef DOTGPU_DROPIN
// drop-in build (oracle/_ref/dot_ref_gpu): the same unmodified reference steppers, but "CHOLMODSolver.hpp" above resolved to
// integration/dropin/CHOLMODSolver.hpp (libdotgpu-backed LinSysSolver) and the energy object is GpuEnergy<reference energy>.
#include "GpuEnergy.hpp"
#include "GpuDOTStepper.hpp"
#endif

// ---- globals the reference translation units expect (main.cpp:27-88) ----
DOT::Config config;  // global => zero-initialised enums
std::ofstream logFile;
std::string outputFolderPath = "output/";
Eigen::MatrixXi SF;
std::vector<int> sTri2Tet;
std::vector<bool> isSurfNode;
std::vector<int> tetIndToSurf;
std::vector<int> surfIndToTet;
Eigen::MatrixXd V_surf;
Eigen::MatrixXi F_surf;
Timer timer, timer_step, timer_temp, timer_temp2, timer_temp3;
double *a11, *a21, *a31, *a12, *a22, *a32, *a13, *a23, *a33;
double *u11, *u21, *u31, *u12, *u22, *u32, *u13, *u23, *u33;
double *v11, *v21, *v31, *v12, *v22, *v32, *v13, *v23, *v33;
double *sigma1, *sigma2, *sigma3;
double *Gmu, *Glambda, *Gsigma0, *Gsigma1, *Gsigma2;

static double* alloc64(size_t n)
{
    void* p = nullptr;
    if (posix_memalign(&p, 64, n * sizeof(double)) != 0) { std::perror("posix_memalign"); std::exit(1); }
    std::memset(p, 0, n * sizeof(double));
    return reinterpret_cast<double*>(p);
}

// ---- tiny .npy writer / reader (v1.0, C order) ----
static void npy_write(const std::string& path, const char* descr, const std::vector<long>& shape,
                      const void* data, size_t bytes)
{
    std::string sh = "(";
    for (size_t i = 0; i < shape.size(); ++i) { sh += std::to_string(shape[i]); sh += (shape.size() == 1 || i + 1 < shape.size()) ? "," : ""; }
    sh += ")";
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
    size_t total = 10 + hdr.size() + 1;
    size_t pad = (64 - total % 64) % 64;
    hdr += std::string(pad, ' ');
    hdr += '\n';
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { std::perror(path.c_str()); std::exit(1); }
    unsigned char magic[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(hdr.size() & 0xff), (unsigned char)(hdr.size() >> 8)};
    std::fwrite(magic, 1, 10, f);
    std::fwrite(hdr.data(), 1, hdr.size(), f);
    if (bytes) std::fwrite(data, 1, bytes, f);
    std::fclose(f);
}
static void npy_f64(const std::string& p, const std::vector<long>& shape, const double* d)
{
    size_t n = 1; for (long s : shape) n *= s;
    npy_write(p, "<f8", shape, d, n * 8);
}
static void npy_i32(const std::string& p, const std::vector<long>& shape, const int* d)
{
    size_t n = 1; for (long s : shape) n *= s;
    npy_write(p, "<i4", shape, d, n * 4);
}
static void npy_i64(const std::string& p, const std::vector<long>& shape, const long long* d)
{
    size_t n = 1; for (long s : shape) n *= s;
    npy_write(p, "<i8", shape, d, n * 8);
}
static std::vector<double> npy_read_f64(const std::string& path, std::vector<long>& shape)
{
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { std::perror(path.c_str()); std::exit(1); }
    unsigned char magic[10];
    if (std::fread(magic, 1, 10, f) != 10) std::exit(1);
    size_t hl = magic[8] | (magic[9] << 8);
    std::string hdr(hl, ' ');
    if (std::fread(&hdr[0], 1, hl, f) != hl) std::exit(1);
    if (hdr.find("<f8") == std::string::npos || hdr.find("False") == std::string::npos) {
        std::cerr << "npy_read_f64: need C-order <f8: " << path << std::endl; std::exit(1);
    }
    size_t a = hdr.find("'shape': (") + 10, b = hdr.find(')', a);
    std::stringstream ss(hdr.substr(a, b - a));
    shape.clear();
    std::string tok;
    while (std::getline(ss, tok, ',')) { if (tok.find_first_of("0123456789") != std::string::npos) shape.push_back(std::stol(tok)); }
    size_t n = 1; for (long s : shape) n *= s;
    std::vector<double> out(n);
    if (std::fread(out.data(), 8, n, f) != n) { std::cerr << "short read " << path << std::endl; std::exit(1); }
    std::fclose(f);
    return out;
}

typedef DOT::Optimizer<DIM> Opt;
typedef DOT::DOTTimeStepper<DIM> DotOpt;
typedef DOT::CHOLMODSolver<Eigen::VectorXi, Eigen::VectorXd> CholSolver;

static std::vector<double> rowmajor(const Eigen::MatrixXd& M)
{
    std::vector<double> o((size_t)M.rows() * M.cols());
    for (long i = 0; i < M.rows(); ++i) for (long j = 0; j < M.cols(); ++j) o[i * M.cols() + j] = M(i, j);
    return o;
}

static void dump_solver(const std::string& prefix, DOT::LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* s, bool pattern)
{
    if (pattern) {
        npy_i32(prefix + "ia.npy", {(long)s->ia.size()}, s->ia.data());
        npy_i32(prefix + "ja.npy", {(long)s->ja.size()}, s->ja.data());
    }
    npy_f64(prefix + "a.npy", {(long)s->a.size()}, s->a.data());
}

static void dump_setup(const std::string& dir, Opt* opt, DotOpt* dot)
{
    const DOT::Mesh<DIM>& m = opt->result;
    long nV = m.V.rows(), nT = m.F.rows();
    npy_f64(dir + "V_rest.npy", {nV, 3}, rowmajor(m.V_rest).data());
    {
        std::vector<int> F((size_t)nT * 4);
        for (long t = 0; t < nT; ++t) for (int k = 0; k < 4; ++k) F[t * 4 + k] = m.F(t, k);
        npy_i32(dir + "F.npy", {nT, 4}, F.data());
    }
    {
        std::vector<double> B((size_t)nT * 9);
        for (long t = 0; t < nT; ++t) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[t * 9 + i * 3 + j] = m.restTriInv[t](i, j);
        npy_f64(dir + "restTriInv.npy", {nT, 3, 3}, B.data());
    }
    npy_f64(dir + "triArea.npy", {nT}, m.triArea.data());
    npy_f64(dir + "mu.npy", {nT}, m.u.data());
    npy_f64(dir + "lambda.npy", {nT}, m.lambda.data());
    {
        std::vector<double> mass(nV);
        for (long v = 0; v < nV; ++v) mass[v] = m.massMatrix.coeff(v, v);
        npy_f64(dir + "mass.npy", {nV}, mass.data());
    }
    {
        std::vector<int> fx(m.fixedVert.begin(), m.fixedVert.end());
        npy_i32(dir + "fixed.npy", {(long)fx.size()}, fx.data());
    }
    {
        double s[6] = {opt->targetGRes, opt->dt, opt->gravity[0], opt->gravity[1], opt->gravity[2], opt->relGL2Tol};
        npy_f64(dir + "scalars.npy", {6}, s);
    }
    dump_solver(dir + "global_", opt->linSysSolver, true);
    if (dot) {
        long k = (long)dot->mesh_subdomain.size();
        std::vector<long long> epart(nT, -1);
        for (long s = 0; s < k; ++s) for (long i = 0; i < dot->elemList_subdomain[s].size(); ++i) epart[dot->elemList_subdomain[s][i]] = s;
        npy_i64(dir + "epart.npy", {nT}, epart.data());
        npy_i32(dir + "dup.npy", {nV}, dot->dup.data());
        for (long s = 0; s < k; ++s) {
            std::string p = dir + "sbd" + std::to_string(s) + "_";
            npy_i32(p + "l2g.npy", {(long)dot->localVIToGlobal_subdomain[s].size()}, dot->localVIToGlobal_subdomain[s].data());
            std::vector<int> fx(dot->mesh_subdomain[s].fixedVert.begin(), dot->mesh_subdomain[s].fixedVert.end());
            npy_i32(p + "fixed.npy", {(long)fx.size()}, fx.data());
            dump_solver(p, dot->linSysSolver_subdomain[s], true);
        }
    }
}

// kernel-level dump of the current state (call only between frames: it re-runs the
// redoSVD=1 energy pass, which leaves svd[]/F consistent with result.V).
static void dump_state(const std::string& dir, Opt* opt, DotOpt* dot, long heCap)
{
    DOT::Mesh<DIM>& m = opt->result;
    long nV = m.V.rows(), nT = m.F.rows();
    npy_f64(dir + "V.npy", {nV, 3}, rowmajor(m.V).data());
    npy_f64(dir + "V_n.npy", {nV, 3}, rowmajor(opt->resultV_n).data());
    npy_f64(dir + "xTilta.npy", {nV, 3}, rowmajor(opt->xTilta).data());
    npy_f64(dir + "velocity.npy", {nV * 3}, opt->velocity.data());
    {
        std::vector<int> fx(m.fixedVert.begin(), m.fixedVert.end());
        npy_i32(dir + "fixed.npy", {(long)fx.size()}, fx.data());
    }
    double E = 0, Eel = 0;
    opt->computeEnergyVal(m, 1, E);                       // Optimizer.cpp:1183 (elastic + inertia)
    Eel = opt->energyVal_ET[0];
    Eigen::VectorXd g;
    opt->computeGradient(m, false, g);                    // Optimizer.cpp:1220
    Eigen::VectorXd gel = opt->gradient_ET[0];
    double sc[2] = {E, Eel};
    npy_f64(dir + "E.npy", {2}, sc);
    npy_f64(dir + "g.npy", {nV * 3}, g.data());
    npy_f64(dir + "g_elastic.npy", {nV * 3}, gel.data());
    {
        std::vector<double> Fm((size_t)nT * 9), U((size_t)nT * 9), Vm((size_t)nT * 9), S((size_t)nT * 3);
        for (long t = 0; t < nT; ++t) {
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
                Fm[t * 9 + i * 3 + j] = opt->F[t](i, j);
                U[t * 9 + i * 3 + j] = opt->svd[t].matrixU()(i, j);
                Vm[t * 9 + i * 3 + j] = opt->svd[t].matrixV()(i, j);
            }
            for (int i = 0; i < 3; ++i) S[t * 3 + i] = opt->svd[t].singularValues()[i];
        }
        npy_f64(dir + "F.npy", {nT, 3, 3}, Fm.data());
        npy_f64(dir + "U.npy", {nT, 3, 3}, U.data());
        npy_f64(dir + "Vsvd.npy", {nT, 3, 3}, Vm.data());
        npy_f64(dir + "Sigma.npy", {nT, 3}, S.data());
    }
    // per-element energies (unweighted by coef) through the SIMD path
    {
        Eigen::VectorXd epe;
        opt->energyTerms[0]->getEnergyValPerElemBySVD(m, 0, opt->svd, opt->F, opt->U, opt->V, opt->Sigma, epe);
        npy_f64(dir + "E_per_elem.npy", {nT}, epe.data());
    }
    // elemental PD-projected Hessians (Energy.cpp:673-701), coef = dt^2
    {
        std::vector<bool> all(nT, true);
        std::vector<Eigen::Matrix<double, 12, 12>> He;
        std::vector<Eigen::Matrix<int, 1, 4>> vInds;
        opt->energyTerms[0]->computeElemHessianByPK(m, false, opt->svd, opt->F, opt->dtSq * opt->energyParams[0], all, He, vInds, true);
        long n = (heCap < 0 || heCap > nT) ? nT : heCap;
        std::vector<double> H((size_t)n * 144);
        for (long t = 0; t < n; ++t) for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) H[t * 144 + i * 12 + j] = He[t](i, j);
        npy_f64(dir + "He.npy", {n, 12, 12}, H.data());
        // checksum over all elements so that the full set is pinned even when capped
        std::vector<double> frob(nT);
        for (long t = 0; t < nT; ++t) frob[t] = He[t].squaredNorm();
        npy_f64(dir + "He_sqnorm.npy", {nT}, frob.data());
    }
    if (dot) {
        // Hessian refresh at this state (DOTTimeStepper.cpp:349-380) -> CSR values + factor
        dot->updateHessianAndFactor();
        dump_solver(dir + "global_", opt->linSysSolver, false);
        long k = (long)dot->mesh_subdomain.size();
        for (long s = 0; s < k; ++s) dump_solver(dir + "sbd" + std::to_string(s) + "_", dot->linSysSolver_subdomain[s], false);
        // one preconditioner application p = D^-1 sum_s R_s^T H_s^-1 R_s (-g)   (DOTTimeStepper.cpp:406-450)
        Eigen::VectorXd q = -g, p = Eigen::VectorXd::Zero(nV * 3);
        for (long s = 0; s < k; ++s) {
            long nl = dot->mesh_subdomain[s].V.rows();
            Eigen::VectorXd rhs(nl * 3), ps;
            for (long l = 0; l < nl; ++l) rhs.segment<3>(l * 3) = q.segment<3>(dot->localVIToGlobal_subdomain[s][l] * 3);
            dot->linSysSolver_subdomain[s]->solve(rhs, ps);
            npy_f64(dir + "sbd" + std::to_string(s) + "_p.npy", {nl * 3}, ps.data());
            for (long l = 0; l < nl; ++l) p.segment<3>(dot->localVIToGlobal_subdomain[s][l] * 3) += ps.segment<3>(l * 3);
        }
        for (long v = 0; v < nV; ++v) if (dot->dup[v] > 1) p.segment<3>(v * 3) /= dot->dup[v];
        npy_f64(dir + "p.npy", {nV * 3}, p.data());
        Eigen::VectorXd Hp;
        opt->linSysSolver->multiply(p, Hp);
        npy_f64(dir + "Hp.npy", {nV * 3}, Hp.data());
    }
}

static void usage()
{
    std::cerr << "dot_ref --script <file.txt> [--mesh <file.msh>] [--energy SNH|FCR] [--parts K] [--stepper DOT|Newton]\n"
                 "        [--tol T] [--dt DT] [--anim <script name>] [--frames N] [--threads N] [--quiet] [--labels-only]\n"
                 "        [--dump-dir D] [--dump-frames a,b,c] [--he-cap N] [--kernel-state V.npy] [--stats-json file]\n"
                 "        [--final-V V.npy] [--full-precision] [--resident (drop-in build: GpuDOTStepper)]\n";
}

int main(int argc, char** argv)
{
    std::string script, meshOverride, energy, stepper, anim, dumpDir, kernelState, statsJson, finalV, saveMsh;
    bool saveStatus = false;
    bool fullPrecision = false;
    int parts = -1, frames = 10, threads = 0;
    long heCap = -1;
    double tol = -1, dtOverride = -1;
    bool quiet = false, labelsOnly = false, cpuEnergy = false, resident = false;
    std::set<int> dumpFrames;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { usage(); std::exit(2); } return argv[++i]; };
        if (a == "--script") script = next();
        else if (a == "--mesh") meshOverride = next();
        else if (a == "--energy") energy = next();
        else if (a == "--parts") parts = std::stoi(next());
        else if (a == "--stepper") stepper = next();
        else if (a == "--tol") tol = std::stod(next());
        else if (a == "--dt") dtOverride = std::stod(next());
        else if (a == "--anim") anim = next();
        else if (a == "--frames") frames = std::stoi(next());
        else if (a == "--threads") threads = std::stoi(next());
        else if (a == "--quiet") quiet = true;
        else if (a == "--labels-only") labelsOnly = true;
        else if (a == "--cpu-energy") cpuEnergy = true;  // drop-in build only: keep the reference's CPU energy, GPU solvers only
        else if (a == "--resident") resident = true;      // drop-in build only: GpuDOTStepper (device-resident stepper) instead of DOTTimeStepper
        else if (a == "--dump-dir") dumpDir = next();
        else if (a == "--he-cap") heCap = std::stol(next());
        else if (a == "--kernel-state") kernelState = next();
        else if (a == "--stats-json") statsJson = next();
        else if (a == "--save-status") saveStatus = true;     // Optimizer::saveStatus() after the last frame -> <output folder>/status<n> (+ <n>.obj)
        else if (a == "--save-msh") saveMsh = next();         // Mesh::saveAsMesh -> IglUtils::saveTetMesh of the final configuration
        else if (a == "--final-V") finalV = next();            // positions after the last frame, [nV,3] float64 .npy
        else if (a == "--full-precision") fullPrecision = true;  // iterStats.txt with 17 significant digits (default: the reference's 6)
        else if (a == "--dump-frames") { std::stringstream ss(next()); std::string t; while (std::getline(ss, t, ',')) dumpFrames.insert(std::stoi(t)); }
        else { usage(); return 2; }
    }
    if (script.empty()) { usage(); return 2; }
    if (threads > 0) omp_set_num_threads(threads);

    if (config.loadFromFile(script) != 0) { std::cerr << "cannot load script " << script << std::endl; return 1; }
    if (!energy.empty()) config.energyType = DOT::Config::getEnergyTypeByStr(energy);
    if (!stepper.empty()) config.timeStepperType = DOT::Config::getTimeStepperTypeByStr(stepper);
    if (parts > 0) config.partitionAmt = parts;
    if (dtOverride > 0) config.dt = dtOverride;
    if (!anim.empty()) config.animScriptType = DOT::AnimScripter<DIM>::getAnimScriptTypeByStr(anim);
    if (!meshOverride.empty()) config.inputShapePath = meshOverride;
    if (config.shapeType != DOT::P_INPUT) { std::cerr << "only `shape input <msh>` scripts are supported" << std::endl; return 1; }

    std::streambuf* coutBuf = std::cout.rdbuf();
    std::ofstream devnull("/dev/null");
    FILE* savedStdout = nullptr;
    int savedFd = -1;
    if (quiet) {
        std::cout.rdbuf(devnull.rdbuf());
        // METIS prints through printf (dbglvl 511, METIS.hpp:288): silence the C stream too
        fflush(stdout);
        savedFd = dup(fileno(stdout));
        if (!freopen("/dev/null", "w", stdout)) return 1;
    }
    (void)savedStdout;

    // ---- main.cpp:673-712 ----
    Eigen::MatrixXd V, UV;
    Eigen::MatrixXi F;
    DOT::IglUtils::readTetMesh(config.inputShapePath, V, F, SF);
    if (config.rotDeg != 0.0) {
        const Eigen::Matrix3d rotMtr = Eigen::AngleAxis<double>(config.rotDeg / 180.0 * M_PI, config.rotAxis).toRotationMatrix();
        for (int vI = 0; vI < V.rows(); ++vI) V.row(vI) = (rotMtr * V.row(vI).transpose()).transpose();
    }
    V *= config.size / (V.colwise().maxCoeff() - V.colwise().minCoeff()).maxCoeff();
    V.rowwise() -= V.colwise().minCoeff();
    UV = V.leftCols(DIM);
    std::vector<std::vector<int>> borderVerts_primitive;
    DOT::IglUtils::findBorderVerts(V, borderVerts_primitive, config.handleRatio);
    DOT::IglUtils::buildSTri2Tet(F, SF, sTri2Tet);

    // ---- main.cpp:782-830 ----
    DOT::Mesh<DIM>* temp = new DOT::Mesh<DIM>(V, F, UV, config.YM, config.PR, config.rho);
    temp->computeBoundaryVert(SF);
    temp->borderVerts_primitive = borderVerts_primitive;
    if (config.blockSize > 0) config.partitionAmt = temp->V_rest.rows() / config.blockSize + 1;
    {
        isSurfNode.assign(temp->V.rows(), false);
        for (int tI = 0; tI < SF.rows(); ++tI) { isSurfNode[SF(tI, 0)] = isSurfNode[SF(tI, 1)] = isSurfNode[SF(tI, 2)] = true; }
        tetIndToSurf.assign(temp->V.rows(), -1);
        surfIndToTet.assign(temp->V.rows(), -1);
        int sVI = 0;
        for (int vI = 0; vI < (int)isSurfNode.size(); ++vI) if (isSurfNode[vI]) { tetIndToSurf[vI] = sVI; surfIndToTet[sVI] = vI; ++sVI; }
        V_surf.resize(sVI, 3);
        F_surf.resize(SF.rows(), 3);
        for (int tI = 0; tI < SF.rows(); ++tI) for (int c = 0; c < 3; ++c) F_surf(tI, c) = tetIndToSurf[SF(tI, c)];
    }
    if (labelsOnly) {
        // METIS labels exactly as ADMMDDTimeStepper's ctor obtains them (ADMMDDTimeStepper.cpp:88-92):
        // k-way partition of the dual graph with the option vector of METIS.hpp:265-321.
        DOT::METIS<DIM> partitions(*temp);
        partitions.partMesh(config.partitionAmt);
        std::vector<long long> ep(partitions.epart.begin(), partitions.epart.end());
        mkdir(dumpDir.c_str(), 0777);
        mkdir((dumpDir + "/setup").c_str(), 0777);
        npy_i64(dumpDir + "/setup/epart.npy", {(long)ep.size()}, ep.data());
        return 0;
    }
    {   // initSIMD (main.cpp:521-597)
        size_t size = std::ceil(temp->F.rows() / 4.f) * 4;
        double** all[] = {&a11, &a21, &a31, &a12, &a22, &a32, &a13, &a23, &a33, &u11, &u21, &u31, &u12, &u22, &u32, &u13, &u23, &u33,
                          &v11, &v21, &v31, &v12, &v22, &v32, &v13, &v23, &v33, &sigma1, &sigma2, &sigma3, &Gmu, &Glambda, &Gsigma0, &Gsigma1, &Gsigma2};
        for (auto p : all) *p = alloc64(size);
    }
    mkdir("output", 0777);
    outputFolderPath = "output/ref_" + std::to_string((long)getpid()) + "/";
    mkdir(outputFolderPath.c_str(), 0777);
    logFile.open(outputFolderPath + "log.txt");

    // ---- main.cpp:865-888 ----
    timer.new_activity("descent");
    const char* stepActs[] = {"matrixComputation", "matrixAssembly", "symbolicFactorization", "numericalFactorization", "backSolve", "lineSearch_other",
                              "modifyGrad", "modifySearchDir", "updateHistory", "lineSearch_eVal", "fullyImplicit_eComp", "solve_extraComp", "compGrad", "CCD"};
    for (auto a : stepActs) timer_step.new_activity(a);
    const char* t3Acts[] = {"init", "initPrimal", "initDual", "initWeights", "initCons", "subdSolve", "consSolve"};
    for (auto a : t3Acts) timer_temp3.new_activity(a);

    // ---- main.cpp:891-942 ----
    std::vector<DOT::Energy<DIM>*> energyTerms;
    std::vector<double> energyParams;
    energyParams.emplace_back(1.0);
#ifdef DOTGPU_DROPIN
    if (!cpuEnergy) {
        switch (config.energyType) {
            case DOT::ET_SNH: energyTerms.emplace_back(new DOT::GpuEnergy<DOT::StableNHEnergy<DIM>>(DOTGPU_ENERGY_SNH)); break;
            case DOT::ET_FCR: energyTerms.emplace_back(new DOT::GpuEnergy<DOT::FixedCoRotEnergy<DIM>>(DOTGPU_ENERGY_FCR)); break;
        }
    } else
#endif
    switch (config.energyType) {
        case DOT::ET_SNH: energyTerms.emplace_back(new DOT::StableNHEnergy<DIM>()); break;
        case DOT::ET_FCR: energyTerms.emplace_back(new DOT::FixedCoRotEnergy<DIM>()); break;
    }
    (void)cpuEnergy;
    auto t0 = std::chrono::steady_clock::now();
    Opt* opt = nullptr;
    DotOpt* dot = nullptr;
#ifdef DOTGPU_DROPIN
    if (resident) {
        if (!dumpDir.empty() || !kernelState.empty()) { std::cerr << "--resident cannot be combined with the dump modes" << std::endl; return 1; }
        opt = new DOT::GpuDOTStepper(*temp, energyTerms, energyParams, false, config);
    } else
#endif
    switch (config.timeStepperType) {
        case DOT::TST_NEWTON: opt = new Opt(*temp, energyTerms, energyParams, false, config); break;
        case DOT::TST_DOT: dot = new DotOpt(*temp, energyTerms, energyParams, false, config); opt = dot; break;
        // the other L-BFGS initialisers that share the kernels (main.cpp:922-931)
        case DOT::TST_LBFGSH: opt = new DOT::LBFGSTimeStepper<DIM>(*temp, energyTerms, energyParams, DOT::D0T_H, false, config); break;
        case DOT::TST_LBFGSJH: {
            auto* jh = new DOT::LBFGSTimeStepper<DIM>(*temp, energyTerms, energyParams, DOT::D0T_JH, false, config);
            opt = jh;
            if (!dumpDir.empty()) {  // node labels of METIS<3>::partMesh_nodes (LBFGSTimeStepper.cpp:71-74)
                std::vector<long long> np(jh->nodePart.begin(), jh->nodePart.end());
                mkdir(dumpDir.c_str(), 0777);
                mkdir((dumpDir + "/setup").c_str(), 0777);
                npy_i64(dumpDir + "/setup/npart.npy", {(long)np.size()}, np.data());
            }
            break;
        }
        default: std::cerr << "driver supports timeStepper DOT, Newton, LBFGSH and LBFGSJH only" << std::endl; return 1;
    }
    opt->setTime(config.duration, config.dt);
    opt->precompute();
    opt->setAllowEDecRelTol(false);
    double setupSec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    auto mk = [&](const std::string& d) { mkdir(d.c_str(), 0777); return d + "/"; };
    if (!dumpDir.empty()) { mk(dumpDir); dump_setup(mk(dumpDir + "/setup"), opt, dot); }

    if (!kernelState.empty()) {
        // kernel mode: overwrite positions with a caller-supplied state and dump everything at it
        std::vector<long> sh;
        std::vector<double> Vin = npy_read_f64(kernelState, sh);
        if (sh.size() != 2 || sh[0] != opt->result.V.rows() || sh[1] != 3) { std::cerr << "bad kernel-state shape" << std::endl; return 1; }
        for (long v = 0; v < sh[0]; ++v) for (int c = 0; c < 3; ++c) opt->result.V(v, c) = Vin[v * 3 + c];
        dump_state(mk(dumpDir + "/kernel"), opt, dot, heCap);
        if (quiet) { std::cout.rdbuf(coutBuf); }
        return 0;
    }

    if (fullPrecision) opt->file_iterStats.precision(17);
    // ---- frame loop: main.cpp:92-132 ----
    std::vector<double> frameSec;
    std::vector<int> frameIters;
    auto tl0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) {
        auto tf0 = std::chrono::steady_clock::now();
        if (tol > 0) opt->setRelGL2Tol(tol); else opt->setRelGL2Tol();
        int before = opt->getInnerIterAmt();
        opt->solve(1);
        frameSec.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - tf0).count());
        frameIters.push_back(opt->getInnerIterAmt() - before);
        if (!dumpDir.empty() && dumpFrames.count(f + 1)) dump_state(mk(dumpDir + "/frame" + std::to_string(f + 1)), opt, dot, heCap);
    }
    double loopSec = 0;
    for (double s : frameSec) loopSec += s;
    (void)tl0;
    opt->file_iterStats.flush();

    if (quiet) {
        std::cout.rdbuf(coutBuf);
        fflush(stdout);
        dup2(savedFd, fileno(stdout));
        close(savedFd);
    }
    const DOT::Mesh<DIM>& R = opt->getResult();
    double sumV = R.V.sum(), sqV = R.V.squaredNorm();
    if (saveStatus) opt->saveStatus();
    if (!saveMsh.empty()) DOT::IglUtils::saveTetMesh(saveMsh, R.V, R.F, SF, false);
    if (!finalV.empty()) {
        std::vector<double> rm((size_t)R.V.rows() * 3);
        for (long v = 0; v < R.V.rows(); ++v) for (int c = 0; c < 3; ++c) rm[v * 3 + c] = R.V(v, c);
        npy_f64(finalV, {(long)R.V.rows(), 3}, rm.data());
    }
    std::ostringstream js;
    js.precision(17);
    js << "{\"frames\": " << frames << ", \"inner_iters\": " << opt->getInnerIterAmt() << ", \"loop_sec\": " << loopSec
       << ", \"setup_sec\": " << setupSec << ", \"fps\": " << (frames / loopSec) << ", \"threads\": " << omp_get_max_threads()
       << ", \"nT\": " << R.F.rows() << ", \"nV\": " << R.V.rows() << ", \"parts\": " << config.partitionAmt
       << ", \"energy\": \"" << DOT::Config::getStrByEnergyType(config.energyType) << "\""
       << ", \"sumV\": " << sumV << ", \"sqnormV\": " << sqV << ", \"line_search_halvings\": " << opt->numOfLineSearch
       << ", \"targetGRes\": " << opt->targetGRes << ", \"iter_stats\": \"" << outputFolderPath << "iterStats.txt\""
       << ", \"output_folder\": \"" << outputFolderPath << "\", \"timestep\": " << opt->globalIterNum;
    js << ", \"timers_sec\": {";
    for (int a = 0; a < 14; ++a) js << (a ? ", " : "") << "\"" << stepActs[a] << "\": " << timer_step.timing(a);
    js << "}, \"frame_sec\": [";
    for (size_t i = 0; i < frameSec.size(); ++i) js << (i ? ", " : "") << frameSec[i];
    js << "], \"frame_iters\": [";
    for (size_t i = 0; i < frameIters.size(); ++i) js << (i ? ", " : "") << frameIters[i];
    js << "]}";
    std::cout << js.str() << std::endl;
    if (!statsJson.empty()) { std::ofstream o(statsJson); o << js.str() << std::endl; }
    if (!dumpDir.empty()) {
        // copy iterStats next to the dumps so tests can replay the iteration log
        std::ifstream in(outputFolderPath + "iterStats.txt");
        std::ofstream out(dumpDir + "/iterStats.txt");
        out << in.rdbuf();
    }
    return 0;
}
