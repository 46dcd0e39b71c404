// DOT's METIS decomposition behind the C ABI.
//
// Reference: METIS<dim>::partMesh (src/Utils/METIS.hpp:109-160) with the option vector of initParam (:265-321), called from the
// ADMMDDTimeStepper constructor (src/TimeStepper/ADMMDDTimeStepper.cpp:88-92).  Subdomain labels must be bit-exact, so the
// partitioner is NOT re-implemented: this file dlopen()s libdotmetis.so = the METIS 5.1.0 the reference vendors
// (SuiteSparse/metis-5.1.0, IDXTYPEWIDTH 64 / REALTYPEWIDTH 32, metis.h:69,79), compiled by dot_b200/build.py from the sources
// where they lie, and calls METIS_PartMeshDual with exactly the reference's arguments.  Differences: dbglvl 0 instead of 511
// (printing only; labels are tested bit-for-bit against the reference wrapper's, tests/test_host_cpu.py).
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "mesh_host.h"

namespace dotgpu {
namespace {

typedef int64_t idx_t;   // IDXTYPEWIDTH 64
typedef float real_t;    // REALTYPEWIDTH 32
constexpr int kNOptions = 40;  // METIS_NOPTIONS (metis.h:190)
// moptions_et (metis.h:289-309)
enum { OPT_PTYPE = 0, OPT_OBJTYPE, OPT_CTYPE, OPT_IPTYPE, OPT_RTYPE, OPT_DBGLVL, OPT_NITER, OPT_NCUTS, OPT_SEED, OPT_NO2HOP, OPT_MINCONN,
       OPT_CONTIG, OPT_COMPRESS, OPT_CCORDER, OPT_PFACTOR, OPT_NSEPS, OPT_UFACTOR, OPT_NUMBERING };
constexpr idx_t PTYPE_KWAY = 1, OBJTYPE_CUT = 0, CTYPE_SHEM = 1, IPTYPE_METISRB = 4, RTYPE_GREEDY = 1;
constexpr int kMetisOK = 1;

typedef int (*fn_defaults)(idx_t*);
typedef int (*fn_partmeshdual)(idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, real_t*, idx_t*, idx_t*, idx_t*, idx_t*);
typedef int (*fn_vsep)(idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*);
typedef int (*fn_partmeshnodal)(idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, real_t*, idx_t*, idx_t*, idx_t*, idx_t*);

struct MetisLib {
    void* h = nullptr;
    fn_defaults defaults = nullptr;
    fn_partmeshdual part = nullptr;
    fn_partmeshnodal part_nodal = nullptr;
    fn_vsep vsep = nullptr;
    std::mutex mtx;   // METIS 5.1.0 seeds a process-global RNG per call: calls are serialised so that every call is reproducible
    std::string err;
};

MetisLib& metis() {
    static MetisLib L;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> cand;
        if (const char* e = std::getenv("DOTGPU_METIS_LIB")) cand.push_back(e);
        Dl_info info;
        if (dladdr((void*)&metis, &info) && info.dli_fname) {  // next to libdotgpu.so
            std::string p(info.dli_fname);
            size_t s = p.find_last_of('/');
            cand.push_back((s == std::string::npos ? std::string(".") : p.substr(0, s)) + "/libdotmetis.so");
        }
        cand.push_back("libdotmetis.so");
        for (const std::string& c : cand) {
            L.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (L.h) break;
            L.err = dlerror();
        }
        if (!L.h) return;
        L.defaults = (fn_defaults)dlsym(L.h, "METIS_SetDefaultOptions");
        L.part = (fn_partmeshdual)dlsym(L.h, "METIS_PartMeshDual");
        L.part_nodal = (fn_partmeshnodal)dlsym(L.h, "METIS_PartMeshNodal");
        L.vsep = (fn_vsep)dlsym(L.h, "METIS_ComputeVertexSeparator");
        if (!L.defaults || !L.part) L.err = "libdotmetis.so lacks METIS_SetDefaultOptions / METIS_PartMeshDual";
    });
    return L;
}

}  // namespace

// Multilevel vertex separator of a graph (METIS_ComputeVertexSeparator: MlevelNodeBisectionMultiple, the bisection step of METIS'
// own nested dissection) for the fill-reducing ordering of chol_symbolic.cpp.  part[i] = 0 / 1 (the two sides) or 2 (separator).
// Returns false when libdotmetis.so is not available (the caller falls back to its level-structure separators).
bool metis_vertex_separator(int n, const int64_t* xadj, const int64_t* adjncy, int64_t* part) {
    MetisLib& L = metis();
    if (!L.h || !L.defaults || !L.vsep) return false;
    std::lock_guard<std::mutex> lock(L.mtx);
    idx_t options[kNOptions];
    L.defaults(options);
    options[OPT_SEED] = -1;
    options[OPT_DBGLVL] = 0;
    idx_t nv = n, sepsize = 0;
    const int status = L.vsep(&nv, const_cast<idx_t*>((const idx_t*)xadj), const_cast<idx_t*>((const idx_t*)adjncy), nullptr, options, &sepsize,
                              (idx_t*)part);
    return status == kMetisOK;
}

static void metis_part(int nV, int nT, const int32_t* tets, int k, bool nodal, int32_t* out) {
    DG_REQUIRE(nV > 0 && nT > 0 && tets && out, "null or empty mesh");
    DG_REQUIRE(k >= 2, "the number of partitions must be at least 2 (METIS.hpp:300-302)");
    MetisLib& L = metis();
    std::lock_guard<std::mutex> lock(L.mtx);
    if (!L.h || !L.defaults || !L.part || !L.part_nodal)
        throw Error(DOTGPU_ERR_STATE, "libdotmetis.so (the reference's vendored METIS 5.1.0, built by dot_b200/build.py) is not available: " + L.err +
                                          "; pass labels produced elsewhere instead");
    // mesh in METIS' element-node form (METIS.hpp:89-105)
    std::vector<idx_t> eptr((size_t)nT + 1), eind((size_t)nT * 4);
    for (int t = 0; t < nT; ++t) {
        eptr[t] = 4 * (idx_t)t;
        for (int j = 0; j < 4; ++j) eind[4 * (size_t)t + j] = tets[4 * (size_t)t + j];
    }
    eptr[nT] = 4 * (idx_t)nT;
    idx_t options[kNOptions];
    L.defaults(options);
    options[OPT_PTYPE] = PTYPE_KWAY;
    options[OPT_OBJTYPE] = OBJTYPE_CUT;
    options[OPT_CTYPE] = CTYPE_SHEM;
    options[OPT_IPTYPE] = IPTYPE_METISRB;
    options[OPT_RTYPE] = RTYPE_GREEDY;
    options[OPT_MINCONN] = 1;
    options[OPT_CONTIG] = 1;
    options[OPT_NCUTS] = 3;
    options[OPT_NSEPS] = 3;
    options[OPT_NITER] = 10;
    options[OPT_DBGLVL] = 0;   // reference: 511 (prints every refinement move); no effect on the labels
    options[OPT_SEED] = -1;    // -> 4321 inside METIS (libmetis/util.c:23)
    options[OPT_UFACTOR] = 30;
    idx_t ne = nT, nn = nV, ncommon = 3, nparts = k, objval = 0;
    std::vector<idx_t> wgt((size_t)(nodal ? nV : nT), 1), epart((size_t)nT), npart((size_t)nV);
    std::vector<real_t> tpwgts((size_t)k);
    for (auto& w : tpwgts) w = (real_t)(1.0 / k);   // std::vector<real_t>(nparts, 1.0 / nparts)
    const int status = nodal ? L.part_nodal(&ne, &nn, eptr.data(), eind.data(), wgt.data(), nullptr, &nparts, tpwgts.data(), options, &objval,
                                            epart.data(), npart.data())
                             : L.part(&ne, &nn, eptr.data(), eind.data(), wgt.data(), nullptr, &ncommon, &nparts, tpwgts.data(), options, &objval,
                                      epart.data(), npart.data());
    if (status != kMetisOK) throw Error(DOTGPU_ERR_INVALID, "METIS failed with status " + std::to_string(status));
    if (nodal) for (int v = 0; v < nV; ++v) out[v] = (int32_t)npart[v];
    else for (int t = 0; t < nT; ++t) out[t] = (int32_t)epart[t];
}

void metis_partition(int nV, int nT, const int32_t* tets, int k, int32_t* epart_out) { metis_part(nV, nT, tets, k, false, epart_out); }
void metis_partition_nodes(int nV, int nT, const int32_t* tets, int k, int32_t* npart_out) { metis_part(nV, nT, tets, k, true, npart_out); }

}  // namespace dotgpu
