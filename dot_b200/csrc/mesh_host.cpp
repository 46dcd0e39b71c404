#include "mesh_host.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace dotgpu {

void mesh_features(int nV, int nT, const double* X, const int32_t* T, double YM, double PR, double rho, double* DmInv,
                   double* vol, double* mass, double* mu, double* lam) {
    for (int v = 0; v < nV; ++v) mass[v] = 0.0;
    const double mu0 = YM / 2.0 / (1.0 + PR);
    const double lam0 = YM * PR / (1.0 + PR) / (1.0 - 2.0 * PR);
    for (int t = 0; t < nT; ++t) {
        const double* p0 = X + 3 * (size_t)T[4 * t];
        const double* p1 = X + 3 * (size_t)T[4 * t + 1];
        const double* p2 = X + 3 * (size_t)T[4 * t + 2];
        const double* p3 = X + 3 * (size_t)T[4 * t + 3];
        // Dm columns = P2-P1, P3-P1, P4-P1 (Mesh.cpp:627-633); m[i][j] row i col j
        double m[3][3];
        for (int i = 0; i < 3; ++i) {
            m[i][0] = p1[i] - p0[i];
            m[i][1] = p2[i] - p0[i];
            m[i][2] = p3[i] - p0[i];
        }
        double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
        double c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
        double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
        double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
        double inv = 1.0 / det;
        double* B = DmInv + 9 * (size_t)t;
        B[0] = c00 * inv;
        B[1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * inv;
        B[2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * inv;
        B[3] = c01 * inv;
        B[4] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * inv;
        B[5] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * inv;
        B[6] = c02 * inv;
        B[7] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * inv;
        B[8] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * inv;
        vol[t] = det / 6.0;  // triArea (Mesh.cpp:639)
        mu[t] = mu0;
        lam[t] = lam0;
        // barycentric mass measured from the 4th vertex (Mesh.cpp:565-575)
        double a[3], b[3], c[3];
        for (int i = 0; i < 3; ++i) {
            a[i] = p0[i] - p3[i];
            b[i] = p1[i] - p3[i];
            c[i] = p2[i] - p3[i];
        }
        double cr0 = b[1] * c[2] - b[2] * c[1], cr1 = b[2] * c[0] - b[0] * c[2], cr2 = b[0] * c[1] - b[1] * c[0];
        double v = std::fabs(a[0] * cr0 + a[1] * cr1 + a[2] * cr2) / 6.0;
        for (int k = 0; k < 4; ++k) mass[T[4 * t + k]] += v / 4.0;
    }
    for (int v = 0; v < nV; ++v) mass[v] *= rho;
}

void vertex_adjacency(int nV, int nT, const int32_t* T, std::vector<int>& ptr, std::vector<int>& idx) {
    std::vector<int> cnt(nV + 1, 0);
    for (size_t i = 0; i < (size_t)4 * nT; ++i) cnt[T[i] + 1] += 3;
    for (int v = 0; v < nV; ++v) cnt[v + 1] += cnt[v];
    std::vector<int> raw(cnt[nV]), cur(cnt.begin(), cnt.end() - 1);
    for (int t = 0; t < nT; ++t)
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b)
                if (a != b) raw[cur[T[4 * t + a]]++] = T[4 * t + b];
    ptr.assign(nV + 1, 0);
    idx.clear();
    idx.reserve(raw.size() / 3);
    for (int v = 0; v < nV; ++v) {
        std::sort(raw.begin() + cnt[v], raw.begin() + cnt[v + 1]);
        int last = -1;
        for (int i = cnt[v]; i < cnt[v + 1]; ++i)
            if (raw[i] != last) {
                idx.push_back(raw[i]);
                last = raw[i];
            }
        ptr[v + 1] = (int)idx.size();
    }
}

void build_pattern(int nverts, const std::vector<int>& ap, const std::vector<int>& ai, const std::vector<uint8_t>& fixed,
                   MatrixPattern& out) {
    out.nverts = nverts;
    out.fixed = fixed;
    out.ia.assign(3 * (size_t)nverts + 1, 0);
    out.ja.clear();
    out.bptr.assign(nverts + 1, 0);
    out.bcol.clear();
    for (int v = 0; v < nverts; ++v) {
        if (fixed[v]) {
            for (int c = 0; c < 3; ++c) {
                out.ja.push_back(3 * v + c);
                out.ia[3 * v + c + 1] = (int32_t)out.ja.size();
            }
            out.bcol.push_back(v);
        } else {
            size_t b0 = out.bcol.size();
            out.bcol.push_back(v);
            for (int i = ap[v]; i < ap[v + 1]; ++i) {
                int w = ai[i];
                if (w > v && !fixed[w]) out.bcol.push_back(w);
            }
            size_t nb = out.bcol.size() - b0;
            for (int c = 0; c < 3; ++c) {
                for (size_t j = 0; j < nb; ++j) {
                    int w = out.bcol[b0 + j];
                    for (int d = (j == 0 ? c : 0); d < 3; ++d) out.ja.push_back(3 * w + d);
                }
                out.ia[3 * v + c + 1] = (int32_t)out.ja.size();
            }
        }
        out.bptr[v + 1] = (int32_t)out.bcol.size();
    }
}

namespace {
inline int find_block(const MatrixPattern& p, int v, int w) {
    auto b = p.bcol.begin() + p.bptr[v], e = p.bcol.begin() + p.bptr[v + 1];
    // first entry is v itself, the rest ascending and > v
    if (w == v) return p.bptr[v];
    auto it = std::lower_bound(b + 1, e, w);
    if (it == e || *it != w) return -1;
    return (int)(it - p.bcol.begin());
}

struct ListBuilder {
    std::vector<std::vector<int32_t>> rows;  // per block of the current block row
    void start(size_t nb) {
        rows.resize(nb);
        for (auto& r : rows) r.clear();
    }
    void flush(FillList& f) {
        for (auto& r : rows) {
            f.src.insert(f.src.end(), r.begin(), r.end());
            f.ptr.push_back((int64_t)f.src.size());
        }
    }
};
}  // namespace

void DDHost::build(int nV_, int nT_, const int32_t* T, const int32_t* epart, int k_, const uint8_t* fixed_mask,
                   const double* V_rest, double rho, const double* mass_global, bool with_fill, const std::vector<char>* sub_mask) {
    nV = nV_;
    nT = nT_;
    k = k_;
    if (k < 1) throw std::invalid_argument("k < 1");
    for (int t = 0; t < nT; ++t)
        if (epart[t] < 0 || epart[t] >= k) throw std::invalid_argument("element label out of range");
    std::vector<uint8_t> fixed(nV, 0);
    if (fixed_mask) fixed.assign(fixed_mask, fixed_mask + nV);

    // vFLoc of the global mesh (ascending tet)
    std::vector<int> vfp(nV + 1, 0), vfi((size_t)4 * nT);
    for (size_t i = 0; i < (size_t)4 * nT; ++i) vfp[T[i] + 1]++;
    for (int v = 0; v < nV; ++v) vfp[v + 1] += vfp[v];
    {
        std::vector<int> cur(vfp.begin(), vfp.end() - 1);
        for (int t = 0; t < nT; ++t)
            for (int c = 0; c < 4; ++c) vfi[cur[T[4 * t + c]]++] = 4 * t + c;
    }
    std::vector<int> gap, gai;
    vertex_adjacency(nV, nT, T, gap, gai);
    build_pattern(nV, gap, gai, fixed, gpat);

    // element lists and dup
    subs.assign(k, SubdomainHost());
    for (int t = 0; t < nT; ++t) subs[epart[t]].elems.push_back(t);
    dup.assign(nV, 0);
    std::vector<int> g2l(nV, -1);
    for (int s = 0; s < k; ++s) {
        SubdomainHost& sd = subs[s];
        sd.tets_local.resize(4 * sd.elems.size());
        for (size_t li = 0; li < sd.elems.size(); ++li)
            for (int c = 0; c < 4; ++c) {
                int g = T[4 * (size_t)sd.elems[li] + c];
                if (g2l[g] < 0) {
                    g2l[g] = (int)sd.l2g.size();
                    sd.l2g.push_back(g);
                }
                sd.tets_local[4 * li + c] = g2l[g];
            }
        for (int g : sd.l2g) {
            dup[g]++;
            g2l[g] = -1;
        }
    }

    // global fill list
    ListBuilder lb;
    if (with_fill) {
        gfill.ptr.assign(1, 0);
        gfill.src.clear();
        gfill.consts.assign(1, 1.0);
        for (int v = 0; v < nV; ++v) gfill.consts.push_back(mass_global ? mass_global[v] : 0.0);
        for (int v = 0; v < nV; ++v) {
            size_t nb = gpat.bptr[v + 1] - gpat.bptr[v];
            lb.start(nb);
            if (fixed[v]) {
                lb.rows[0].push_back(-1);  // identity (IglUtils.hpp:148-157)
            } else {
                for (int i = vfp[v]; i < vfp[v + 1]; ++i) {
                    int t = vfi[i] >> 2, a = vfi[i] & 3;
                    for (int b = 0; b < 4; ++b) {
                        int w = T[4 * (size_t)t + b];
                        if (fixed[w] || w < v) continue;
                        int blk = find_block(gpat, v, w);
                        lb.rows[blk - gpat.bptr[v]].push_back(16 * t + 4 * a + b);
                    }
                }
                lb.rows[0].push_back(-(1 + v) - 1);  // + m_v (DOTTimeStepper.cpp:597-607)
            }
            lb.flush(gfill);
        }
    }

    for (int s = 0; s < k; ++s) {
        SubdomainHost& sd = subs[s];
        if (sub_mask && !(*sub_mask)[s]) continue;
        const int nl = (int)sd.l2g.size(), ne = (int)sd.elems.size();
        for (int l = 0; l < nl; ++l) g2l[sd.l2g[l]] = l;
        std::vector<uint8_t> fl(nl, 0);
        for (int l = 0; l < nl; ++l)
            if (fixed[sd.l2g[l]]) fl[l] = 1;
        for (int l = 0; l < nl; ++l)
            if (fl[l]) sd.fixed_local.push_back(l);
        for (int l = 0; l < nl; ++l)
            if (dup[sd.l2g[l]] > 1) sd.iface.push_back(sd.l2g[l]);
        std::sort(sd.iface.begin(), sd.iface.end());
        // sub-mesh mass (Mesh::computeMassMatrix on the sub-mesh)
        sd.mass_local.assign(nl, 0.0);
        if (V_rest) {
            for (int li = 0; li < ne; ++li) {
                const int32_t* tt = T + 4 * (size_t)sd.elems[li];
                const double *p0 = V_rest + 3 * (size_t)tt[0], *p1 = V_rest + 3 * (size_t)tt[1], *p2 = V_rest + 3 * (size_t)tt[2],
                             *p3 = V_rest + 3 * (size_t)tt[3];
                double a[3], b[3], c[3];
                for (int i = 0; i < 3; ++i) {
                    a[i] = p0[i] - p3[i];
                    b[i] = p1[i] - p3[i];
                    c[i] = p2[i] - p3[i];
                }
                double cr0 = b[1] * c[2] - b[2] * c[1], cr1 = b[2] * c[0] - b[0] * c[2], cr2 = b[0] * c[1] - b[1] * c[0];
                double v = std::fabs(a[0] * cr0 + a[1] * cr1 + a[2] * cr2) / 6.0;
                for (int c4 = 0; c4 < 4; ++c4) sd.mass_local[sd.tets_local[4 * li + c4]] += v / 4.0;
            }
            for (int l = 0; l < nl; ++l) sd.mass_local[l] *= rho;
        }
        // vNeighborExt (ADMMDDTimeStepper.cpp:474-486)
        std::vector<int> lap, lai;
        vertex_adjacency(nl, ne, sd.tets_local.data(), lap, lai);
        {
            std::vector<std::vector<int>> extra(nl);
            bool any = false;
            for (int g : sd.iface) {
                int lv = g2l[g];
                for (int i = gap[g]; i < gap[g + 1]; ++i) {
                    int lu = g2l[gai[i]];
                    if (lu >= 0 && !std::binary_search(lai.begin() + lap[lv], lai.begin() + lap[lv + 1], lu)) {
                        extra[lv].push_back(lu);
                        any = true;
                    }
                }
            }
            if (any) {
                std::vector<int> np(nl + 1, 0), ni;
                ni.reserve(lai.size() + 64);
                for (int l = 0; l < nl; ++l) {
                    size_t b0 = ni.size();
                    ni.insert(ni.end(), lai.begin() + lap[l], lai.begin() + lap[l + 1]);
                    ni.insert(ni.end(), extra[l].begin(), extra[l].end());
                    std::sort(ni.begin() + b0, ni.end());
                    ni.erase(std::unique(ni.begin() + b0, ni.end()), ni.end());
                    np[l + 1] = (int)ni.size();
                }
                lap.swap(np);
                lai.swap(ni);
            }
        }
        build_pattern(nl, lap, lai, fl, sd.pat);

        if (with_fill) {
            // local vFLoc
            std::vector<int> lfp(nl + 1, 0), lfi((size_t)4 * ne);
            for (size_t i = 0; i < (size_t)4 * ne; ++i) lfp[sd.tets_local[i] + 1]++;
            for (int l = 0; l < nl; ++l) lfp[l + 1] += lfp[l];
            {
                std::vector<int> cur(lfp.begin(), lfp.end() - 1);
                for (int li = 0; li < ne; ++li)
                    for (int c = 0; c < 4; ++c) lfi[cur[sd.tets_local[4 * li + c]]++] = 4 * li + c;
            }
            FillList& f = sd.fill;
            f.ptr.assign(1, 0);
            f.src.clear();
            f.consts.assign(1, 1.0);
            for (int l = 0; l < nl; ++l) f.consts.push_back(sd.mass_local[l]);  // index 1+l
            for (int lv = 0; lv < nl; ++lv) {
                size_t nb = sd.pat.bptr[lv + 1] - sd.pat.bptr[lv];
                lb.start(nb);
                const int g = sd.l2g[lv];
                if (fl[lv]) {
                    lb.rows[0].push_back(-1);                 // setCoeff(.,.,1)
                    lb.rows[0].push_back(-(1 + lv) - 1);      // + m_local, also on fixed rows (DOTTimeStepper.cpp:677-686)
                } else {
                    for (int i = lfp[lv]; i < lfp[lv + 1]; ++i) {
                        int li = lfi[i] >> 2, a = lfi[i] & 3;
                        int t = sd.elems[li];
                        for (int b = 0; b < 4; ++b) {
                            int lu = sd.tets_local[4 * (size_t)li + b];
                            if (fl[lu] || lu < lv) continue;
                            int blk = find_block(sd.pat, lv, lu);
                            lb.rows[blk - sd.pat.bptr[lv]].push_back(16 * t + 4 * a + b);
                        }
                    }
                    lb.rows[0].push_back(-(1 + lv) - 1);
                    if (dup[g] > 1) {  // interface completion (DOTTimeStepper.cpp:696-791)
                        f.consts.push_back((mass_global ? mass_global[g] : 0.0) - sd.mass_local[lv]);
                        lb.rows[0].push_back(-(int)f.consts.size());
                        for (int i = vfp[g]; i < vfp[g + 1]; ++i) {
                            int t = vfi[i] >> 2, a = vfi[i] & 3;
                            if (epart[t] == s) continue;
                            lb.rows[0].push_back(16 * t + 4 * a + a);
                            for (int b = 0; b < 4; ++b) {
                                if (b == a) continue;
                                int u = T[4 * (size_t)t + b];
                                if (fixed[u] || dup[u] <= 1 || g2l[u] < 0) continue;
                                int lu = g2l[u];
                                if (lu < lv) continue;  // addCoeff drops row > col
                                int blk = find_block(sd.pat, lv, lu);
                                if (blk < 0) throw std::logic_error("interface block missing from pattern");
                                lb.rows[blk - sd.pat.bptr[lv]].push_back(16 * t + 4 * a + b);
                            }
                        }
                    }
                }
                lb.flush(f);
            }
        }
        for (int l = 0; l < nl; ++l) g2l[sd.l2g[l]] = -1;
    }
}

// SURVEY 8(f4), LBFGS-JH (LBFGSTimeStepper.cpp:60-92, 241-262, 318-331): block Jacobi of the global PD-projected Hessian over a NODE
// partition.  Block s = the nodes with npart == s (ascending, METIS::getNodeList), its matrix = the global matrix restricted to those
// nodes (LinSysSolver::getTriplets, LinSysSolver.hpp:256-284): EVERY tet incident to a block vertex contributes the 3x3 blocks between
// vertices of the block, the lumped mass sits on the free diagonals, fixed vertices are identity rows.  Expressed with the same ordered
// gather lists as the subdomain matrices, so fill / factorisation / solves / scatter run unchanged (dup == 1 everywhere).
void DDHost::build_node_blocks(int nV_, int nT_, const int32_t* T, const int32_t* npart, int k_, const uint8_t* fixed_mask,
                               const double* mass_global) {
    nV = nV_;
    nT = nT_;
    k = k_;
    if (k < 1) throw std::invalid_argument("k < 1");
    for (int v = 0; v < nV; ++v)
        if (npart[v] < 0 || npart[v] >= k) throw std::invalid_argument("node label out of range");
    std::vector<uint8_t> fixed(nV, 0);
    if (fixed_mask) fixed.assign(fixed_mask, fixed_mask + nV);
    std::vector<int> vfp(nV + 1, 0), vfi((size_t)4 * nT);
    for (size_t i = 0; i < (size_t)4 * nT; ++i) vfp[T[i] + 1]++;
    for (int v = 0; v < nV; ++v) vfp[v + 1] += vfp[v];
    {
        std::vector<int> cur(vfp.begin(), vfp.end() - 1);
        for (int t = 0; t < nT; ++t)
            for (int c = 0; c < 4; ++c) vfi[cur[T[4 * t + c]]++] = 4 * t + c;
    }
    std::vector<int> gap, gai;
    vertex_adjacency(nV, nT, T, gap, gai);
    build_pattern(nV, gap, gai, fixed, gpat);
    ListBuilder lb;
    // global matrix (the SpMV of initStepSize is not used by LBFGS-JH, but the checkers read it)
    gfill.ptr.assign(1, 0);
    gfill.src.clear();
    gfill.consts.assign(1, 1.0);
    for (int v = 0; v < nV; ++v) gfill.consts.push_back(mass_global ? mass_global[v] : 0.0);
    auto fill_rows = [&](const MatrixPattern& P, FillList& f, const std::vector<int32_t>& l2g, const std::vector<int>& g2l) {
        for (int lv = 0; lv < P.nverts; ++lv) {
            const int g = l2g[lv];
            lb.start((size_t)(P.bptr[lv + 1] - P.bptr[lv]));
            if (fixed[g]) {
                lb.rows[0].push_back(-1);  // identity (IglUtils.hpp:148-157)
            } else {
                for (int i = vfp[g]; i < vfp[g + 1]; ++i) {
                    const int t = vfi[i] >> 2, a = vfi[i] & 3;
                    for (int b = 0; b < 4; ++b) {
                        const int w = T[4 * (size_t)t + b];
                        const int lu = g2l[w];
                        if (fixed[w] || lu < lv) continue;     // lu < 0: outside the block
                        const int blk = find_block(P, lv, lu);
                        lb.rows[blk - P.bptr[lv]].push_back(16 * t + 4 * a + b);
                    }
                }
                lb.rows[0].push_back(-(1 + lv) - 1);  // + m_v
            }
            lb.flush(f);
        }
    };
    {
        std::vector<int32_t> ident(nV);
        std::vector<int> g2l(nV);
        for (int v = 0; v < nV; ++v) ident[v] = g2l[v] = v;
        fill_rows(gpat, gfill, ident, g2l);
    }
    subs.assign(k, SubdomainHost());
    dup.assign(nV, 1);
    for (int v = 0; v < nV; ++v) subs[npart[v]].l2g.push_back(v);  // ascending node ids
    std::vector<int> g2l(nV, -1);
    for (int s = 0; s < k; ++s) {
        SubdomainHost& sd = subs[s];
        const int nl = (int)sd.l2g.size();
        if (nl == 0) throw std::invalid_argument("empty node block");
        for (int l = 0; l < nl; ++l) g2l[sd.l2g[l]] = l;
        std::vector<uint8_t> fl(nl, 0);
        sd.mass_local.assign(nl, 0.0);
        for (int l = 0; l < nl; ++l) {
            fl[l] = fixed[sd.l2g[l]];
            if (fl[l]) sd.fixed_local.push_back(l);
            sd.mass_local[l] = mass_global ? mass_global[sd.l2g[l]] : 0.0;
        }
        std::vector<int> lap(nl + 1, 0), lai;
        for (int l = 0; l < nl; ++l) {
            const int g = sd.l2g[l];
            for (int i = gap[g]; i < gap[g + 1]; ++i)
                if (g2l[gai[i]] >= 0) lai.push_back(g2l[gai[i]]);  // ascending: l2g is ascending
            lap[l + 1] = (int)lai.size();
        }
        build_pattern(nl, lap, lai, fl, sd.pat);
        sd.fill.ptr.assign(1, 0);
        sd.fill.src.clear();
        sd.fill.consts.assign(1, 1.0);
        for (int l = 0; l < nl; ++l) sd.fill.consts.push_back(sd.mass_local[l]);
        fill_rows(sd.pat, sd.fill, sd.l2g, g2l);
        for (int l = 0; l < nl; ++l) g2l[sd.l2g[l]] = -1;
    }
}

}  // namespace dotgpu
