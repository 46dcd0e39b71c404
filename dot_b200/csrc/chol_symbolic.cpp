#include "chol_symbolic.h"

#include "mesh_host.h"

#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <stdexcept>

namespace dotgpu {

namespace {

struct Graph {
    int nn = 0;
    std::vector<int> ptr, idx;
};

struct NDOrder {
    const Graph& g;
    int leaf;
    int metis_min = 0;      // node sets of at least this size take their separator from METIS' multilevel bisection (0: never)
    std::vector<int> local_id;
    bool bisector = false;  // second separator candidate; measured on B200: -4 % nnz(L), -8 % flops but one more level and no gain in time
    std::vector<int> stamp, lev, queue;
    int cur_stamp = 0;
    std::vector<std::vector<int>> supers;  // emitted supernodes (node lists), children before parents

    NDOrder(const Graph& g_, int leaf_) : g(g_), leaf(leaf_), stamp(g_.nn, 0), lev(g_.nn, -1) {
        const char* e = std::getenv("DOTGPU_ND_BISECTOR");
        if (e) bisector = *e == '1';
        // Separators: the multilevel vertex separator of the reference's vendored METIS (libdotmetis.so, optional) where the node
        // set is large enough to matter; level-structure separators below that / without the library.  Measured:
        // nnz(L) -25 % on bar17K against the level-structure separators alone.  DOTGPU_ND_METIS=0 switches it off.
        metis_min = 4 * leaf_;
        if (const char* m = std::getenv("DOTGPU_ND_METIS")) metis_min = std::atoi(m) <= 0 ? 0 : std::max(std::atoi(m), 8);
        local_id.assign(g_.nn, -1);
    }

    // BFS inside the node set marked with `mark` in stamp[]; returns levels in lev[], order in queue
    int bfs(int start, int mark, int visited_mark) {
        queue.clear();
        queue.push_back(start);
        stamp[start] = visited_mark;
        lev[start] = 0;
        size_t head = 0;
        int maxlev = 0;
        while (head < queue.size()) {
            int v = queue[head++];
            for (int i = g.ptr[v]; i < g.ptr[v + 1]; ++i) {
                int w = g.idx[i];
                if (stamp[w] == mark) {
                    stamp[w] = visited_mark;
                    lev[w] = lev[v] + 1;
                    maxlev = std::max(maxlev, lev[w]);
                    queue.push_back(w);
                }
            }
        }
        return maxlev + 1;
    }

    // METIS separator of the subgraph induced by S; false if unavailable / degenerate
    bool metis_separator(const std::vector<int>& S, std::vector<int>& sep) {
        const int n = (int)S.size();
        for (int i = 0; i < n; ++i) local_id[S[i]] = i;
        std::vector<int64_t> xadj(n + 1, 0), adj, part(n, 0);
        for (int i = 0; i < n; ++i) {
            const int v = S[i];
            for (int k = g.ptr[v]; k < g.ptr[v + 1]; ++k)
                if (local_id[g.idx[k]] >= 0) adj.push_back(local_id[g.idx[k]]);
            xadj[i + 1] = (int64_t)adj.size();
        }
        for (int i = 0; i < n; ++i) local_id[S[i]] = -1;
        if (adj.empty() || !metis_vertex_separator(n, xadj.data(), adj.data(), part.data())) return false;
        int cnt[3] = {0, 0, 0};
        for (int i = 0; i < n; ++i) cnt[part[i] < 0 || part[i] > 2 ? 2 : part[i]]++;
        if (cnt[2] == 0 || cnt[0] == 0 || cnt[1] == 0 || cnt[2] * 2 > n) return false;
        sep.clear();
        for (int i = 0; i < n; ++i)
            if (part[i] == 2) sep.push_back(S[i]);
        return true;
    }

    // S minus the separator: connected components are ordered recursively, then the separator becomes a supernode
    void split_and_recurse(std::vector<int>& S, std::vector<int>& sep) {
        const int mrest = ++cur_stamp;
        for (int v : S) stamp[v] = mrest;
        const int msep = ++cur_stamp;
        for (int v : sep) stamp[v] = msep;
        std::vector<std::vector<int>> comps;
        for (int v : S) {
            if (stamp[v] != mrest) continue;
            int mc = ++cur_stamp;
            std::vector<int> comp;
            comp.push_back(v);
            stamp[v] = mc;
            size_t head = 0;
            while (head < comp.size()) {
                int a = comp[head++];
                for (int i = g.ptr[a]; i < g.ptr[a + 1]; ++i) {
                    int w = g.idx[i];
                    if (stamp[w] == mrest) {
                        stamp[w] = mc;
                        comp.push_back(w);
                    }
                }
            }
            comps.push_back(std::move(comp));
        }
        std::vector<int>().swap(S);
        for (auto& c : comps) order(c);
        if (!sep.empty()) supers.push_back(std::move(sep));
    }

    void order(std::vector<int>& S) {  // S connected
        if ((int)S.size() <= leaf) {
            supers.push_back(S);
            return;
        }
        if (metis_min > 0 && (int)S.size() >= metis_min) {
            std::vector<int> msep;
            if (metis_separator(S, msep)) {
                split_and_recurse(S, msep);
                return;
            }
        }
        int m0 = ++cur_stamp;
        for (int v : S) stamp[v] = m0;
        int m1 = ++cur_stamp;
        bfs(S[0], m0, m1);
        int far = queue.back();
        int m2 = ++cur_stamp;
        int nlev = bfs(far, m1, m2);  // queue.back() = the other end of the pseudo-diameter
        if (nlev <= 2) {
            supers.push_back(S);
            return;
        }
        std::vector<int> cnt(nlev, 0);
        for (int v : S) cnt[lev[v]]++;
        const int total = (int)S.size();
        int best = -1, best_size = 1 << 30;
        int before = cnt[0];
        for (int L = 1; L <= nlev - 2; ++L) {
            int after = total - before - cnt[L];
            if (before >= 0.3 * total && after >= 0.3 * total && cnt[L] < best_size) {
                best = L;
                best_size = cnt[L];
            }
            before += cnt[L];
        }
        if (best < 0) {
            int bestd = 1 << 30;
            before = cnt[0];
            for (int L = 1; L <= nlev - 2; ++L) {
                int after = total - before - cnt[L];
                int d = std::abs(before - after);
                if (d < bestd) {
                    bestd = d;
                    best = L;
                }
                before += cnt[L];
            }
        }
        // candidate 1: nodes of level `best` that touch level best+1 (the rest of the level stays on the near side)
        std::vector<int> sep;
        int msep = ++cur_stamp;
        for (int v : S) {
            if (lev[v] != best) continue;
            bool touches = false;
            for (int i = g.ptr[v]; i < g.ptr[v + 1] && !touches; ++i) {
                int w = g.idx[i];
                if (stamp[w] == m2 && lev[w] == best + 1) touches = true;
            }
            if (touches) sep.push_back(v);
        }
        // candidate 2 ("bisector"): distances from BOTH ends of the pseudo-diameter; a node belongs to the end it is closer to,
        // the separator is the smaller of the two boundary layers.  Level shells around one end are curved; the bisector of two
        // far-apart ends is close to a plane and usually thinner on blob-shaped subdomains.
        if (bisector) {
            std::vector<int> dist_b(S.size());
            for (size_t i = 0; i < S.size(); ++i) dist_b[i] = lev[S[i]];
            const int far2 = queue.back();
            int m3 = ++cur_stamp;
            bfs(far2, m2, m3);  // lev[] = distance from the other end; all nodes of S now carry stamp m3
            std::vector<char> side(S.size());
            int nA = 0;
            for (size_t i = 0; i < S.size(); ++i) {
                side[i] = lev[S[i]] < dist_b[i] ? 0 : 1;  // 0: closer to far2
                nA += side[i] == 0;
            }
            const int nB = total - nA;
            if (nA >= 0.3 * total && nB >= 0.3 * total) {
                // mark sides through lev[] (0/1) for neighbour tests
                for (size_t i = 0; i < S.size(); ++i) lev[S[i]] = side[i];
                std::vector<int> bA, bB;
                for (size_t i = 0; i < S.size(); ++i) {
                    const int v = S[i];
                    bool touches = false;
                    for (int k = g.ptr[v]; k < g.ptr[v + 1] && !touches; ++k) {
                        int w = g.idx[k];
                        if (stamp[w] == m3 && lev[w] != side[i]) touches = true;
                    }
                    if (touches) (side[i] == 0 ? bA : bB).push_back(v);
                }
                std::vector<int>& cand = bA.size() <= bB.size() ? bA : bB;
                if (!cand.empty() && cand.size() < sep.size()) sep.swap(cand);
            }
            m2 = m3;  // every node of S carries stamp m3 now
        }
        for (int v : sep) stamp[v] = msep;
        // connected components of S \ sep
        int mrest = m2;
        std::vector<std::vector<int>> comps;
        for (int v : S) {
            if (stamp[v] != mrest) continue;
            int mc = ++cur_stamp;
            std::vector<int> comp;
            comp.push_back(v);
            stamp[v] = mc;
            size_t head = 0;
            while (head < comp.size()) {
                int a = comp[head++];
                for (int i = g.ptr[a]; i < g.ptr[a + 1]; ++i) {
                    int w = g.idx[i];
                    if (stamp[w] == mrest) {
                        stamp[w] = mc;
                        comp.push_back(w);
                    }
                }
            }
            comps.push_back(std::move(comp));
        }
        std::vector<int>().swap(S);
        for (auto& c : comps) order(c);
        if (!sep.empty()) supers.push_back(std::move(sep));
    }
};

}  // namespace

void Symbolic::analyze(int n_, const int32_t* ia, const int32_t* ja, int leaf_nodes) {
    n = n_;
    if (n <= 0) throw std::invalid_argument("empty matrix");
    const int B = (n % 3 == 0) ? 3 : 1;
    const int nn = n / B;
    // ---- node graph (symmetric, no self loops) ----
    Graph g;
    g.nn = nn;
    {
        std::vector<std::pair<int, int>> edges;
        edges.reserve((size_t)ia[n] / (B * B) * 2 + 16);
        for (int i = 0; i < n; ++i) {
            if (ia[i + 1] <= ia[i] || ja[ia[i]] != i) throw std::invalid_argument("pattern row without leading diagonal entry");
            int a = i / B, last = -1;
            for (int k = ia[i]; k < ia[i + 1]; ++k) {
                int j = ja[k];
                if (j < i || j >= n) throw std::invalid_argument("pattern is not upper triangular CSR");
                int b = j / B;
                if (b != a && b != last) {
                    edges.emplace_back(a, b);
                    edges.emplace_back(b, a);
                    last = b;
                }
            }
        }
        std::sort(edges.begin(), edges.end());
        edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
        g.ptr.assign(nn + 1, 0);
        for (auto& e : edges) g.ptr[e.first + 1]++;
        for (int v = 0; v < nn; ++v) g.ptr[v + 1] += g.ptr[v];
        g.idx.resize(edges.size());
        for (size_t i = 0; i < edges.size(); ++i) g.idx[i] = edges[i].second;  // sorted by (first, second)
    }
    // ---- nested dissection per connected component ----
    if (const char* e = std::getenv("DOTGPU_ND_LEAF")) leaf_nodes = std::max(2, std::atoi(e));  // experiments
    NDOrder nd(g, leaf_nodes);
    {
        std::vector<char> seen(nn, 0);
        std::vector<int> comp;
        for (int v = 0; v < nn; ++v) {
            if (seen[v]) continue;
            comp.clear();
            comp.push_back(v);
            seen[v] = 1;
            size_t head = 0;
            while (head < comp.size()) {
                int a = comp[head++];
                for (int i = g.ptr[a]; i < g.ptr[a + 1]; ++i) {
                    int w = g.idx[i];
                    if (!seen[w]) {
                        seen[w] = 1;
                        comp.push_back(w);
                    }
                }
            }
            std::vector<int> S(comp);
            nd.order(S);
        }
    }
    // ---- node-level structure of a supernode list: numbering, below-row sets, supernodal etree ----
    std::vector<int> newnode(nn, -1), snode_of_node(nn, -1), node_ptr, snode_of_newnode(nn);
    std::vector<std::vector<int>> rown;
    auto structure = [&](const std::vector<std::vector<int>>& supers) {
        const int ns_ = (int)supers.size();
        node_ptr.assign(ns_ + 1, 0);
        int c = 0;
        for (int s = 0; s < ns_; ++s) {
            for (int v : supers[s]) {
                newnode[v] = c++;
                snode_of_node[v] = s;
            }
            node_ptr[s + 1] = c;
        }
        if (c != nn) throw std::logic_error("ordering lost nodes");
        for (int v = 0; v < nn; ++v) snode_of_newnode[newnode[v]] = snode_of_node[v];
        rown.assign(ns_, std::vector<int>());
        std::vector<std::vector<int>> pending(ns_);
        parent.assign(ns_, -1);
        std::vector<int> mark(nn, -1);
        for (int s = 0; s < ns_; ++s) {
            std::vector<int>& r = rown[s];
            const int last = node_ptr[s + 1] - 1;
            for (int v : supers[s]) {
                for (int i = g.ptr[v]; i < g.ptr[v + 1]; ++i) {
                    int w = newnode[g.idx[i]];
                    if (w > last && mark[w] != s) {
                        mark[w] = s;
                        r.push_back(w);
                    }
                }
            }
            for (int w : pending[s])
                if (w > last && mark[w] != s) {
                    mark[w] = s;
                    r.push_back(w);
                }
            std::vector<int>().swap(pending[s]);
            std::sort(r.begin(), r.end());
            if (!r.empty()) {
                int p = snode_of_newnode[r[0]];
                parent[s] = p;
                const int plast = node_ptr[p + 1] - 1;
                for (int w : r)
                    if (w > plast) pending[p].push_back(w);
            }
        }
    };
    // ---- chain amalgamation without fill: an only child whose update rows are exactly its parent's front (columns + rows)
    //      forms one dense supernode with it (CHOLMOD's fundamental-supernode rule, cholmod_super_symbolic.c:393-458).  Every
    //      merge removes one dependent level from the factorisation and from both solve sweeps. ----
    for (;;) {
        structure(nd.supers);
        const int ns_ = (int)nd.supers.size();
        std::vector<int> nchild(ns_, 0);
        for (int s = 0; s < ns_; ++s)
            if (parent[s] >= 0) nchild[parent[s]]++;
        bool merged = false;
        std::vector<std::vector<int>> out;
        out.reserve(ns_);
        for (int s = 0; s < ns_; ++s) {
            const int p = parent[s];
            if (p == s + 1 && nchild[p] == 1 &&
                rown[s].size() == nd.supers[p].size() + rown[p].size()) {
                nd.supers[p].insert(nd.supers[p].begin(), nd.supers[s].begin(), nd.supers[s].end());
                merged = true;  // s disappears into p; p may merge further in the next pass
            } else {
                out.push_back(std::move(nd.supers[s]));
            }
        }
        nd.supers.swap(out);
        if (!merged) break;
    }
    nsuper = (int)nd.supers.size();
    perm.resize(n);
    iperm.resize(n);
    for (int v = 0; v < nn; ++v)
        for (int c = 0; c < B; ++c) {
            perm[B * newnode[v] + c] = B * v + c;
            iperm[B * v + c] = B * newnode[v] + c;
        }
    super_ptr.resize(nsuper + 1);
    for (int s = 0; s <= nsuper; ++s) super_ptr[s] = B * node_ptr[s];
    // ---- expand to scalar rows ----
    row_ptr.assign(nsuper + 1, 0);
    for (int s = 0; s < nsuper; ++s)
        row_ptr[s + 1] = row_ptr[s] + (int64_t)B * ((node_ptr[s + 1] - node_ptr[s]) + (int64_t)rown[s].size());
    rows.resize(row_ptr[nsuper]);
    rel.assign(row_ptr[nsuper], -1);
    for (int s = 0; s < nsuper; ++s) {
        int64_t o = row_ptr[s];
        for (int c = super_ptr[s]; c < super_ptr[s + 1]; ++c) rows[o++] = c;
        for (int w : rown[s])
            for (int c = 0; c < B; ++c) rows[o++] = B * w + c;
    }
    // relative indices into the parent front
    for (int s = 0; s < nsuper; ++s) {
        int p = parent[s];
        if (p < 0) continue;
        const int32_t* pr = rows.data() + row_ptr[p];
        const int pm = (int)(row_ptr[p + 1] - row_ptr[p]);
        const int ns = super_ptr[s + 1] - super_ptr[s];
        int pos = 0;
        for (int64_t o = row_ptr[s] + ns; o < row_ptr[s + 1]; ++o) {
            while (pos < pm && pr[pos] < rows[o]) ++pos;
            if (pos >= pm || pr[pos] != rows[o]) throw std::logic_error("child row missing in parent front");
            rel[o] = pos;
        }
    }
    // ---- levels, children ----
    level.assign(nsuper, 0);
    for (int s = 0; s < nsuper; ++s)
        if (parent[s] >= 0) level[parent[s]] = std::max(level[parent[s]], level[s] + 1);
    nlevels = 0;
    for (int s = 0; s < nsuper; ++s) nlevels = std::max(nlevels, level[s] + 1);
    level_ptr.assign(nlevels + 1, 0);
    for (int s = 0; s < nsuper; ++s) level_ptr[level[s] + 1]++;
    for (int l = 0; l < nlevels; ++l) level_ptr[l + 1] += level_ptr[l];
    level_list.resize(nsuper);
    {
        std::vector<int> cur(level_ptr.begin(), level_ptr.end() - 1);
        for (int s = 0; s < nsuper; ++s) level_list[cur[level[s]]++] = s;
    }
    child_ptr.assign(nsuper + 1, 0);
    for (int s = 0; s < nsuper; ++s)
        if (parent[s] >= 0) child_ptr[parent[s] + 1]++;
    for (int s = 0; s < nsuper; ++s) child_ptr[s + 1] += child_ptr[s];
    child_list.resize(child_ptr[nsuper]);
    {
        std::vector<int> cur(child_ptr.begin(), child_ptr.end() - 1);
        for (int s = 0; s < nsuper; ++s)
            if (parent[s] >= 0) child_list[cur[parent[s]]++] = s;
    }
    // ---- storage offsets, statistics ----
    panel_off.assign(nsuper + 1, 0);
    cb_off.assign(nsuper + 1, 0);
    u_off.assign(nsuper + 1, 0);
    flops = 0.0;
    max_front = max_nscol = 0;
    for (int s = 0; s < nsuper; ++s) {
        int64_t m = row_ptr[s + 1] - row_ptr[s], ns = super_ptr[s + 1] - super_ptr[s], nb = m - ns;
        panel_off[s + 1] = panel_off[s] + m * ns;
        cb_off[s + 1] = cb_off[s] + nb * nb;
        u_off[s + 1] = u_off[s] + nb;
        flops += (double)ns * ns * ns / 3.0 + (double)nb * ns * ns + (double)nb * nb * ns;
        max_front = std::max<int>(max_front, (int)m);
        max_nscol = std::max<int>(max_nscol, (int)ns);
    }
    nnz_l = panel_off[nsuper];
    // ---- scatter map of A into the panels ----
    std::vector<int> snode_of_col(n);
    for (int s = 0; s < nsuper; ++s)
        for (int c = super_ptr[s]; c < super_ptr[s + 1]; ++c) snode_of_col[c] = s;
    amap.resize(ia[n]);
    for (int i = 0; i < n; ++i) {
        for (int k = ia[i]; k < ia[i + 1]; ++k) {
            int pi = iperm[i], pj = iperm[ja[k]];
            int c = std::min(pi, pj), r = std::max(pi, pj);
            int s = snode_of_col[c];
            const int32_t* b = rows.data() + row_ptr[s];
            const int32_t* e = rows.data() + row_ptr[s + 1];
            const int32_t* it = std::lower_bound(b, e, r);
            if (it == e || *it != r) throw std::logic_error("matrix entry outside the symbolic structure");
            amap[k] = panel_off[s] + (int64_t)(it - b) * (super_ptr[s + 1] - super_ptr[s]) + (c - super_ptr[s]);
        }
    }
    // ---- extend-add gather lists (forward solve) ----
    ea_ptr.assign(row_ptr[nsuper] + 1, 0);
    for (int s = 0; s < nsuper; ++s) {
        int p = parent[s];
        if (p < 0) continue;
        const int ns = super_ptr[s + 1] - super_ptr[s];
        for (int64_t o = row_ptr[s] + ns; o < row_ptr[s + 1]; ++o) ea_ptr[row_ptr[p] + rel[o] + 1]++;
    }
    for (int64_t i = 0; i < row_ptr[nsuper]; ++i) ea_ptr[i + 1] += ea_ptr[i];
    ea_src.resize(ea_ptr[row_ptr[nsuper]]);
    {
        std::vector<int64_t> cur(ea_ptr.begin(), ea_ptr.end() - 1);
        for (int s = 0; s < nsuper; ++s) {  // ascending child order => deterministic summation order
            int p = parent[s];
            if (p < 0) continue;
            const int ns = super_ptr[s + 1] - super_ptr[s];
            for (int64_t o = row_ptr[s] + ns; o < row_ptr[s + 1]; ++o)
                ea_src[cur[row_ptr[p] + rel[o]]++] = u_off[s] + (o - row_ptr[s] - ns);
        }
    }
}

}  // namespace dotgpu
