"""CPU-side tests: the C-ABI library loads and exports what include/dotgpu.h declares, the host logic
(mesh features, domain decomposition, patterns, scripted motion, symbolic analysis) matches the
reference's golden vectors / the oracle.  No compute call needs a GPU here."""
import os
import re

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import dot_b200 as D
from golden_util import Golden, rel
from oracle import dot_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["tiny_snh_k4_twist", "tiny_fcr_k4_twistnsns", "small_snh_k4_twist", "small_fcr_k3_stretch", "small_snh_k5_tsns_dt24",
         "tiny_fcr_inverted", "small_fcr_perturbed"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dotgpu.h")).read()
    names = sorted(set(re.findall(r"\b(dotgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    L = D.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.dotgpu_version() >= 100


def test_no_cpu_fallback_without_device():
    if D.device_count() > 0:
        pytest.skip("a CUDA device is present")
    g = Golden("tiny_snh_k4_twist")
    Dm, vol, mass, mu, lam = D.mesh_features(g["setup/V_rest"], g["setup/F"])
    with pytest.raises(D.DotGpuError) as e:
        D.Energy("SNH", g["setup/F"], Dm, vol, mu, lam, g["setup/V_rest"].shape[0])
    assert e.value.code == -2
    with pytest.raises(D.DotGpuError):
        D.Solver(g["setup/sbd0_ia"], g["setup/sbd0_ja"], device=0)
    fm = np.zeros(g["setup/V_rest"].shape[0], dtype=np.uint8)
    with pytest.raises(D.DotGpuError):
        D.Stepper(g["setup/V_rest"], g["setup/F"], g["setup/epart"], fm, k=4)


def test_bad_arguments_are_reported_not_fatal():
    with pytest.raises(D.DotGpuError) as e:
        D.DD(4, np.array([[0, 1, 2, 9]], dtype=np.int32), np.array([0], dtype=np.int32), 1)
    assert e.value.code == -1
    with pytest.raises(D.DotGpuError):
        D.DD(4, np.array([[0, 1, 2, 3]], dtype=np.int32), np.array([5], dtype=np.int32), 2)  # label out of range
    with pytest.raises(D.DotGpuError):
        D.Solver(np.array([0, 1, 2], dtype=np.int32), np.array([1, 1], dtype=np.int32), device=-1)  # row 0 lacks its diagonal


@pytest.mark.parametrize("name", CASES)
def test_mesh_features_match_reference(name):
    g = Golden(name)
    Dm, vol, mass, mu, lam = D.mesh_features(g["setup/V_rest"], g["setup/F"])
    assert rel(Dm, g["setup/restTriInv"]) < 1e-13
    assert rel(vol, g["setup/triArea"]) < 1e-13
    assert rel(mass, g["setup/mass"]) < 1e-13
    assert np.array_equal(mu, g["setup/mu"]) and np.array_equal(lam, g["setup/lambda"])


@pytest.mark.parametrize("name", CASES)
def test_domain_decomposition_bit_exact(name):
    g = Golden(name)
    nV = g["setup/V_rest"].shape[0]
    fm = np.zeros(nV, dtype=np.uint8)
    fm[g["setup/fixed"]] = 1
    dd = D.DD(nV, g["setup/F"], g["setup/epart"], g.k, fm)
    ia, ja = dd.pattern(-1)
    assert np.array_equal(ia, g["setup/global_ia"]) and np.array_equal(ja, g["setup/global_ja"])
    assert np.array_equal(dd.dup(), g["setup/dup"])
    for s in range(g.k):
        assert np.array_equal(dd.l2g(s), g["setup/sbd%d_l2g" % s])
        assert np.array_equal(dd.fixed_local(s), g["setup/sbd%d_fixed" % s])
        ia, ja = dd.pattern(s)
        assert np.array_equal(ia, g["setup/sbd%d_ia" % s]) and np.array_equal(ja, g["setup/sbd%d_ja" % s])


def test_single_subdomain_and_empty_fixed_set():
    g = Golden("tiny_snh_k4_twist")
    nV = g["setup/V_rest"].shape[0]
    T = g["setup/F"]
    dd = D.DD(nV, T, np.zeros(T.shape[0], dtype=np.int32), 1, None)
    assert np.array_equal(dd.dup(), np.ones(nV, dtype=np.int32))
    m = O.Mesh(g["setup/V_rest"], T)
    ia, ja = O.set_pattern(m.v_neighbor(), [])
    gia, gja = dd.pattern(-1)
    assert np.array_equal(ia, gia) and np.array_equal(ja, gja)
    # one subdomain: local numbering is first-touch, pattern is the permuted global one
    assert sorted(dd.l2g(0).tolist()) == list(range(nV))


@pytest.mark.parametrize("kind", ["twist", "stretch", "squash", "twistnstretch", "twistnsns", "twistnsns_old", "stretchnsquash", "null"])
def test_anim_scripter_matches_oracle(kind):
    g = Golden("small_snh_k4_twist")
    V = g["setup/V_rest"]
    a = D.Anim(kind, V)
    ref = O.AnimScripter(kind, V, O.border_verts(V, 0.01))
    assert np.array_equal(np.nonzero(a.fixed_mask())[0], ref.fixed())
    x = V.copy()
    xr = V.copy()
    for f in range(60):
        a.step(x, 0.025)
        xr = ref.step(xr, 0.025)
        assert np.abs(x - xr).max() < 1e-14, f
    if kind != "null":
        assert np.abs(x - V).max() > 1e-3


def test_anim_fixed_set_matches_reference():
    for name in ["tiny_snh_k4_twist", "small_fcr_k3_stretch", "small_snh_k5_tsns_dt24"]:
        g = Golden(name)
        a = D.Anim(g.meta["anim"], g["setup/V_rest"])
        assert np.array_equal(np.nonzero(a.fixed_mask())[0], g["setup/fixed"])


def _multifrontal_numpy(n, ia, ja, a, sym):
    """Checker: dense-front multifrontal Cholesky driven by libdotgpu's symbolic structure."""
    perm, sp, rp, rows, parent = sym["perm"], sym["super_ptr"], sym["row_ptr"], sym["rows"], sym["parent"]
    A = O.csr_upper_to_full(ia, ja, a).toarray()
    Ap = A[np.ix_(perm, perm)]
    ns = len(sp) - 1
    fronts = [None] * ns
    cbs = [None] * ns
    children = [[] for _ in range(ns)]
    for s in range(ns):
        if parent[s] >= 0:
            children[parent[s]].append(s)
    Lfull = np.zeros((n, n))
    for s in range(ns):
        r = rows[rp[s]:rp[s + 1]]
        k = sp[s + 1] - sp[s]
        assert np.array_equal(r[:k], np.arange(sp[s], sp[s + 1])) and (np.diff(r) > 0).all()
        F = np.zeros((len(r), len(r)))
        F[:, :k] = Ap[np.ix_(r, r[:k])]
        F[:k, :] = F[:, :k].T
        for c in children[s]:
            rc = rows[rp[c]:rp[c + 1]][sp[c + 1] - sp[c]:]
            pos = np.searchsorted(r, rc)
            assert np.array_equal(r[pos], rc), "child rows not contained in parent front"
            F[np.ix_(pos, pos)] += cbs[c]
        L11 = np.linalg.cholesky(F[:k, :k])
        L21 = np.linalg.solve(L11, F[:k, k:]).T
        cbs[s] = F[k:, k:] - L21 @ L21.T
        Lfull[np.ix_(r[:k], r[:k])] = L11
        Lfull[np.ix_(r[k:], r[:k])] = L21
        # A entries outside the front would be lost: check the structure covers the column
        col_nz = np.nonzero(Ap[sp[s]:, sp[s]:sp[s + 1]].any(axis=1))[0] + sp[s]
        assert np.isin(col_nz, r).all()
    return perm, Lfull


@pytest.mark.parametrize("name,sub", [("tiny_snh_k4_twist", 0), ("small_snh_k4_twist", 1), ("small_fcr_k3_stretch", 2),
                                      ("small_snh_k5_tsns_dt24", 4)])
def test_symbolic_analysis_drives_a_correct_factorisation(name, sub):
    g = Golden(name)
    st = g.states()[-1]
    ia, ja, a = g["setup/sbd%d_ia" % sub], g["setup/sbd%d_ja" % sub], g[st + "/sbd%d_a" % sub]
    s = D.Solver(ia, ja, device=-1)
    info = s.info()
    sym = s.symbolic()
    n = info.n
    assert sorted(sym["perm"].tolist()) == list(range(n))
    assert info.nsuper == len(sym["parent"]) and info.nlevels == sym["level"].max() + 1
    for c, p in enumerate(sym["parent"]):
        if p >= 0:
            assert p > c and sym["level"][p] > sym["level"][c]
    perm, L = _multifrontal_numpy(n, ia, ja, a, sym)
    A = O.csr_upper_to_full(ia, ja, a)
    b = np.random.default_rng(0).standard_normal(n)
    x = np.empty(n)
    y = np.linalg.solve(L, b[perm])
    x[perm] = np.linalg.solve(L.T, y)
    xr = spla.spsolve(A, b)
    assert rel(x, xr) < 1e-10
    assert info.nnz_l >= np.count_nonzero(np.abs(L) > 0)


def test_symbolic_on_global_pattern_is_sane():
    g = Golden("small_snh_k4_twist")
    ia, ja = g["setup/global_ia"], g["setup/global_ja"]
    s = D.Solver(ia, ja, device=-1)
    info = s.info()
    assert info.n == len(ia) - 1 and info.nsuper > 4 and info.nlevels >= 3
    assert info.nnz_l < 40 * info.nnz_a
    a = g[g.states()[-1] + "/global_a"]
    perm, L = _multifrontal_numpy(info.n, ia, ja, a, s.symbolic())
    A = O.csr_upper_to_full(ia, ja, a).toarray()
    assert rel(L @ L.T, A[np.ix_(perm, perm)]) < 1e-12


@pytest.mark.parametrize("mesh,k", [("bar5K_like", 6), ("bar17K_like", 8), ("bunny5K", 6), ("bar17K", 8), ("horse38K", 16)])
def test_partition_labels_bit_exact_with_reference_metis(mesh, k):
    """a14: dotgpu_partition (the reference's vendored METIS 5.1.0 behind the C ABI, option vector of Utils/METIS.hpp:265-321)
    reproduces the labels of the reference's own wrapper bit for bit - structured bars and the reference's input meshes."""
    import os
    from dot_b200 import meshgen
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if not os.path.exists(os.path.join(os.path.dirname(D.lib_path()), "libdotmetis.so")):
        pytest.skip("libdotmetis.so not built (needs the reference's vendored METIS sources at build time)")
    if mesh in meshgen.PRESETS:
        V, T = meshgen.preset(mesh)
    else:
        V, T = meshgen.load_mesh_npz(os.path.join(gold, "mesh_%s.npz" % mesh))
    ep = D.partition(V.shape[0], T, k)
    ref = np.load(os.path.join(gold, "labels_%s_k%d.npz" % (mesh, k)))["epart"]
    assert ep.dtype == np.int32 and np.array_equal(ep, ref)
    with pytest.raises(D.DotGpuError):
        D.partition(V.shape[0], T, 1)
