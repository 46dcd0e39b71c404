"""TEST INFRASTRUCTURE - CPU restatement (numpy/scipy) of DOT's per-frame optimisation time step.

Not part of the product path: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.  Parity is PINNED: tests/test_oracle_golden.py
checks every function here against tests/golden/*.npz, which oracle/gen_golden.py produced
by running the unmodified reference (oracle/_ref/dot_ref) in the build container.

Each function cites the reference lines it restates (paths relative to /root/reference/src).
All arithmetic fp64, indices int32/int64.  Written for clarity and small meshes; it is
vectorised where that is free, loops where the reference's std::set/std::map order matters.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

FCR, SNH = "FCR", "SNH"


# ---------------------------------------------------------------------------------------
# a1  mesh features  (Mesh.cpp:589-700 computeFeatures, :552-585 mass, :741-744 Lame)
# ---------------------------------------------------------------------------------------
@dataclass
class Mesh:
    V_rest: np.ndarray            # [nV,3]
    T: np.ndarray                 # [nT,4] int
    YM: float = 1e5
    PR: float = 0.4
    rho: float = 1000.0
    DmInv: np.ndarray = field(init=False)   # restTriInv [nT,3,3]
    vol: np.ndarray = field(init=False)     # triArea (signed rest volume)
    mass: np.ndarray = field(init=False)    # lumped, [nV]
    mu: np.ndarray = field(init=False)
    lam: np.ndarray = field(init=False)

    def __post_init__(self):
        X, T = self.V_rest, self.T
        Dm = np.stack([X[T[:, 1]] - X[T[:, 0]], X[T[:, 2]] - X[T[:, 0]], X[T[:, 3]] - X[T[:, 0]]], axis=2)
        self.DmInv = np.linalg.inv(Dm)
        self.vol = np.linalg.det(Dm) / 6.0
        # barycentric lumped mass with |volume| measured from vertex 3 (Mesh.cpp:565-575)
        a, b, c = X[T[:, 0]] - X[T[:, 3]], X[T[:, 1]] - X[T[:, 3]], X[T[:, 2]] - X[T[:, 3]]
        v = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c))) / 6.0
        m = np.zeros(X.shape[0])
        for k in range(4):
            np.add.at(m, T[:, k], v / 4.0)
        self.mass = m * self.rho
        nT = T.shape[0]
        self.mu = np.full(nT, self.YM / 2.0 / (1.0 + self.PR))
        self.lam = np.full(nT, self.YM * self.PR / (1.0 + self.PR) / (1.0 - 2.0 * self.PR))

    @property
    def nV(self):
        return self.V_rest.shape[0]

    @property
    def nT(self):
        return self.T.shape[0]

    def v_neighbor(self):
        """vNeighbor as sorted lists (std::set order), Mesh.cpp:681-690."""
        nb = [set() for _ in range(self.nV)]
        for t in self.T:
            for i in range(4):
                for j in range(i + 1, 4):
                    nb[t[i]].add(int(t[j]))
                    nb[t[j]].add(int(t[i]))
        return [sorted(s) for s in nb]

    def v_floc(self):
        """vFLoc[v] = ascending (tet, local) pairs, Mesh.cpp:606-611."""
        out = [[] for _ in range(self.nV)]
        for t in range(self.nT):
            for k in range(4):
                out[self.T[t, k]].append((t, k))
        return out


# ---------------------------------------------------------------------------------------
# a2  deformation gradient (Energy.cpp:457-473)
# ---------------------------------------------------------------------------------------
def deformation_gradient(mesh: Mesh, x: np.ndarray) -> np.ndarray:
    T = mesh.T
    Ds = np.stack([x[T[:, 1]] - x[T[:, 0]], x[T[:, 2]] - x[T[:, 0]], x[T[:, 3]] - x[T[:, 0]]], axis=2)
    return Ds @ mesh.DmInv


# ---------------------------------------------------------------------------------------
# a3/a4  3x3 SVD with the reference's conventions: U S V^T = F, det U = det V = +1,
# |s1|>=|s2|>=|s3|, sign carried by s3 (IglUtils.cpp:929-1085; AutoFlipSVD.hpp:34-81)
# ---------------------------------------------------------------------------------------
SVD_TOL = 8.0 * np.finfo(float).eps


def svd_lapack(F: np.ndarray):
    """LAPACK SVD + sign fix-up; cross-check only (see svd_rot for why it is not the default)."""
    U, s, Vt = np.linalg.svd(F)
    V = np.swapaxes(Vt, 1, 2).copy()
    U = U.copy()
    s = s.copy()
    fu = np.linalg.det(U) < 0
    U[fu, :, 2] *= -1
    s[fu, 2] *= -1
    fv = np.linalg.det(V) < 0
    V[fv, :, 2] *= -1
    s[fv, 2] *= -1
    return U, s, V


def svd_rot(F: np.ndarray, max_sweeps: int = 12):
    """Batched Jacobi SVD.  The reference diagonalises A^T A by cyclic Jacobi sweeps in the order
    (1,2),(1,3),(2,3), forms B = A V, sorts the columns by norm and gets U, sigma by a QR of B
    (SVD_EFTYCHIOS/Singular_Value_Decomposition_Main_Kernel_Body.hpp:51-85, 91, 577-870, 944-1151).
    Restated here as the one-sided (Hestenes) form of the same iteration - rotations are chosen from
    the entries of (AV)^T(AV) but applied to the columns of A V directly, which keeps full relative
    accuracy - followed by the same sort and a Gram-Schmidt/cross-product QR that puts the sign on
    sigma_3.  A rotation is skipped when |a_p.a_q| <= 8 eps |a_p||a_q|, so an (numerically) identity
    F returns sigma = (1,1,1) exactly, as the reference's does: this matters because the reference's
    2x2 PD projection (IglUtils.hpp:270-309) is discontinuous at a zero eigenvalue, which is exactly
    where every B block sits in the rest state.
    """
    A = np.array(F, dtype=float, copy=True)
    n = A.shape[0]
    V = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    for _ in range(max_sweeps):
        any_rot = False
        for p, q in ((0, 1), (0, 2), (1, 2)):
            ap, aq = A[:, :, p], A[:, :, q]
            alpha = (ap * ap).sum(axis=1)
            beta = (aq * aq).sum(axis=1)
            gamma = (ap * aq).sum(axis=1)
            act = np.abs(gamma) > SVD_TOL * np.sqrt(alpha * beta)
            if not act.any():
                continue
            any_rot = True
            g = np.where(act, gamma, 1.0)
            zeta = (beta - alpha) / (2.0 * g)
            t = np.where(zeta >= 0.0, 1.0, -1.0) / (np.abs(zeta) + np.sqrt(1.0 + zeta * zeta))
            c = 1.0 / np.sqrt(1.0 + t * t)
            sn = c * t
            c = np.where(act, c, 1.0)[:, None]
            sn = np.where(act, sn, 0.0)[:, None]
            A[:, :, p], A[:, :, q] = c * ap - sn * aq, sn * ap + c * aq
            vp, vq = V[:, :, p].copy(), V[:, :, q].copy()
            V[:, :, p], V[:, :, q] = c * vp - sn * vq, sn * vp + c * vq
        if not any_rot:
            break
    # sort columns by norm, descending; keep det V = +1
    nrm2 = (A * A).sum(axis=1)
    for p, q in ((0, 1), (0, 2), (1, 2)):       # 3-element sorting network
        sw = nrm2[:, p] < nrm2[:, q]
        if sw.any():
            # swap p<->q and negate the column that lands in q (a proper rotation)
            Ap, Aq = A[sw, :, p].copy(), A[sw, :, q].copy()
            A[sw, :, p], A[sw, :, q] = Aq, -Ap
            Vp, Vq = V[sw, :, p].copy(), V[sw, :, q].copy()
            V[sw, :, p], V[sw, :, q] = Vq, -Vp
            np_, nq_ = nrm2[sw, p].copy(), nrm2[sw, q].copy()
            nrm2[sw, p], nrm2[sw, q] = nq_, np_
    s = np.zeros((n, 3))
    U = np.zeros((n, 3, 3))
    s0 = np.sqrt(nrm2[:, 0])
    ok0 = s0 > 0.0
    u0 = np.where(ok0[:, None], A[:, :, 0] / np.where(ok0, s0, 1.0)[:, None], np.array([1.0, 0.0, 0.0]))
    b1 = A[:, :, 1] - (u0 * A[:, :, 1]).sum(axis=1)[:, None] * u0
    n1 = np.sqrt((b1 * b1).sum(axis=1))
    ok1 = n1 > 1e-300
    # rank <= 1: any unit vector orthogonal to u0 (axis least aligned with u0)
    k = np.argmin(np.abs(u0), axis=1)
    e = np.eye(3)[k]
    alt = e - (u0 * e).sum(axis=1)[:, None] * u0
    alt /= np.sqrt((alt * alt).sum(axis=1))[:, None]
    u1 = np.where(ok1[:, None], b1 / np.where(ok1, n1, 1.0)[:, None], alt)
    u2 = np.cross(u0, u1)
    s[:, 0] = s0
    s[:, 1] = (u1 * A[:, :, 1]).sum(axis=1)
    s[:, 2] = (u2 * A[:, :, 2]).sum(axis=1)
    U[:, :, 0], U[:, :, 1], U[:, :, 2] = u0, u1, u2
    return U, s, V


# ---------------------------------------------------------------------------------------
# a5  sigma-space energies and derivatives
#   FCR: FixedCoRotEnergy.cpp:83-172   SNH: StableNHEnergy.cpp:80-230 (SNH_WITHLOG off)
# ---------------------------------------------------------------------------------------
def psi(energy, s, mu, lam):
    J = s[:, 0] * s[:, 1] * s[:, 2]
    if energy == FCR:
        return mu * ((s - 1.0) ** 2).sum(axis=1) + lam / 2.0 * (J - 1.0) ** 2
    JmA = J - (1.0 + mu / lam)
    return (mu * ((s ** 2).sum(axis=1) - 3.0) + lam * JmA * JmA) / 2.0


def dpsi_dsigma(energy, s, mu, lam):
    J = s[:, 0] * s[:, 1] * s[:, 2]
    noI = np.stack([s[:, 1] * s[:, 2], s[:, 2] * s[:, 0], s[:, 0] * s[:, 1]], axis=1)
    if energy == FCR:
        return (2.0 * mu)[:, None] * (s - 1.0) + noI * (lam * (J - 1.0))[:, None]
    term2 = lam * (J - (1.0 + mu / lam))
    return s * mu[:, None] + term2[:, None] * noI


def d2psi_dsigma2(energy, s, mu, lam):
    n = s.shape[0]
    J = s[:, 0] * s[:, 1] * s[:, 2]
    noI = np.stack([s[:, 1] * s[:, 2], s[:, 2] * s[:, 0], s[:, 0] * s[:, 1]], axis=1)
    A = np.zeros((n, 3, 3))
    if energy == FCR:
        for i in range(3):
            A[:, i, i] = 2.0 * mu + lam * noI[:, i] * noI[:, i]
        A[:, 0, 1] = A[:, 1, 0] = lam * (s[:, 2] * (J - 1.0) + noI[:, 0] * noI[:, 1])
        A[:, 0, 2] = A[:, 2, 0] = lam * (s[:, 1] * (J - 1.0) + noI[:, 0] * noI[:, 2])
        A[:, 1, 2] = A[:, 2, 1] = lam * (s[:, 0] * (J - 1.0) + noI[:, 2] * noI[:, 1])
    else:
        l2 = lam * (2.0 * J - (1.0 + mu / lam))
        for i in range(3):
            A[:, i, i] = mu + lam * noI[:, i] * noI[:, i]
        A[:, 0, 1] = A[:, 1, 0] = s[:, 2] * l2
        A[:, 0, 2] = A[:, 2, 0] = s[:, 1] * l2
        A[:, 1, 2] = A[:, 2, 1] = s[:, 0] * l2
    return A


def b_left_coef(energy, s, mu, lam):
    J = s[:, 0] * s[:, 1] * s[:, 2]
    o = np.stack([s[:, 2], s[:, 0], s[:, 1]], axis=1)
    if energy == FCR:
        return mu[:, None] - (lam / 2.0)[:, None] * o * (J - 1.0)[:, None]
    term0 = lam * (J - (1.0 + mu / lam))
    return (mu[:, None] - term0[:, None] * o) / 2.0


def elastic_energy_per_elem(energy, mesh: Mesh, s):
    """Psi(sigma)*vol (Energy.cpp:294-423, :842, :900)."""
    return psi(energy, s, mesh.mu, mesh.lam) * mesh.vol


# closed forms without SVD (StableNHEnergy.cpp:91-94,246-249; FixedCoRotEnergy.cpp:87-91,179-182)
def cofactor(F):
    C = np.empty_like(F)
    C[:, 0, 0] = F[:, 1, 1] * F[:, 2, 2] - F[:, 1, 2] * F[:, 2, 1]
    C[:, 0, 1] = F[:, 1, 2] * F[:, 2, 0] - F[:, 1, 0] * F[:, 2, 2]
    C[:, 0, 2] = F[:, 1, 0] * F[:, 2, 1] - F[:, 1, 1] * F[:, 2, 0]
    C[:, 1, 0] = F[:, 0, 2] * F[:, 2, 1] - F[:, 0, 1] * F[:, 2, 2]
    C[:, 1, 1] = F[:, 0, 0] * F[:, 2, 2] - F[:, 0, 2] * F[:, 2, 0]
    C[:, 1, 2] = F[:, 0, 1] * F[:, 2, 0] - F[:, 0, 0] * F[:, 2, 1]
    C[:, 2, 0] = F[:, 0, 1] * F[:, 1, 2] - F[:, 0, 2] * F[:, 1, 1]
    C[:, 2, 1] = F[:, 0, 2] * F[:, 1, 0] - F[:, 0, 0] * F[:, 1, 2]
    C[:, 2, 2] = F[:, 0, 0] * F[:, 1, 1] - F[:, 0, 1] * F[:, 1, 0]
    return C


def first_piola_closed_form(energy, F, mu, lam, R=None):
    J = np.linalg.det(F)
    C = cofactor(F)
    if energy == SNH:
        return mu[:, None, None] * F + (lam * (J - (1.0 + mu / lam)))[:, None, None] * C
    return (2.0 * mu)[:, None, None] * (F - R) + (lam * (J - 1.0))[:, None, None] * C


# ---------------------------------------------------------------------------------------
# a6  P = U diag(dPsi/dsigma) V^T, elemental gradient (Energy.cpp:910-973; IglUtils.cpp:836-870)
# ---------------------------------------------------------------------------------------
def first_piola(energy, U, s, V, mu, lam):
    ph = dpsi_dsigma(energy, s, mu, lam)
    return np.einsum("tia,ta,tja->tij", U, ph, V)


def elem_gradient(mesh: Mesh, P, coef):
    """g_e[3+3a+b] = w * DmInv[a,:].P[b,:], g_e[b] = -sum_a ...; w = coef*vol."""
    w = coef * mesh.vol
    G = np.einsum("taj,tbj->tab", mesh.DmInv, P) * w[:, None, None]     # [t,a,b]
    g = np.zeros((mesh.nT, 12))
    g[:, 3:] = G.reshape(mesh.nT, 9)
    g[:, 0:3] = -(G[:, 0] + G[:, 1] + G[:, 2])
    return g


# ---------------------------------------------------------------------------------------
# a7  vertex gather (Energy.cpp:543-563) + inertia (Optimizer.cpp:1204-1211, 1239-1252)
# ---------------------------------------------------------------------------------------
def gather_gradient(mesh: Mesh, ge, fixed_mask):
    g = np.zeros((mesh.nV, 3))
    # ascending tet order per vertex == a stable scatter in tet order
    for k in range(4):
        pass
    order = np.argsort(np.repeat(np.arange(mesh.nT), 4), kind="stable")
    vidx = mesh.T.reshape(-1)[order]
    vals = ge.reshape(-1, 3)[order]
    np.add.at(g, vidx, vals)
    g[fixed_mask] = 0.0
    return g.reshape(-1)


def incremental_potential(energy, mesh: Mesh, x, xTilde, dt):
    F = deformation_gradient(mesh, x)
    U, s, V = svd_rot(F)
    Eel = (dt * dt) * elastic_energy_per_elem(energy, mesh, s).sum()
    Ein = (((x - xTilde) ** 2).sum(axis=1) * mesh.mass / 2.0).sum()      # ALL vertices
    return Eel + Ein, Eel, (F, U, s, V)


def full_gradient(energy, mesh: Mesh, x, xTilde, dt, fixed_mask, svd=None):
    if svd is None:
        F = deformation_gradient(mesh, x)
        U, s, V = svd_rot(F)
    else:
        F, U, s, V = svd
    P = first_piola(energy, U, s, V, mesh.mu, mesh.lam)
    ge = elem_gradient(mesh, P, dt * dt)
    g_el = gather_gradient(mesh, ge, fixed_mask)
    g = g_el.reshape(-1, 3).copy()
    free = ~fixed_mask
    g[free] += mesh.mass[free, None] * (x[free] - xTilde[free])
    return g.reshape(-1), g_el, ge


# ---------------------------------------------------------------------------------------
# a8  elemental PD-projected Hessian (Energy.cpp:673-777, 1129-1270; IglUtils.hpp:252-309)
# ---------------------------------------------------------------------------------------
def make_pd3(A):
    w, Q = np.linalg.eigh(A)
    bad = w[:, 0] < 0.0
    if bad.any():
        wc = np.maximum(w[bad], 0.0)
        A = A.copy()
        A[bad] = np.einsum("tia,ta,tja->tij", Q[bad], wc, Q[bad])
    return A


def make_pd2(a, b, d):
    """IglUtils.hpp:270-309 on [[a,b],[b,d]] -> (a,b,d)."""
    a, b, d = a.copy(), b.copy(), d.copy()
    b2 = b * b
    D = a * d - b2
    T2 = (a + d) / 2.0
    sq = np.sqrt(np.maximum(T2 * T2 - D, 0.0))
    L2 = T2 - sq
    L1 = T2 + sq
    neg = L2 < 0.0
    zero = neg & (L1 <= 0.0)
    diag = neg & ~zero & (b2 == 0.0)
    rank1 = neg & ~zero & (b2 != 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        L1md = L1 - d
        r = L1md / L1
        na = np.where(rank1, r * L1md, a)
        nb = np.where(rank1, b * r, b)
        nd = np.where(rank1, b2 / L1, d)
    na = np.where(diag, L1, na)
    nb = np.where(diag, 0.0, nb)
    nd = np.where(diag, 0.0, nd)
    na = np.where(zero, 0.0, na)
    nb = np.where(zero, 0.0, nb)
    nd = np.where(zero, 0.0, nd)
    return na, nb, nd


def dP_dF(energy, U, s, V, mu, lam, w, project=True):
    """w * dP/dF as [nT,9,9] with F vectorised row-major (ij = 3i+j)."""
    n = s.shape[0]
    dE = dpsi_dsigma(energy, s, mu, lam)
    A = d2psi_dsigma2(energy, s, mu, lam)
    if project:
        A = make_pd3(A)
    BL = b_left_coef(energy, s, mu, lam)
    M = np.zeros((n, 9, 9))
    idx = [0, 4, 8]
    for i in range(3):
        for j in range(3):
            M[:, idx[i], idx[j]] = A[:, i, j]
    B = []
    for c in range(3):
        cp = (c + 1) % 3
        right = dE[:, c] + dE[:, cp]
        ss = s[:, c] + s[:, cp]
        right = right / (2.0 * np.where(ss < 1e-6, 1e-6, ss))
        left = BL[:, c]
        b00, b01, b11 = left + right, left - right, left + right
        if project:
            b00, b01, b11 = make_pd2(b00, b01, b11)
        B.append((b00, b01, b11))
    M[:, 1, 1], M[:, 1, 3], M[:, 3, 1], M[:, 3, 3] = B[0][0], B[0][1], B[0][1], B[0][2]
    M[:, 5, 5], M[:, 5, 7], M[:, 7, 5], M[:, 7, 7] = B[1][0], B[1][1], B[1][1], B[1][2]
    M[:, 2, 2], M[:, 2, 6], M[:, 6, 2], M[:, 6, 6] = B[2][2], B[2][1], B[2][1], B[2][0]
    M *= np.asarray(w).reshape(-1, 1, 1)
    Q = np.einsum("tia,tjb->tijab", U, V).reshape(n, 9, 9)
    return Q @ M @ np.swapaxes(Q, 1, 2)


def dF_dx(DmInv):
    """9x12 per tet (IglUtils.cpp:821-833)."""
    n = DmInv.shape[0]
    G = np.zeros((n, 9, 12))
    w0 = -(DmInv[:, 0] + DmInv[:, 1] + DmInv[:, 2])       # [t, j]
    for i in range(3):
        for j in range(3):
            G[:, 3 * i + j, i] = w0[:, j]
            for k in range(3):
                G[:, 3 * i + j, 3 * (k + 1) + i] = DmInv[:, k, j]
    return G


def elem_hessians(energy, mesh: Mesh, U, s, V, coef, project=True):
    w = coef * mesh.vol
    dp = dP_dF(energy, U, s, V, mesh.mu, mesh.lam, w, project)
    G = dF_dx(mesh.DmInv)
    return np.swapaxes(G, 1, 2) @ dp @ G


# ---------------------------------------------------------------------------------------
# a11 CSR-upper pattern of a 3x3-block matrix (LinSysSolver.hpp:37-135), 0-based output
# ---------------------------------------------------------------------------------------
def set_pattern(v_neighbor, fixed):
    fixed = set(int(f) for f in fixed)
    ia = [0]
    ja = []
    for v, nbs in enumerate(v_neighbor):
        if v in fixed:
            for c in range(3):
                ja.append(3 * v + c)
                ia.append(ia[-1] + 1)
            continue
        cols = [3 * v, 3 * v + 1, 3 * v + 2]
        for w in nbs:
            if w not in fixed and w > v:
                cols += [3 * w, 3 * w + 1, 3 * w + 2]
        for c in range(3):
            ja += cols[c:]
            ia.append(ia[-1] + len(cols) - c)
    return np.asarray(ia, dtype=np.int32), np.asarray(ja, dtype=np.int32)


def _slot_lookup(ia, ja):
    n = len(ia) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ia))
    keys = rows * (3 * n + 3) + ja
    return keys, (3 * n + 3)


def _add(a, keys, stride, r, c, val):
    k = np.searchsorted(keys, r * stride + c)
    assert keys[k] == r * stride + c, "entry not in pattern"
    a[k] += val


# ---------------------------------------------------------------------------------------
# a9  global matrix fill (DOTTimeStepper.cpp:574-616; IglUtils.hpp:143-220)
# ---------------------------------------------------------------------------------------
def fill_global(mesh: Mesh, He, ia, ja, fixed_mask):
    a = np.zeros(len(ja))
    keys, st = _slot_lookup(ia, ja)
    floc = mesh.v_floc()
    for v in range(mesh.nV):
        if fixed_mask[v]:
            for c in range(3):
                k = np.searchsorted(keys, (3 * v + c) * st + 3 * v + c)
                a[k] = 1.0
            continue
        for (t, k) in floc[v]:
            for kk in range(4):
                u = mesh.T[t, kk]
                if fixed_mask[u]:
                    continue
                for i in range(3):
                    for j in range(3):
                        r, c = 3 * v + i, 3 * u + j
                        if r <= c:
                            _add(a, keys, st, r, c, He[t, 3 * k + i, 3 * kk + j])
        for c in range(3):
            _add(a, keys, st, 3 * v + c, 3 * v + c, mesh.mass[v])
    return a


# ---------------------------------------------------------------------------------------
# a14 domain decomposition set-up from METIS labels
#     (ADMMDDTimeStepper.cpp:155-278, 457-496; Mesh.cpp:855-905; DOTTimeStepper.cpp:38-56)
# ---------------------------------------------------------------------------------------
@dataclass
class Subdomain:
    elems: np.ndarray          # ascending global tet ids
    l2g: np.ndarray            # local vertex -> global (first-touch order)
    g2l: dict
    T_local: np.ndarray
    fixed_local: np.ndarray
    mass_local: np.ndarray
    interface: list            # ascending global ids shared with another subdomain
    ia: np.ndarray = None
    ja: np.ndarray = None


def decompose(mesh: Mesh, epart, fixed):
    k = int(epart.max()) + 1
    fixed = sorted(int(f) for f in fixed)
    subs = []
    count = np.zeros(mesh.nV, dtype=np.int32)
    for s in range(k):
        elems = np.nonzero(epart == s)[0].astype(np.int32)
        g2l, l2g = {}, []
        Tl = np.empty((len(elems), 4), dtype=np.int32)
        for li, t in enumerate(elems):
            for c in range(4):
                gv = int(mesh.T[t, c])
                if gv not in g2l:
                    g2l[gv] = len(l2g)
                    l2g.append(gv)
                Tl[li, c] = g2l[gv]
        l2g = np.asarray(l2g, dtype=np.int32)
        count[l2g] += 1
        ml = np.zeros(len(l2g))
        X = mesh.V_rest
        Tg = mesh.T[elems]
        a, b, c3 = X[Tg[:, 0]] - X[Tg[:, 3]], X[Tg[:, 1]] - X[Tg[:, 3]], X[Tg[:, 2]] - X[Tg[:, 3]]
        v = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c3))) / 6.0
        for c in range(4):
            np.add.at(ml, Tl[:, c], v / 4.0)
        fl = np.asarray(sorted(g2l[f] for f in fixed if f in g2l), dtype=np.int32)
        subs.append(Subdomain(elems, l2g, g2l, Tl, fl, ml * mesh.rho, []))
    dup = count
    nbg = mesh.v_neighbor()
    for s, sd in enumerate(subs):
        sd.interface = sorted(int(g) for g in sd.l2g if dup[g] > 1)
        nb = [set() for _ in range(len(sd.l2g))]
        for t in sd.T_local:
            for i in range(4):
                for j in range(i + 1, 4):
                    nb[t[i]].add(int(t[j]))
                    nb[t[j]].add(int(t[i]))
        for g in sd.interface:
            lv = sd.g2l[g]
            for ng in nbg[g]:
                if ng in sd.g2l:
                    nb[lv].add(sd.g2l[ng])
        sd.ia, sd.ja = set_pattern([sorted(x) for x in nb], sd.fixed_local)
    return subs, dup


# ---------------------------------------------------------------------------------------
# a10 subdomain matrix fill incl. interface completion (DOTTimeStepper.cpp:619-797)
# ---------------------------------------------------------------------------------------
def fill_subdomain(mesh: Mesh, sd: Subdomain, He, fixed_mask):
    a = np.zeros(len(sd.ja))
    keys, st = _slot_lookup(sd.ia, sd.ja)
    nl = len(sd.l2g)
    fl = np.zeros(nl, dtype=bool)
    fl[sd.fixed_local] = True
    floc = [[] for _ in range(nl)]
    for li in range(len(sd.elems)):
        for c in range(4):
            floc[sd.T_local[li, c]].append((li, c))
    for v in range(nl):
        if fl[v]:
            for c in range(3):
                k = np.searchsorted(keys, (3 * v + c) * st + 3 * v + c)
                a[k] = 1.0
        else:
            for (li, k) in floc[v]:
                t = sd.elems[li]
                for kk in range(4):
                    u = sd.T_local[li, kk]
                    if fl[u]:
                        continue
                    for i in range(3):
                        for j in range(3):
                            r, c = 3 * v + i, 3 * u + j
                            if r <= c:
                                _add(a, keys, st, r, c, He[t, 3 * k + i, 3 * kk + j])
        # NOTE (DOTTimeStepper.cpp:677-686): sub-mesh mass is added for every local vertex,
        # fixed ones included (a fixed diagonal therefore holds 1 + m_local).
        for c in range(3):
            _add(a, keys, st, 3 * v + c, 3 * v + c, sd.mass_local[v])
    in_sub = set(int(t) for t in sd.elems)
    inter = set(sd.interface)
    gfloc = mesh.v_floc()
    for g in sd.interface:
        if fixed_mask[g]:
            continue
        lv = sd.g2l[g]
        md = mesh.mass[g] - sd.mass_local[lv]
        for c in range(3):
            _add(a, keys, st, 3 * lv + c, 3 * lv + c, md)
        for (t, k) in gfloc[g]:
            if t in in_sub:
                continue
            for i in range(3):
                for j in range(3):
                    if i <= j:
                        _add(a, keys, st, 3 * lv + i, 3 * lv + j, He[t, 3 * k + i, 3 * k + j])
            for kk in range(4):
                u = int(mesh.T[t, kk])
                if fixed_mask[u] or kk == k or u not in inter:
                    continue
                lu = sd.g2l[u]
                for i in range(3):
                    for j in range(3):
                        r, c = 3 * lv + i, 3 * lu + j
                        if r <= c:
                            _add(a, keys, st, r, c, He[t, 3 * k + i, 3 * kk + j])
    return a


def csr_upper_to_full(ia, ja, a):
    n = len(ia) - 1
    Up = sp.csr_matrix((a, ja, ia), shape=(n, n))
    return (Up + sp.triu(Up, 1).T).tocsc()


def spmv_sym(ia, ja, a, x):
    """cholmod_sdmult with stype=-1 (CHOLMODSolver.cpp:185-208)."""
    return csr_upper_to_full(ia, ja, a) @ x


# ---------------------------------------------------------------------------------------
# a15 scripted Dirichlet motion (AnimScripter.cpp:29-453; IglUtils.cpp:909-927)
# ---------------------------------------------------------------------------------------
def border_verts(V, ratio=0.01):
    lo, hi = V.min(axis=0), V.max(axis=0)
    rng = hi - lo
    left = np.nonzero(V[:, 0] < lo[0] + rng[0] * ratio)[0]
    right = np.nonzero((V[:, 0] > hi[0] - rng[0] * ratio) & ~(V[:, 0] < lo[0] + rng[0] * ratio))[0]
    return [left, right]


def rot_x(angle):
    """Eigen::AngleAxis(angle, UnitX).toRotationMatrix() term by term."""
    c, s = math.cos(angle), math.sin(angle)
    c1 = 1.0 - c
    R = np.zeros((3, 3))
    R[0, 0] = c1 * 1.0 * 1.0 + c
    R[1, 1] = c1 * 0.0 * 0.0 + c
    R[2, 2] = c1 * 0.0 * 0.0 + c
    R[1, 2] = 0.0 - s
    R[2, 1] = 0.0 + s
    return R


class AnimScripter:
    def __init__(self, kind, V, handles):
        self.kind = kind
        self.handles = handles
        self.center = (V.min(axis=0) + V.max(axis=0)) / 2.0   # bbox.colwise().mean()
        sgn = [1.0, -1.0]
        self.ang = {}
        self.vel = {}
        self.changed = False
        self.released = False
        if kind == "rubberBandPull":                                     # AnimScripter.cpp:219-257
            lo, hi = V.min(axis=0), V.max(axis=0)
            ry = hi[1] - lo[1]
            waist, ends = [], []
            for v in range(V.shape[0]):
                y = V[v, 1]
                if y < lo[1] + ry * 0.02:
                    ends.append(v); self.vel[v] = np.array([0.0, -0.2, 0.0])
                elif y > hi[1] - ry * 0.02:
                    ends.append(v); self.vel[v] = np.array([0.0, 0.2, 0.0])
                elif hi[1] - ry * 0.48 > y > lo[1] + ry * 0.48:
                    waist.append(v); self.vel[v] = np.array([-2.5, 0.0, 0.0])
            self.handles = [np.array(waist, dtype=np.int32), np.array(ends, dtype=np.int32)]
            self.ang = {}
            self.turn = (int(waist[0]), V[waist[0], 0] - 5.0, np.inf)
            self.fixed_now = np.sort(np.concatenate(self.handles)).astype(np.int32)
            return
        spec = {"twist": (-0.1 * math.pi, None), "stretch": (None, -0.1), "squash": (None, 0.03),
                "twistnstretch": (-0.1 * math.pi, -0.1), "twistnsns": (-0.4 * math.pi, -1.2),
                "twistnsns_old": (-0.4 * math.pi, -0.9), "stretchnsquash": (None, -0.9), "null": (None, None)}[kind]
        for b, hv in enumerate(handles):
            for v in hv:
                if spec[0] is not None:
                    self.ang[int(v)] = sgn[b] * spec[0]
                if spec[1] is not None:
                    self.vel[int(v)] = np.array([sgn[b] * spec[1], 0.0, 0.0])
        self.turn = None
        if kind in ("twistnsns", "twistnsns_old", "stretchnsquash"):
            v0 = int(handles[0][0])
            lo = V[v0, 0] - (1.2 if kind == "twistnsns" else 0.8)
            self.turn = (v0, lo, V[v0, 0] + 0.4)

    def fixed(self):
        if self.kind == "null":
            return np.array([0], dtype=np.int32)
        if self.kind == "rubberBandPull":
            return self.fixed_now
        return np.sort(np.concatenate(self.handles)).astype(np.int32)

    def step(self, x, dt):
        d = np.zeros_like(x)
        self.changed = False
        if self.kind == "rubberBandPull":                                # AnimScripter.cpp:404-423
            v0, lo, _ = self.turn
            if not self.released and x[v0, 0] <= lo:
                self.released = True
                self.changed = True
                for v in self.vel:
                    self.vel[v] = np.zeros(3)
                self.fixed_now = np.sort(self.handles[1]).astype(np.int32)
            for v in self.vel:
                d[v] += self.vel[v] * dt
            return x + 1.0 * d
        for v, w in self.ang.items():
            R = rot_x(w * dt)
            d[v] = (R @ (x[v] - self.center) + self.center) - x[v]
        flip = False
        if self.turn is not None:
            v0, lo, hi = self.turn
            flip = (x[v0, 0] <= lo) or (x[v0, 0] >= hi)
        for v in self.vel:
            if flip:
                self.vel[v][0] *= -1.0
            d[v] += self.vel[v] * dt
        return x + 1.0 * d


# ---------------------------------------------------------------------------------------
# a13 tolerance (Optimizer.cpp:613-651) - evaluated on data0 whose fixedVert is {0}
# ---------------------------------------------------------------------------------------
def face_areas(V, T):
    """igl::face_areas: column j = area of the face opposite vertex j."""
    out = np.zeros((T.shape[0], 4))
    for j in range(4):
        o = [i for i in range(4) if i != j]
        a, b, c = V[T[:, o[0]]], V[T[:, o[1]]], V[T[:, o[2]]]
        out[:, j] = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    return out


def target_gres(energy, mesh: Mesh, dt, rel_tol=1e-5, n_fixed0=1):
    I = np.eye(3)[None]
    one = np.ones((1, 3))
    H = dP_dF(energy, I, one, I, mesh.mu[:1], mesh.lam[:1], np.ones(1), project=False)[0]
    fa = face_areas(mesh.V_rest, mesh.T)
    ls = np.zeros(mesh.nV)
    for i in range(4):
        np.add.at(ls, mesh.T[:, i], fa[:, i])
    nV = mesh.nV
    return (rel_tol ** 2) * (H ** 2).sum() * (ls ** 2).sum() * (nV - n_fixed0) / nV * dt ** 4


# ---------------------------------------------------------------------------------------
# a12/a13 the DOT time stepper (DOTTimeStepper.cpp:273-504; Optimizer.cpp:327-368, 442-610, 752-881, 1076-1093)
# ---------------------------------------------------------------------------------------
class DOTStepper:
    def __init__(self, mesh: Mesh, energy, epart, anim_kind="twist", dt=0.025, handle_ratio=0.01, rel_tol=1e-5,
                 history=5):
        self.mesh, self.energy, self.dt, self.m = mesh, energy, dt, history
        self.x = mesh.V_rest.copy()
        self.anim = AnimScripter(anim_kind, mesh.V_rest, border_verts(mesh.V_rest, handle_ratio))
        self.fixed = self.anim.fixed()
        self.fixed_mask = np.zeros(mesh.nV, dtype=bool)
        self.fixed_mask[self.fixed] = True
        self.subs, self.dup = decompose(mesh, epart, self.fixed)
        self.gia, self.gja = set_pattern(mesh.v_neighbor(), self.fixed)
        self.gravity = np.array([0.0, -9.80665, 0.0])
        self.vel = np.zeros_like(self.x)
        self.x_n = self.x.copy()
        self.target = target_gres(energy, mesh, dt, rel_tol)
        self.compute_xtilde()
        self.inner_iters = 0
        self.halvings = 0
        self.log = []
        _, _, svd = incremental_potential(energy, mesh, self.x, self.xTilde, dt)
        self.refresh(svd)

    def restart(self, frames_done, x, vel):
        """Continue from the state the reference had after `frames_done` frames (positions + velocity,
        the content of its `status<n>` files, Optimizer.cpp:1096-1162): replays the scripted handle
        motion, sets x^n, x~ and refreshes the Hessian at x like the end of a frame does."""
        dummy = self.mesh.V_rest.copy()
        self.anim = AnimScripter(self.anim.kind, self.mesh.V_rest, self.anim.handles)
        for _ in range(frames_done):
            dummy = self.anim.step(dummy, self.dt)
        self.x = np.array(x, dtype=float, copy=True)
        self.x_n = self.x.copy()
        self.vel = np.array(vel, dtype=float).reshape(-1, 3).copy()
        self.compute_xtilde()
        _, _, svd = incremental_potential(self.energy, self.mesh, self.x, self.xTilde, self.dt)
        self.refresh(svd)
        self.log = []

    def compute_xtilde(self):                                           # Optimizer.cpp:585-610
        self.xTilde = self.x_n + self.vel * self.dt + self.gravity * self.dt ** 2
        self.xTilde[self.fixed_mask] = self.x_n[self.fixed_mask]

    def refresh(self, svd):                                              # DOTTimeStepper.cpp:349-380
        F, U, s, V = svd
        self.He = elem_hessians(self.energy, self.mesh, U, s, V, self.dt ** 2, True)
        self.ga = fill_global(self.mesh, self.He, self.gia, self.gja, self.fixed_mask)
        self.sa = [fill_subdomain(self.mesh, sd, self.He, self.fixed_mask) for sd in self.subs]
        self.lu = [spla.splu(csr_upper_to_full(sd.ia, sd.ja, a)) for sd, a in zip(self.subs, self.sa)]

    def precondition(self, q):                                           # DOTTimeStepper.cpp:406-450
        p = np.zeros_like(q)
        q3, p3 = q.reshape(-1, 3), p.reshape(-1, 3)
        for sd, lu in zip(self.subs, self.lu):
            ps = lu.solve(q3[sd.l2g].reshape(-1))
            p3[sd.l2g] += ps.reshape(-1, 3)
        m = self.dup > 1
        p3[m] /= self.dup[m, None]
        return p

    def energy_at(self, x):
        return incremental_potential(self.energy, self.mesh, x, self.xTilde, self.dt)

    def init_step(self, p, g):                                           # Optimizer::initStepSize, DOT branch (Optimizer.cpp:1076-1093)
        Hp = spmv_sym(self.gia, self.gja, self.ga, p)
        return max(0.1, min(1.0, -(p @ g) / (p @ Hp)))

    def step_frame(self):
        mesh, dt = self.mesh, self.dt
        self.x = self.anim.step(self.x, dt)                              # Optimizer.cpp:334
        dx, dg, dgdx = [], [], []
        d = self.vel * dt + self.gravity * dt * dt                       # initX(2), Optimizer.cpp:472-493
        d[self.fixed_mask] = 0.0
        self.x = self.x + 1.0 * d
        E, _, svd = self.energy_at(self.x)
        g, _, _ = full_gradient(self.energy, mesh, self.x, self.xTilde, dt, self.fixed_mask, svd)
        self.log.append((0.0, E, float(g @ g)))
        it = 0
        while True:
            q = -g
            ksi = [0.0] * len(dx)
            for i in range(len(dx) - 1, -1, -1):
                ksi[i] = dx[i] @ q / dgdx[i]
                q = q - ksi[i] * dg[i]
            p = self.precondition(q)
            for i in range(len(dx)):
                p = p + dx[i] * (ksi[i] - dg[i] @ p / dgdx[i])
            alpha = self.init_step(p, g)
            x0 = self.x
            while True:
                xt = x0 + alpha * p.reshape(-1, 3)
                Et, _, svd = self.energy_at(xt)
                if not (Et > E and alpha > 0.0):
                    break
                alpha /= 2.0
                self.halvings += 1
            self.x, E = xt, Et
            g_old = g
            g, _, _ = full_gradient(self.energy, mesh, self.x, self.xTilde, dt, self.fixed_mask, svd)
            s_new, y_new = alpha * p, g - g_old
            ys = y_new @ s_new
            if ys > 0.0:
                dx.append(s_new), dg.append(y_new), dgdx.append(ys)
                if len(dx) > self.m:
                    dx.pop(0), dg.pop(0), dgdx.pop(0)
            self.inner_iters += 1
            it += 1
            gg = float(g @ g)
            self.log.append((alpha, E, gg))
            if gg <= self.target or it >= 10000:
                break
        self.refresh(svd)
        self.vel = (self.x - self.x_n) / dt                              # Optimizer.cpp:354-361
        self.x_n = self.x.copy()
        self.compute_xtilde()
        return it


# ---------------------------------------------------------------------------------------
# f4  the other L-BFGS initialisers that share the kernels (LBFGSTimeStepper.cpp:108-265 precompute, 286-335 per-frame refresh,
#     339-420 solve_oneStep): same two-loop recursion, history 5, line search from step 1 (initStepSize: only DOT uses p.Hp)
# ---------------------------------------------------------------------------------------
class LBFGSStepper(DOTStepper):
    """d0 = "H": initial inverse Hessian = the global PD-projected Hessian (+ mass) at the start of the time step, factorised once
    per frame (D0T_H).  d0 = "JH": block Jacobi of that matrix over a node partition (D0T_JH; `npart` = METIS<3>::partMesh_nodes
    labels, node lists ascending): every block is solved on its own, results are scattered without averaging."""

    def __init__(self, mesh, energy, d0="H", npart=None, anim_kind="twist", dt=0.025, handle_ratio=0.01, rel_tol=1e-5, history=5):
        self.d0, self.npart = d0, None if npart is None else np.asarray(npart)
        super().__init__(mesh, energy, np.zeros(mesh.nT, dtype=np.int64), anim_kind, dt, handle_ratio, rel_tol, history)

    def refresh(self, svd):
        F, U, s, V = svd
        self.He = elem_hessians(self.energy, self.mesh, U, s, V, self.dt ** 2, True)
        self.ga = fill_global(self.mesh, self.He, self.gia, self.gja, self.fixed_mask)      # Optimizer::computePrecondMtr
        A = csr_upper_to_full(self.gia, self.gja, self.ga).tocsc()
        if self.d0 == "H":
            self.blocks = [(np.arange(3 * self.mesh.nV), spla.splu(A))]
        else:
            self.blocks = []
            for b in range(int(self.npart.max()) + 1):
                nodes = np.nonzero(self.npart == b)[0]
                dof = (3 * nodes[:, None] + np.arange(3)[None, :]).ravel()
                self.blocks.append((dof, spla.splu(A[dof][:, dof].tocsc())))

    def precondition(self, q):
        p = np.zeros_like(q)
        for dof, lu in self.blocks:
            p[dof] = lu.solve(q[dof])
        return p

    def init_step(self, p, g):
        return 1.0


# ---------------------------------------------------------------------------------------
# f1  Projected Newton: Optimizer::fullyImplicit / solve_oneStep (Optimizer.cpp:654-749), the reference's
#     `timeStepper Newton` ("1 subdomain" in BASELINE.json)
# ---------------------------------------------------------------------------------------
class NewtonStepper:
    """Each iteration: PD-projected Hessian of the incremental potential at the current iterate (computePrecondMtr,
    Optimizer.cpp:1257-1308), factorise, p = -H^-1 g, line search from step 1 (initStepSize, :1088) halving while the energy
    increases (:752-833), new gradient.  Stops when |g|^2 <= targetGRes."""

    def __init__(self, mesh: Mesh, energy, anim_kind="twist", dt=0.025, handle_ratio=0.01, rel_tol=1e-5):
        self.mesh, self.energy, self.dt = mesh, energy, dt
        self.x = mesh.V_rest.copy()
        self.anim = AnimScripter(anim_kind, mesh.V_rest, border_verts(mesh.V_rest, handle_ratio))
        self.fixed = self.anim.fixed()
        self.fixed_mask = np.zeros(mesh.nV, dtype=bool)
        self.fixed_mask[self.fixed] = True
        self.gia, self.gja = set_pattern(mesh.v_neighbor(), self.fixed)
        self.gravity = np.array([0.0, -9.80665, 0.0])
        self.vel = np.zeros_like(self.x)
        self.x_n = self.x.copy()
        self.target = target_gres(energy, mesh, dt, rel_tol)
        self.compute_xtilde()
        self.inner_iters = 0
        self.halvings = 0
        self.log = []

    compute_xtilde = DOTStepper.compute_xtilde
    energy_at = DOTStepper.energy_at

    def restart(self, frames_done, x, vel):
        dummy = self.mesh.V_rest.copy()
        self.anim = AnimScripter(self.anim.kind, self.mesh.V_rest, self.anim.handles)
        for _ in range(frames_done):
            dummy = self.anim.step(dummy, self.dt)
        self.x = np.array(x, dtype=float, copy=True)
        self.x_n = self.x.copy()
        self.vel = np.array(vel, dtype=float).reshape(-1, 3).copy()
        self.compute_xtilde()
        self.log = []

    def hessian(self, svd):
        F, U, s, V = svd
        He = elem_hessians(self.energy, self.mesh, U, s, V, self.dt ** 2, True)
        return fill_global(self.mesh, He, self.gia, self.gja, self.fixed_mask)

    def step_frame(self):
        mesh, dt = self.mesh, self.dt
        self.x = self.anim.step(self.x, dt)
        d = self.vel * dt + self.gravity * dt * dt                       # initX(2)
        d[self.fixed_mask] = 0.0
        self.x = self.x + 1.0 * d
        E, _, svd = self.energy_at(self.x)
        g, _, _ = full_gradient(self.energy, mesh, self.x, self.xTilde, dt, self.fixed_mask, svd)
        self.log.append((0.0, E, float(g @ g)))
        it = 0
        while True:
            a = self.hessian(svd)
            p = spla.splu(csr_upper_to_full(self.gia, self.gja, a)).solve(-g)
            alpha = 1.0
            x0 = self.x
            while True:
                xt = x0 + alpha * p.reshape(-1, 3)
                Et, _, svd = self.energy_at(xt)
                if not (Et > E and alpha > 0.0):
                    break
                alpha /= 2.0
                self.halvings += 1
            self.x, E = xt, Et
            g, _, _ = full_gradient(self.energy, mesh, self.x, self.xTilde, dt, self.fixed_mask, svd)
            self.inner_iters += 1
            it += 1
            gg = float(g @ g)
            self.log.append((alpha, E, gg))
            if gg <= self.target or it >= 10000:
                break
        self.vel = (self.x - self.x_n) / dt
        self.x_n = self.x.copy()
        self.compute_xtilde()
        return it
