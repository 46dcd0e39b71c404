"""ctypes binding of libdotgpu (include/dotgpu.h).  Host buffers are numpy arrays (C-contiguous)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ENERGY_FCR, ENERGY_SNH = 0, 1
ANIM_KINDS = {"null": 0, "stretch": 1, "squash": 2, "stretchnsquash": 3, "twist": 4, "twistnstretch": 5, "twistnsns": 6,
              "twistnsns_old": 7, "rubberBandPull": 8}
_ENERGY = {"FCR": ENERGY_FCR, "SNH": ENERGY_SNH, 0: 0, 1: 1}


class DotGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libdotgpu error %d: %s" % (code, msg))
        self.code = code


def lib_path() -> str:
    return os.environ.get("DOTGPU_LIB") or os.path.join(_HERE, "libdotgpu.so")


class SolverInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("nsuper", C.c_int32), ("nlevels", C.c_int32), ("max_front", C.c_int32), ("max_nscol", C.c_int32),
                ("nnz_a", C.c_int64), ("nnz_l", C.c_int64), ("flops", C.c_double), ("device_bytes", C.c_int64)]


class StepperConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("energy_type", C.c_int32), ("num_subdomains", C.c_int32), ("history", C.c_int32),
                ("dt", C.c_double), ("gravity", C.c_double * 3), ("rel_tol", C.c_double), ("YM", C.c_double), ("PR", C.c_double),
                ("rho", C.c_double), ("max_iters", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("target_fixed_count", C.c_int32), ("flags", C.c_int32), ("node_part", C.c_void_p)]


class FrameStats(C.Structure):
    _fields_ = [("iters", C.c_int32), ("halvings", C.c_int32), ("energy_evals", C.c_int32), ("converged", C.c_int32),
                ("E", C.c_double), ("grad_sqnorm", C.c_double), ("target", C.c_double), ("ms_total", C.c_double),
                ("ms_solve", C.c_double), ("ms_refresh", C.c_double), ("ms_precond", C.c_double), ("precond_calls", C.c_int32),
                ("line_search_failed", C.c_int32)]


def lib():
    """Loads libdotgpu.so; raises if it has not been built (python -m dot_b200.build)."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise ImportError("libdotgpu.so is missing: run `python -m dot_b200.build` (needs nvcc); there is no CPU fallback")
        L = C.CDLL(p, mode=C.RTLD_GLOBAL)
        L.dotgpu_last_error.restype = C.c_char_p
        L.dotgpu_dd_nnz.restype = C.c_int64
        L.dotgpu_stepper_launch_count.restype = C.c_int64
        for f in ("dotgpu_energy_destroy", "dotgpu_solver_destroy", "dotgpu_dd_destroy", "dotgpu_anim_destroy",
                  "dotgpu_stepper_destroy", "dotgpu_stepper_default_config"):
            getattr(L, f).restype = None
        _LIB = L
    return _LIB


def _chk(rc):
    if rc != 0:
        raise DotGpuError(rc, lib().dotgpu_last_error().decode())


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count() -> int:
    return int(lib().dotgpu_device_count())


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _chk(lib().dotgpu_nccl_unique_id(buf))
    return buf.raw


def owned_subdomains(k, rank, world):
    """Subdomain ids rank `rank` of `world` factors and solves (the multi-GPU sharding of the path)."""
    n = lib().dotgpu_owned_subdomains(int(k), int(rank), int(world), None)
    if n < 0:
        raise DotGpuError(n, "bad rank/world")
    out = np.empty(n, dtype=np.int32)
    lib().dotgpu_owned_subdomains(int(k), int(rank), int(world), _p(out))
    return out


def balanced_owner(weights, world):
    """owner[s] of every subdomain: longest-processing-time greedy on the weights (nnz(L_s) in the stepper)."""
    w = _f64(weights)
    out = np.empty(w.shape[0], dtype=np.int32)
    _chk(lib().dotgpu_balanced_owner(w.shape[0], _p(w), int(world), _p(out)))
    return out


def partition(nV, tets, k):
    """METIS<3>::partMesh with the reference's vendored METIS and option vector: element labels [nT] int32 (bit-exact)."""
    T = _i32(tets)
    ep = np.empty(T.shape[0], dtype=np.int32)
    _chk(lib().dotgpu_partition(int(nV), T.shape[0], _p(T), int(k), _p(ep)))
    return ep


def partition_nodes(nV, tets, k):
    """METIS<3>::partMesh_nodes (METIS_PartMeshNodal, the reference's vendored METIS and option vector): node labels [nV] int32."""
    T = _i32(tets)
    npart = np.empty(int(nV), dtype=np.int32)
    _chk(lib().dotgpu_partition_nodes(int(nV), T.shape[0], _p(T), int(k), _p(npart)))
    return npart


def mesh_features(V_rest, tets, YM=1e5, PR=0.4, rho=1000.0):
    """Mesh::computeFeatures: returns DmInv [nT,3,3], vol [nT], mass [nV], mu [nT], lam [nT]."""
    V = _f64(V_rest)
    T = _i32(tets)
    nV, nT = V.shape[0], T.shape[0]
    Dm, vol, mass = np.empty((nT, 3, 3)), np.empty(nT), np.empty(nV)
    mu, lam = np.empty(nT), np.empty(nT)
    _chk(lib().dotgpu_mesh_features(nV, nT, _p(V), _p(T), C.c_double(YM), C.c_double(PR), C.c_double(rho), _p(Dm), _p(vol), _p(mass),
                                    _p(mu), _p(lam)))
    return Dm, vol, mass, mu, lam


class Energy:
    """Sits where the reference's Energy<3> subclasses sit (FixedCoRotEnergy / StableNHEnergy)."""

    def __init__(self, energy, tets, DmInv, vol, mu, lam, nV, fixed_mask=None, device=0):
        self.T = _i32(tets)
        self.nT, self.nV = self.T.shape[0], int(nV)
        self.h = C.c_void_p()
        fm = _u8(fixed_mask) if fixed_mask is not None else None
        _chk(lib().dotgpu_energy_create(C.byref(self.h), device, _ENERGY[energy], self.nV, self.nT, _p(self.T), _p(_f64(DmInv)),
                                        _p(_f64(vol)), _p(_f64(mu)), _p(_f64(lam)), _p(fm)))

    def close(self):
        if self.h:
            lib().dotgpu_energy_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def set_fixed(self, mask):
        _chk(lib().dotgpu_energy_set_fixed(self.h, _p(_u8(mask))))

    def compute_energy_val(self, x, coef=1.0):
        E = C.c_double()
        _chk(lib().dotgpu_energy_value(self.h, _p(_f64(x)), C.c_double(coef), C.byref(E)))
        return E.value

    def energy_per_elem(self, x):
        out = np.empty(self.nT)
        _chk(lib().dotgpu_energy_per_elem(self.h, _p(_f64(x)), _p(out)))
        return out

    def compute_gradient(self, x, coef=1.0):
        g = np.empty(3 * self.nV)
        _chk(lib().dotgpu_energy_gradient(self.h, _p(_f64(x)), C.c_double(coef), _p(g)))
        return g

    def svd(self, x):
        F, U, V = np.empty((self.nT, 3, 3)), np.empty((self.nT, 3, 3)), np.empty((self.nT, 3, 3))
        S = np.empty((self.nT, 3))
        _chk(lib().dotgpu_energy_svd(self.h, _p(_f64(x)), _p(F), _p(U), _p(S), _p(V)))
        return F, U, S, V

    def compute_elem_hessians(self, x, coef=1.0, project_spd=True):
        He = np.empty((self.nT, 12, 12))
        vi = np.empty((self.nT, 4), dtype=np.int32)
        _chk(lib().dotgpu_energy_elem_hessians(self.h, _p(_f64(x)), C.c_double(coef), int(project_spd), _p(He), _p(vi)))
        return He, vi


class Solver:
    """Sits where CHOLMODSolver sits: set_pattern+analyze (ctor), set_values, factorize, solve, multiply.
    device=-1 gives a symbolic-only handle (ordering + supernodal structure, no GPU needed)."""

    def __init__(self, ia, ja, device=0):
        self.ia, self.ja = _i32(ia), _i32(ja)
        self.n = self.ia.shape[0] - 1
        self.h = C.c_void_p()
        _chk(lib().dotgpu_solver_create(C.byref(self.h), device, self.n, _p(self.ia), _p(self.ja)))

    def close(self):
        if self.h:
            lib().dotgpu_solver_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def set_values(self, a):
        a = _f64(a)
        assert a.shape[0] == self.ja.shape[0]
        _chk(lib().dotgpu_solver_set_values(self.h, _p(a)))

    def factorize(self):
        _chk(lib().dotgpu_solver_factorize(self.h))

    def solve(self, rhs):
        x = np.empty(self.n)
        _chk(lib().dotgpu_solver_solve(self.h, _p(_f64(rhs)), _p(x)))
        return x

    def multiply(self, x):
        y = np.empty(self.n)
        _chk(lib().dotgpu_solver_multiply(self.h, _p(_f64(x)), _p(y)))
        return y

    def info(self):
        i = SolverInfo()
        _chk(lib().dotgpu_solver_get_info(self.h, C.byref(i)))
        return i

    def symbolic(self):
        i = self.info()
        perm = np.empty(i.n, dtype=np.int32)
        sp = np.empty(i.nsuper + 1, dtype=np.int32)
        rp = np.empty(i.nsuper + 1, dtype=np.int64)
        parent = np.empty(i.nsuper, dtype=np.int32)
        level = np.empty(i.nsuper, dtype=np.int32)
        _chk(lib().dotgpu_solver_get_symbolic(self.h, _p(perm), _p(sp), _p(rp), None, _p(parent), _p(level)))
        rows = np.empty(int(rp[-1]), dtype=np.int32)
        _chk(lib().dotgpu_solver_get_symbolic(self.h, None, None, None, _p(rows), None, None))
        return dict(perm=perm, super_ptr=sp, row_ptr=rp, rows=rows, parent=parent, level=level)


class DD:
    """Domain decomposition from element labels (ADMMDDTimeStepper ctor + precompute, host side)."""

    def __init__(self, nV, tets, epart, k, fixed_mask=None, _handle=None):
        self.k = int(k)
        if _handle is not None:
            self.h = _handle
            return
        T = _i32(tets)
        self.h = C.c_void_p()
        fm = _u8(fixed_mask) if fixed_mask is not None else None
        _chk(lib().dotgpu_dd_create(C.byref(self.h), int(nV), T.shape[0], _p(T), _p(_i32(epart)), self.k, _p(fm)))
        self.nV = int(nV)

    def close(self):
        if self.h:
            lib().dotgpu_dd_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def l2g(self, s):
        n = lib().dotgpu_dd_num_local_verts(self.h, s)
        out = np.empty(n, dtype=np.int32)
        _chk(lib().dotgpu_dd_get_l2g(self.h, s, _p(out)))
        return out

    def fixed_local(self, s):
        n = lib().dotgpu_dd_num_local_verts(self.h, s)
        out = np.empty(n, dtype=np.int32)
        cnt = C.c_int()
        _chk(lib().dotgpu_dd_get_fixed_local(self.h, s, _p(out), C.byref(cnt)))
        return out[:cnt.value].copy()

    def pattern(self, s=-1):
        nnz = lib().dotgpu_dd_nnz(self.h, s)
        n = 3 * (self.nV if s < 0 else lib().dotgpu_dd_num_local_verts(self.h, s))
        ia = np.empty(n + 1, dtype=np.int32)
        ja = np.empty(nnz, dtype=np.int32)
        _chk(lib().dotgpu_dd_get_pattern(self.h, s, _p(ia), _p(ja)))
        return ia, ja

    def dup(self):
        out = np.empty(self.nV, dtype=np.int32)
        _chk(lib().dotgpu_dd_get_dup(self.h, _p(out)))
        return out


class Anim:
    """AnimScripter<3>: handle detection + per-frame scripted Dirichlet motion (host)."""

    def __init__(self, kind, V_rest, handle_ratio=0.01):
        V = _f64(V_rest)
        self.nV = V.shape[0]
        self.h = C.c_void_p()
        _chk(lib().dotgpu_anim_create(C.byref(self.h), ANIM_KINDS[kind] if isinstance(kind, str) else int(kind), self.nV, _p(V),
                                      C.c_double(handle_ratio)))

    def close(self):
        if self.h:
            lib().dotgpu_anim_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def fixed_mask(self):
        m = np.empty(self.nV, dtype=np.uint8)
        _chk(lib().dotgpu_anim_fixed_mask(self.h, _p(m)))
        return m

    def step(self, x, dt):
        """stepAnimScript: moves the handle rows of x in place (x: [nV,3] float64 C-contiguous)."""
        assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
        flag = C.c_int(0)
        _chk(lib().dotgpu_anim_step_ex(self.h, _p(x), C.c_double(dt), C.byref(flag)))
        self.changed = bool(flag.value)     # the Dirichlet set changed in this step (rubberBandPull): call Stepper.set_fixed(anim.fixed_mask(), x)
        return x


class Stepper:
    """Device-resident time stepper: DOT (DOTTimeStepper) or, with newton=True and a single subdomain, Projected Newton
    (Optimizer::solve_oneStep, the reference's `timeStepper Newton`)."""

    def __init__(self, V_rest, tets, epart, fixed_mask, energy="SNH", k=None, dt=0.025, device=0, rel_tol=1e-5, YM=1e5, PR=0.4,
                 rho=1000.0, history=5, rank=0, world=1, nccl_id: bytes | None = None, max_iters=10000, newton=False, gravity=(0.0, -9.80665, 0.0),
                 method="DOT", node_part=None):
        """method: "DOT" (default), "Newton" (= newton=True), "LBFGSH" (global Hessian initialiser, k = 1), "LBFGSJH" (block Jacobi over
        the node partition `node_part`, k blocks) - the reference's `timeStepper` names."""
        V = _f64(V_rest)
        T = _i32(tets)
        newton = newton or method == "Newton"
        if method in ("LBFGSH", "LBFGSJH") or epart is None:
            epart = np.zeros(T.shape[0], dtype=np.int32)
        if method == "LBFGSH":
            k = 1
        ep = _i32(epart)
        self.nV, self.nT = V.shape[0], T.shape[0]
        cfg = StepperConfig()
        lib().dotgpu_stepper_default_config(C.byref(cfg))
        cfg.device = device
        cfg.energy_type = _ENERGY[energy]
        cfg.num_subdomains = int(k if k is not None else ep.max() + 1)
        cfg.history = history
        cfg.dt = dt
        cfg.rel_tol = rel_tol
        cfg.YM, cfg.PR, cfg.rho = YM, PR, rho
        cfg.max_iters = max_iters
        cfg.rank, cfg.world = rank, world
        for i in range(3):
            cfg.gravity[i] = float(gravity[i])
        if newton:
            cfg.flags |= 2  # DOTGPU_FLAG_NEWTON
        if method == "LBFGSH":
            cfg.flags |= 4  # DOTGPU_FLAG_LBFGS_H
        self._npart = None
        if method == "LBFGSJH":
            cfg.flags |= 8  # DOTGPU_FLAG_LBFGS_JH
            self._npart = _i32(node_part)
            assert self._npart.shape[0] == self.nV
            cfg.node_part = self._npart.ctypes.data_as(C.c_void_p)
            cfg.num_subdomains = int(k if k is not None else self._npart.max() + 1)
        self._id = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        cfg.nccl_unique_id = C.cast(self._id, C.c_void_p) if self._id is not None else None
        self.cfg = cfg
        self.h = C.c_void_p()
        _chk(lib().dotgpu_stepper_create(C.byref(self.h), C.byref(cfg), self.nV, self.nT, _p(V), _p(T), _p(ep), _p(_u8(fixed_mask))))

    def set_rel_tol(self, rel_tol):
        """Optimizer::setRelGL2Tol: tolerance of the following time steps."""
        _chk(lib().dotgpu_stepper_set_rel_tol(self.h, C.c_double(rel_tol)))

    def close(self):
        if self.h:
            lib().dotgpu_stepper_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def frame(self, x):
        """x [nV,3]: positions after the scripted Dirichlet move; overwritten with the converged positions."""
        assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"] and x.size == 3 * self.nV
        st = FrameStats()
        _chk(lib().dotgpu_stepper_frame(self.h, _p(x), C.byref(st)))
        return st

    def frame_resident(self, fixed_idx, fixed_pos):
        """Positions stay on the device; only the scripted Dirichlet targets (ids + positions) are passed."""
        idx = _i32(fixed_idx)
        pos = _f64(fixed_pos)
        st = FrameStats()
        _chk(lib().dotgpu_stepper_frame_resident(self.h, _p(idx), _p(pos), idx.shape[0], C.byref(st)))
        return st

    def set_state(self, x, velocity=None):
        _chk(lib().dotgpu_stepper_set_state(self.h, _p(_f64(x)), _p(_f64(velocity)) if velocity is not None else None))

    def set_fixed(self, fixed_mask, x_eval=None):
        """updatePrecondMtrAndFactorize: new Dirichlet set -> re-analysis + refactorisation at x_eval (default: the resident x^n)."""
        _chk(lib().dotgpu_stepper_set_fixed(self.h, _p(_u8(fixed_mask)), _p(_f64(x_eval)) if x_eval is not None else None))

    def get_state(self):
        x, v, xt = np.empty((self.nV, 3)), np.empty(3 * self.nV), np.empty((self.nV, 3))
        _chk(lib().dotgpu_stepper_get_state(self.h, _p(x), _p(v), _p(xt)))
        return x, v, xt

    def iter_log(self):
        buf = np.empty((20000, 3))
        n = lib().dotgpu_stepper_get_iter_log(self.h, _p(buf), buf.shape[0])
        return buf[:n].copy()

    def matrix(self, sub=-1):
        dd = self.dd()
        nnz = lib().dotgpu_dd_nnz(dd.h, sub)
        a = np.empty(nnz)
        _chk(lib().dotgpu_stepper_get_matrix(self.h, sub, _p(a)))
        return a

    def dd(self):
        if not hasattr(self, "_dd"):
            h = C.c_void_p()
            _chk(lib().dotgpu_stepper_get_dd(self.h, C.byref(h)))
            self._dd = DD(self.nV, None, None, self.cfg.num_subdomains, _handle=h)
            self._dd.nV = self.nV
        return self._dd

    def precondition(self, q):
        p = np.empty(3 * self.nV)
        _chk(lib().dotgpu_stepper_precondition(self.h, _p(_f64(q)), _p(p)))
        return p

    def eval(self, x, want_gradient=True):
        E = C.c_double()
        g = np.empty(3 * self.nV) if want_gradient else None
        _chk(lib().dotgpu_stepper_eval(self.h, _p(_f64(x)), C.byref(E), _p(g)))
        return E.value, g

    @property
    def target(self):
        t = C.c_double()
        _chk(lib().dotgpu_stepper_get_target(self.h, C.byref(t)))
        return t.value

    def time_kernels(self, which, reps=20):
        ms = C.c_double()
        _chk(lib().dotgpu_stepper_time_kernels(self.h, int(which), int(reps), C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib().dotgpu_stepper_launch_count(self.h))

    def solve_trace(self):
        """[ctas, 96, 8] uint64 %globaltimer stamps of the last preconditioner application (needs DOTGPU_SOLVE_TRACE=1)."""
        lib().dotgpu_stepper_get_solve_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        nw = lib().dotgpu_stepper_get_solve_trace(self.h, None, 0)
        t = np.zeros(max(nw, 0), dtype=np.uint64)
        if nw > 0:
            lib().dotgpu_stepper_get_solve_trace(self.h, _p(t), nw)
        return t.reshape(-1, 96, 8)

    def fill_stats(self):
        """(stored matrix values, 3x3 blocks, elemental 3x3 blocks gathered) of this rank's matrix fill."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _chk(lib().dotgpu_stepper_get_fill_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def owned(self):
        """Subdomain ids this rank factors and solves."""
        n = lib().dotgpu_stepper_get_owned(self.h, None)
        out = np.empty(max(n, 0), dtype=np.int32)
        lib().dotgpu_stepper_get_owned(self.h, _p(out))
        return [int(v) for v in out]

    def solver_info(self, sub):
        i = SolverInfo()
        _chk(lib().dotgpu_stepper_get_solver_info(self.h, sub, C.byref(i)))
        return i
