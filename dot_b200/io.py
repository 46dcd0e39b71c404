"""File formats of the reference, so that existing scripts / meshes / restart files work with the GPU path.

* `parse_script`  - the key/value script format of `Config::loadFromFile` (reference src/Config.cpp:43-200): the keys the
  shipped `input/*.txt` scripts use.
* `read_msh`      - the `.msh` dialect of `IglUtils::readTetMesh` (src/Utils/IglUtils.cpp:680-749); `meshgen.write_msh` writes it.
* `write_status` / `read_status` - restart files of `Optimizer::saveStatus` / the restart branch of the `Optimizer` constructor
  (src/TimeStepper/Optimizer.cpp:1096-1130, 126-177): `timestep`, `position`, `velocity`, `dx_Elastic` sections, `%le` numbers.
* `IterStatsWriter` - `iterStats.txt` rows as `Optimizer::fullyImplicit` / `DOTTimeStepper::fullyImplicit` print them
  (Optimizer.cpp:666-684): `<frame> <alpha> <E> <|g|^2> 0`, first row of a frame has alpha 0.
* `write_label_obj` - `label.obj` of the domain decomposition: one `v <label> 0 0` line per surface triangle
  (src/TimeStepper/ADMMDDTimeStepper.cpp:375-395).

Host-side, pure numpy; nothing here touches the GPU.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

ENERGIES = {"FCR": "FCR", "SNH": "SNH"}
# `script` names -> AnimScripter kinds known to libdotgpu (AnimScripter.cpp:29-76)
ANIM_SCRIPTS = {"null": "null", "stretch": "stretch", "squash": "squash", "stretchnsquash": "stretchnsquash", "twist": "twist",
                "twistnstretch": "twistnstretch", "twistnsns": "twistnsns", "twistnsns_old": "twistnsns_old", "rubberBandPull": "rubberBandPull"}


@dataclass
class Script:
    energy: str = "FCR"
    time_stepper: str = "Newton"      # "Newton" (Projected Newton), "DOT", "LBFGSH" or "LBFGSJH"
    partitions: int = 4               # `timeStepper DOT k`; k < 2 is rewritten to 4 like Config.cpp:76-80
    block_size: int = -1              # `timeStepper DOT -1 <blockSize>`: k = nV / blockSize + 1 (main.cpp:792-798)
    size: float = 1.0                 # longest bounding-box edge after loading (main.cpp:709)
    duration: float = 10.0
    dt: float = 0.025
    rho: float = 1.0                  # Config::Config defaults (Config.cpp:33-37): rho 1, YM 100, PR 0.4 - the shipped scripts set their own
    YM: float = 100.0
    PR: float = 0.4
    with_gravity: bool = True
    script: str = "null"
    shape: str = "input"
    input_shape_path: str = ""
    tol: list = field(default_factory=list)   # relative tolerances per frame; the last one repeats (main.cpp:108-117)
    warm_start: int = 2
    rot_axis: tuple = (0.0, 0.0, 0.0)
    rot_deg: float = 0.0
    handle_ratio: float = 0.01
    restart: str | None = None
    unknown: list = field(default_factory=list)

    def rel_tol(self, frame: int) -> float:
        """Tolerance of time step `frame` (0-based): config.tol[frame], the last entry repeating, default 1e-5."""
        if not self.tol:
            return 1e-5
        return float(self.tol[frame] if frame < len(self.tol) else self.tol[-1])

    def num_frames(self) -> int:
        return int(math.ceil(self.duration / self.dt))


def parse_script(path: str) -> Script:
    s = Script()
    with open(path) as f:
        lines = f.read().splitlines()
    i = 0
    while i < len(lines):
        tok = lines[i].split()
        i += 1
        if not tok:
            continue
        key, args = tok[0], tok[1:]
        if key == "energy":
            if args[0] not in ENERGIES:
                raise ValueError("energy %r is outside the GPU path (FCR and SNH are supported)" % args[0])
            s.energy = args[0]
        elif key == "timeStepper":
            s.time_stepper = args[0]
            if args[0] in ("DOT", "LBFGSJH"):
                k = int(args[1]) if len(args) > 1 else 4
                if k < 0:
                    s.block_size = int(args[2])
                elif k < 2:
                    k = 4
                s.partitions = k
            elif args[0] not in ("Newton", "LBFGSH"):
                raise ValueError("timeStepper %r is outside the GPU path (DOT, Newton, LBFGSH and LBFGSJH are supported)" % args[0])
        elif key == "size":
            s.size = float(args[0])
        elif key == "time":
            s.duration, s.dt = float(args[0]), float(args[1])
        elif key == "density":
            s.rho = float(args[0])
        elif key == "stiffness":
            s.YM, s.PR = float(args[0]), float(args[1])
        elif key == "turnOffGravity":
            s.with_gravity = False
        elif key == "script":
            if args[0] not in ANIM_SCRIPTS:
                raise ValueError("animation script %r is not implemented" % args[0])
            s.script = args[0]
        elif key == "shape":
            s.shape = args[0]
            if args[0] == "input":
                s.input_shape_path = args[1]
        elif key == "tol":
            n = int(args[0])
            vals = []
            while len(vals) < n and i < len(lines):      # the values follow on the next line(s), read with `file >> tolI`
                vals += [float(v) for v in lines[i].split()]
                i += 1
            s.tol = vals[:n]
        elif key == "warmStart":
            s.warm_start = int(args[0])
        elif key == "rotateModel":
            s.rot_axis = (float(args[0]), float(args[1]), float(args[2]))
            s.rot_deg = float(args[3])
        elif key == "handleRatio":
            s.handle_ratio = float(args[0])
        elif key == "restart":
            s.restart = args[0]
        elif key in ("timeIntegration", "inexactSolve", "resolution", "view", "zoom", "appendStr", "disableCout", "tuning"):
            if key == "tuning":                              # values on the following line(s)
                n = int(args[0])
                got = 0
                while got < n and i < len(lines):
                    got += len(lines[i].split())
                    i += 1
        else:
            s.unknown.append(key)
    return s


def read_msh(path: str):
    """Returns (V [nV,3] float64, T [nT,4] int32 0-based, SF [nS,3] int32 0-based or empty)."""
    with open(path) as f:
        lines = f.read().splitlines()
    i, n = 0, len(lines)

    def seek(tag):
        nonlocal i
        while i < n and not lines[i].startswith(tag):
            i += 1
        if i >= n:
            return False
        i += 1
        return True

    if not seek("$Nodes"):
        raise ValueError("no $Nodes section in %s" % path)
    nV = int(lines[i].split()[1])          # "1 <nV>"
    i += 2                                 # that line and the block header "0 3 0 <nV>"
    V = np.array([[float(v) for v in lines[i + k].split()[1:4]] for k in range(nV)], dtype=np.float64)
    i += nV
    if not seek("$Elements"):
        raise ValueError("no $Elements section in %s" % path)
    nT = int(lines[i].split()[1])
    i += 2
    T = np.array([[int(v) for v in lines[i + k].split()[1:5]] for k in range(nT)], dtype=np.int64) - 1
    i += nT
    SF = np.zeros((0, 3), dtype=np.int32)
    if seek("$Surface"):
        nS = int(lines[i].split()[0])
        i += 1
        if nS > 0:
            SF = np.array([[int(v) for v in lines[i + k].split()[:3]] for k in range(nS)], dtype=np.int64) - 1
    if T.min() < 0 or T.max() >= nV:
        raise ValueError("element index out of range in %s" % path)
    return V, T.astype(np.int32), SF.astype(np.int32)


def rotate_model(V: np.ndarray, axis, deg: float) -> np.ndarray:
    """main.cpp:692-707: Eigen::AngleAxis(deg/180*pi, axis) applied to every node (the axis is used as given, like the reference)."""
    if deg == 0.0:
        return V
    a = np.asarray(axis, dtype=np.float64)
    th = deg / 180.0 * math.pi
    c, s_ = math.cos(th), math.sin(th)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = c * np.eye(3) + s_ * K + (1 - c) * np.outer(a, a)
    return V @ R.T


def write_status(path: str, timestep: int, x: np.ndarray, velocity: np.ndarray, dx_elastic: np.ndarray | None = None) -> None:
    x = np.asarray(x, dtype=np.float64).reshape(-1, 3)
    v = np.asarray(velocity, dtype=np.float64).reshape(-1)
    d = np.zeros_like(x) if dx_elastic is None else np.asarray(dx_elastic, dtype=np.float64).reshape(-1, 3)
    with open(path, "w") as f:
        f.write("timestep %d\n" % timestep)
        f.write("\nposition %d %d\n" % (x.shape[0], 3))
        f.write("".join("%e %e %e\n" % (r[0], r[1], r[2]) for r in x))
        f.write("\nvelocity %d\n" % v.size)
        f.write("".join("%e\n" % r for r in v))
        f.write("\ndx_Elastic %d %d\n" % (d.shape[0], 3))
        f.write("".join("%e %e %e\n" % (r[0], r[1], r[2]) for r in d))


def write_msh_reference(path: str, V: np.ndarray, T: np.ndarray, SF: np.ndarray) -> None:
    """The .msh file exactly as IglUtils::saveTetMesh writes it (IglUtils.cpp:627-679; coordinates with %le, i.e. 7 significant
    digits - `meshgen.write_msh` writes the same dialect with %.17g for loss-free round trips)."""
    lo, hi = V.min(axis=0), V.max(axis=0)
    with open(path, "w") as f:
        f.write("$MeshFormat\n4 0 8\n$EndMeshFormat\n$Entities\n0 0 0 1\n")
        f.write("0 %e %e %e %e %e %e 0 0\n$EndEntities\n" % (lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]))
        f.write("$Nodes\n1 %d\n0 3 0 %d\n" % (V.shape[0], V.shape[0]))
        f.write("".join("%d %e %e %e\n" % (i + 1, v[0], v[1], v[2]) for i, v in enumerate(V)))
        f.write("$EndNodes\n$Elements\n1 %d\n0 3 4 %d\n" % (T.shape[0], T.shape[0]))
        f.write("".join("%d %d %d %d %d\n" % (i + 1, t[0] + 1, t[1] + 1, t[2] + 1, t[3] + 1) for i, t in enumerate(T)))
        f.write("$EndElements\n$Surface\n%d\n" % SF.shape[0])
        f.write("".join("%d %d %d\n" % (s_[0] + 1, s_[1] + 1, s_[2] + 1) for s_ in SF))
        f.write("$EndSurface\n")


def read_status(path: str):
    """Returns dict(timestep, position [nV,3], velocity [3 nV], dx_Elastic [nV,3] or None)."""
    toks = open(path).read().split()
    out = {"timestep": 0, "position": None, "velocity": None, "dx_Elastic": None}
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "timestep":
            out["timestep"] = int(toks[i + 1])
            i += 2
        elif t == "position":
            r, c = int(toks[i + 1]), int(toks[i + 2])
            out["position"] = np.array(toks[i + 3:i + 3 + r * c], dtype=np.float64).reshape(r, c)
            i += 3 + r * c
        elif t == "velocity":
            n = int(toks[i + 1])
            out["velocity"] = np.array(toks[i + 2:i + 2 + n], dtype=np.float64)
            i += 2 + n
        elif t == "dx_Elastic":
            r, c = int(toks[i + 1]), int(toks[i + 2])
            out["dx_Elastic"] = np.array(toks[i + 3:i + 3 + r * c], dtype=np.float64).reshape(r, c)
            i += 3 + r * c
        else:
            i += 1
    return out


class IterStatsWriter:
    """iterStats.txt exactly as the reference's steppers stream it (default ostream formatting = %g):
    `timeStepper DOT`    -> `<frame> <alpha> <E> <|g|^2>`          (DOTTimeStepper.cpp:298, 306, 329; Optimizer.cpp:862)
    `timeStepper Newton` -> `<frame> <alpha> <E> <|g|^2> 0`        (Optimizer::fullyImplicit, Optimizer.cpp:666-684)"""

    def __init__(self, path: str, time_stepper: str = "DOT"):
        self.f = open(path, "w")
        self.tail = " 0" if time_stepper == "Newton" else ""

    def frame(self, frame_index: int, log: np.ndarray) -> None:
        """log: rows (alpha, E, |g|^2) of one time step, row 0 = after initX (what Stepper.iter_log() returns)."""
        for a, E, gg in np.asarray(log).reshape(-1, 3):
            self.f.write("%d %g %g %g%s\n" % (frame_index, a, E, gg, self.tail))
        self.f.flush()

    def close(self):
        self.f.close()


# the reference's per-activity timers (main.cpp:867-890): timer ("descent"), timer_step (14 activities), timer_temp3 (7 ADMM activities)
TIMER_STEP_NAMES = ["matrixComputation", "matrixAssembly", "symbolicFactorization", "numericalFactorization", "backSolve", "lineSearch_other",
                    "modifyGrad", "modifySearchDir", "updateHistory", "lineSearch_eVal", "fullyImplicit_eComp", "solve_extraComp", "compGrad", "CCD"]
TIMER_TEMP3_NAMES = ["init", "initPrimal", "initDual", "initWeights", "initCons", "subdSolve", "consSolve"]


def _timer_block(names, secs) -> str:
    """Timer::print (Utils/Timer.hpp:58-69): '<n> activities:', one right-aligned width-10 '%g' value + ' s: <name>' per activity, the total."""
    out = ["%d activities:" % len(names)]
    out += ["%10s s: %s" % ("%g" % float(t), n) for n, t in zip(names, secs)]
    out.append("%10s s: Total" % ("%g" % float(sum(secs))))
    return "\n".join(out) + "\n"


def write_info_txt(path: str, n_verts: int, n_tets: int, iter_num: int, inner_iters: int, descent_sec: float, step_sec: dict | None = None) -> None:
    """`info.txt` as main.cpp::saveInfoForPresent writes it (main.cpp:338-358): sizes, outer / inner iteration counts, the three timers,
    a trailing '0 0'.  step_sec maps names of TIMER_STEP_NAMES to seconds (missing: 0); the ADMM timer block is all zeros here."""
    step_sec = step_sec or {}
    unknown = set(step_sec) - set(TIMER_STEP_NAMES)
    if unknown:
        raise ValueError("not a timer_step activity: %s" % sorted(unknown))
    with open(path, "w") as f:
        f.write("%d %d\n" % (n_verts, n_tets))
        f.write("%d %d 0 0 0\n" % (iter_num, inner_iters))                      # 1.0 - energyParams[0] = 0
        f.write(_timer_block(["descent"], [descent_sec]))
        f.write(_timer_block(TIMER_STEP_NAMES, [step_sec.get(n, 0.0) for n in TIMER_STEP_NAMES]))
        f.write(_timer_block(TIMER_TEMP3_NAMES, [0.0] * len(TIMER_TEMP3_NAMES)))
        f.write("0 0\n")


def write_label_obj(path: str, surface_tris: np.ndarray, tri_to_tet: np.ndarray, epart: np.ndarray) -> None:
    """One `v <label> 0 0` line per surface triangle, label = subdomain of the tet the triangle belongs to."""
    with open(path, "w") as f:
        for t in tri_to_tet[: surface_tris.shape[0]]:
            f.write("v %d 0 0\n" % int(epart[int(t)]))


def surface_to_tet(T: np.ndarray, SF: np.ndarray) -> np.ndarray:
    """IglUtils::buildSTri2Tet (IglUtils.cpp:591-625): the tet that owns every surface triangle."""
    key = {}
    for t, (a, b, c, d) in enumerate(T.tolist()):
        for tri in ((a, c, b), (a, d, c), (a, b, d), (b, c, d)):
            key[tuple(sorted(tri))] = t
    return np.array([key[tuple(sorted(tri))] for tri in SF.tolist()], dtype=np.int32)


def partition_rcb(V: np.ndarray, T: np.ndarray, k: int) -> np.ndarray:
    """Recursive coordinate bisection of the tet centroids into k parts.  A stand-in for runs where the reference's METIS
    labels are not available (METIS cannot be shipped with this repo): the result is a valid decomposition, NOT the
    reference's labels - pass `--labels` for bit-exact subdomains."""
    c = V[T].mean(axis=1)
    part = np.zeros(T.shape[0], dtype=np.int32)

    def split(idx, lo, hi):
        if hi - lo <= 1:
            part[idx] = lo
            return
        kl = (hi - lo) // 2
        ax = int(np.argmax(c[idx].max(axis=0) - c[idx].min(axis=0)))
        order = idx[np.argsort(c[idx, ax], kind="stable")]
        cut = int(round(len(order) * kl / (hi - lo)))
        split(order[:cut], lo, lo + kl)
        split(order[cut:], lo + kl, hi)

    split(np.arange(T.shape[0]), 0, k)
    return part
