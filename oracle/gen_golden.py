#!/usr/bin/env python
"""TEST INFRASTRUCTURE - generates tests/golden/*.npz by running the UNMODIFIED reference
(oracle/_ref/dot_ref, built by oracle/ref/build_ref.sh from /root/reference) on small
synthetic bars.  Run it here (the container that has /root/reference); the GPU box only
sees the committed .npz files.

    python oracle/gen_golden.py            # all cases
    python oracle/gen_golden.py --labels   # also the METIS labels of the bench meshes

Every array in a fixture is the reference's own number: mesh set-up (restTriInv, triArea,
mass, CSR patterns of the global and per-subdomain matrices, METIS labels, local->global
maps, dup) and, per dumped state, positions, F, U, Sigma, V, per-element energies, the
assembled gradient, elemental PD-projected Hessians, CSR values after the Hessian refresh,
one preconditioner application p and H*p, plus iterStats.txt of the run.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dot_b200 import meshgen  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "dot_ref")
GOLD = os.path.join(ROOT, "tests", "golden")

# name, mesh preset, energy, parts, anim script, frames, dump frames, he_cap, dt
CASES = [
    ("tiny_snh_k4_twist", "bar_tiny", "SNH", 4, "twist", 6, [1, 3, 6], -1, 0.025),
    ("tiny_fcr_k4_twistnsns", "bar_tiny", "FCR", 4, "twistnsns", 6, [2, 6], -1, 0.025),
    ("small_snh_k4_twist", "bar_small", "SNH", 4, "twist", 5, [5], 48, 0.025),
    ("small_fcr_k3_stretch", "bar_small", "FCR", 3, "stretch", 5, [5], 48, 0.025),
    ("small_snh_k5_tsns_dt24", "bar_small", "SNH", 5, "twistnsns_old", 8, [8], 16, 0.0416667),
]
# back-tracking cases: a large time step makes the reference's line search halve (21 resp. 92 halvings in these runs,
# Optimizer.cpp:803-833); iterStats.txt is written with 17 significant digits (--full-precision)
HALVING_CASES = [
    ("small_snh_k4_twist_dt200", "bar_small", "SNH", 4, "twist", 6, [1, 6], 16, 0.2),
    ("small_fcr_k3_tsns_dt200", "bar_small", "FCR", 3, "twistnsns", 8, [1, 8], 16, 0.2),
]
# a15 / DOTTimeStepper::updatePrecondMtrAndFactorize: `script rubberBandPull` drags the waist of the bar 5 units sideways and RELEASES it
# at time step 81 (the Dirichlet set changes mid-run -> new patterns, symbolic analysis, factorisation); default tol, 17-digit iterStats
RUBBER_CASES = [("bar2K_snh_k4_rubberband", "bar2K", "SNH", 4, "rubberBandPull", 83, [79, 80, 81, 83], 4, 0.025)]
# the reference's own input meshes (BASELINE.json configs C1, C2, C5): nodes / tets as parsed from input/tetMeshes/*.msh
# plus the METIS labels of the reference's wrapper for the subdomain counts the configs name
MESH_CASES = [("bunny5K", [6]), ("bar17K", [8]), ("horse38K", [16])]
# Projected Newton (`timeStepper Newton`, Optimizer::solve_oneStep - the reference's "1 subdomain" case): same tuple, parts unused
PN_CASES = [
    ("small_fcr_newton_twist", "bar_small", "FCR", 4, "twist", 4, [1, 4], 16, 0.025),
    ("tiny_snh_newton_tsns", "bar_tiny", "SNH", 4, "twistnsns", 5, [1, 5], -1, 0.025),
]
# the L-BFGS initialisers that share the kernels - LBFGS-H (global projected Hessian of the start of the time step) and
# LBFGS-JH (block Jacobi of it over a METIS node partition), LBFGSTimeStepper.cpp:108-265, 339-420; 17-digit iterStats
LBFGS_CASES = [
    ("small_snh_lbfgsh_twist", "bar_small", "SNH", 4, "twist", 5, [1, 5], 16, 0.025, "LBFGSH"),
    ("small_fcr_lbfgsjh4_tsns", "bar_small", "FCR", 4, "twistnsns", 5, [1, 5], 16, 0.025, "LBFGSJH"),
]
# kernel-level states: name, mesh, energy, parts, perturbation amplitude (x cell size), seed
KERNEL_CASES = [
    ("tiny_fcr_inverted", "bar_tiny", "FCR", 4, 0.45, 12345),
    ("tiny_snh_inverted", "bar_tiny", "SNH", 4, 0.45, 12345),
    ("small_fcr_perturbed", "bar_small", "FCR", 4, 0.15, 7),
]
# METIS labels for the benchmark meshes (bench.py cannot call METIS on the GPU box)
LABEL_CASES = [("bar5K_like", 6), ("bar17K_like", 8), ("bar136K_like", 64), ("bar1M", 128)]


def run_ref(args, cwd):
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([REF] + args, cwd=cwd, env=env, check=True, capture_output=True, text=True).stdout
    last = [l for l in out.strip().splitlines() if l.startswith("{")]
    return json.loads(last[-1]) if last else {}


def collect(d):
    out = {}
    for f in sorted(glob.glob(os.path.join(d, "**", "*.npy"), recursive=True)):
        key = os.path.relpath(f, d)[:-4].replace(os.sep, "/")
        out[key] = np.load(f)
    return out


def make_mesh(tmp, preset):
    V, T = meshgen.preset(preset)
    msh = os.path.join(tmp, preset + ".msh")
    meshgen.write_msh(msh, V, T)
    return V, T, msh


def gen_case(name, preset, energy, parts, anim, frames, dumps, he_cap, dt, stepper="DOT", full_precision=False, tol=None):
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        V, T, msh = make_mesh(tmp, preset)
        script = os.path.join(tmp, "s.txt")
        meshgen.write_script(script, msh, energy=energy, parts=parts, anim=anim, dt=dt, stepper=stepper if stepper.startswith("LBFGS") else "DOT")
        dd = os.path.join(tmp, "dump")
        args = ["--script", script, "--frames", str(frames), "--quiet", "--dump-dir", dd,
                "--dump-frames", ",".join(map(str, dumps)), "--he-cap", str(he_cap)]
        if stepper == "Newton":
            args += ["--stepper", stepper]
        if full_precision:
            args += ["--full-precision"]
        if tol is not None:
            args += ["--tol", repr(tol)]
        stats = run_ref(args, tmp)
        arrs = collect(dd)
        arrs["iterStats"] = np.array(open(os.path.join(dd, "iterStats.txt")).read())
        arrs["meta"] = np.array(json.dumps(dict(name=name, preset=preset, energy=energy, parts=parts, anim=anim, stepper=stepper,
                                                frames=frames, dumps=dumps, dt=dt, stats={k: stats[k] for k in
                                                ("inner_iters", "sumV", "sqnormV", "line_search_halvings", "targetGRes", "frame_iters")})))
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrs)
        print(name, "->", len(arrs), "arrays", stats.get("inner_iters"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def gen_kernel_case(name, preset, energy, parts, amp, seed):
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        V, T, msh = make_mesh(tmp, preset)
        script = os.path.join(tmp, "s.txt")
        meshgen.write_script(script, msh, energy=energy, parts=parts, anim="twist")
        Vn = meshgen.normalise_like_loader(V)
        h = Vn[:, 1].max() / meshgen.PRESETS[preset][1]
        rng = np.random.default_rng(seed)
        X = Vn + amp * h * rng.uniform(-1, 1, Vn.shape)
        np.save(os.path.join(tmp, "state.npy"), np.ascontiguousarray(X))
        dd = os.path.join(tmp, "dump")
        run_ref(["--script", script, "--quiet", "--dump-dir", dd, "--kernel-state", os.path.join(tmp, "state.npy")], tmp)
        arrs = collect(dd)
        arrs["meta"] = np.array(json.dumps(dict(name=name, preset=preset, energy=energy, parts=parts, anim="twist", amp=amp, seed=seed, dt=0.025)))
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrs)
        ninv = int((arrs["kernel/Sigma"][:, 2] < 0).sum())
        print(name, "->", len(arrs), "arrays; inverted tets:", ninv)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def gen_mesh(name, parts_list):
    """tests/golden/mesh_<name>.npz = the reference's input/tetMeshes/<name>.msh as arrays (V float64 exactly as parsed,
    T int32 0-based) and labels_<name>_k<k>.npz from the reference's METIS wrapper on that mesh."""
    from dot_b200 import io as dio
    ref_root = os.environ.get("DOT_REFERENCE", "/root/reference")
    msh = os.path.join(ref_root, "input", "tetMeshes", name + ".msh")
    V, T, SF = dio.read_msh(msh)
    np.savez_compressed(os.path.join(GOLD, "mesh_%s.npz" % name), V=V, T=T.astype(np.int32))
    print("mesh", name, V.shape, T.shape)
    for parts in parts_list:
        tmp = tempfile.mkdtemp(prefix="golden_")
        try:
            script = os.path.join(tmp, "s.txt")
            meshgen.write_script(script, msh, energy="SNH", parts=parts, anim="twist")
            dd = os.path.join(tmp, "dump")
            run_ref(["--script", script, "--frames", "0", "--quiet", "--dump-dir", dd, "--labels-only"], tmp)
            ep = np.load(os.path.join(dd, "setup", "epart.npy"))
            assert ep.min() >= 0 and ep.max() == parts - 1
            np.savez_compressed(os.path.join(GOLD, "labels_%s_k%d.npz" % (name, parts)), epart=ep.astype(np.uint8))
            print("labels", name, parts, np.bincount(ep).tolist()[:8], "...")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


def gen_io_fixture():
    """tests/golden/io_small.npz: the files the REFERENCE writes for one short run - status<n> (Optimizer::saveStatus), iterStats.txt
    (DOT and Newton flavours), the final .msh (IglUtils::saveTetMesh) - as raw text, next to the same state as binary arrays, so that
    dot_b200/io.py's writers / readers can be compared with them byte for byte / field for field."""
    out = {}
    for stepper in ("DOT", "Newton"):
        tmp = tempfile.mkdtemp(prefix="golden_")
        try:
            V, T, msh = make_mesh(tmp, "bar_tiny")
            script = os.path.join(tmp, "s.txt")
            meshgen.write_script(script, msh, energy="SNH", parts=4, anim="twist", stepper=stepper)
            dd = os.path.join(tmp, "dump")
            stats = run_ref(["--script", script, "--frames", "3", "--quiet", "--dump-dir", dd, "--dump-frames", "3", "--he-cap", "1", "--save-status",
                             "--save-msh", os.path.join(tmp, "final.msh")], tmp)
            of = os.path.join(tmp, stats["output_folder"])
            key = stepper.lower()
            out[key + "/iterStats_txt"] = np.array(open(os.path.join(of, "iterStats.txt")).read())
            if stepper == "DOT":
                n = stats["timestep"]
                out["status_txt"] = np.array(open(os.path.join(of, "status%d" % n)).read())
                out["status_timestep"] = np.array(n)
                out["msh_txt"] = np.array(open(os.path.join(tmp, "final.msh")).read())
                out["V"] = np.load(os.path.join(dd, "frame3", "V.npy"))
                out["velocity"] = np.load(os.path.join(dd, "frame3", "velocity.npy"))
                out["T"] = T.astype(np.int32)
                out["SF"] = meshgen.surface_tris(T)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    np.savez_compressed(os.path.join(GOLD, "io_small.npz"), **out)
    print("io_small ->", sorted(out))


def gen_labels(preset, parts):
    """Labels only: run set-up (METIS + DD) and keep epart."""
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        V, T, msh = make_mesh(tmp, preset)
        script = os.path.join(tmp, "s.txt")
        meshgen.write_script(script, msh, energy="SNH", parts=parts, anim="twist")
        dd = os.path.join(tmp, "dump")
        run_ref(["--script", script, "--frames", "0", "--quiet", "--dump-dir", dd, "--labels-only"], tmp)
        ep = np.load(os.path.join(dd, "setup", "epart.npy"))
        assert ep.min() >= 0 and ep.max() == parts - 1
        np.savez_compressed(os.path.join(GOLD, "labels_%s_k%d.npz" % (preset, parts)),
                            epart=ep.astype(np.uint8 if parts <= 256 else np.int32))
        print("labels", preset, parts, np.bincount(ep).tolist()[:8], "...")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--labels", action="store_true")
    ap.add_argument("--io", action="store_true", help="files written by the reference (status, iterStats, .msh) as a fixture")
    ap.add_argument("--meshes", action="store_true", help="fixtures of the reference's own input meshes + their METIS labels")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    if not os.path.exists(REF):
        sys.exit("build oracle/_ref first: bash oracle/ref/build_ref.sh")
    for c in CASES:
        if not a.only or a.only in c[0]:
            gen_case(*c)
    for c in HALVING_CASES:
        if not a.only or a.only in c[0]:
            gen_case(*c, full_precision=True)
    for c in RUBBER_CASES:
        if not a.only or a.only in c[0]:
            gen_case(*c, full_precision=True)
            # keep the states only (positions, velocities, Dirichlet sets): the matrices of 4 dumped frames would be 7 MB
            pth = os.path.join(GOLD, c[0] + ".npz")
            z = np.load(pth)
            keep = {k: z[k] for k in z.files if k in ("meta", "iterStats", "setup/V_rest", "setup/F", "setup/epart") or
                    (k.startswith("frame") and k.split("/")[1] in ("V", "velocity", "fixed", "xTilta"))}
            np.savez_compressed(pth, **keep)
    for c in LBFGS_CASES:
        if not a.only or a.only in c[0]:
            gen_case(*c[:9], stepper=c[9], full_precision=True)
    for c in PN_CASES:
        if not a.only or a.only in c[0]:
            gen_case(*c, stepper="Newton")
    for c in KERNEL_CASES:
        if not a.only or a.only in c[0]:
            gen_kernel_case(*c)
    if a.io:
        gen_io_fixture()
    if a.meshes:
        for c in MESH_CASES:
            if not a.only or a.only in c[0]:
                gen_mesh(*c)
    if a.labels:
        for c in LABEL_CASES:
            if not a.only or a.only in c[0]:
                gen_labels(*c)
