"""Headless runner for the reference's scripts on the GPU path:  python -m dot_b200 <script.txt> [options]

Does what the reference's `DOT_bin 100 <script>` does for the scripts inside the GPU path (energy FCR|SNH, timeStepper
DOT k | Newton, `shape input <msh>`): load + rotate + normalise the mesh (main.cpp:673-712), scripted Dirichlet motion, one
converged time step per frame, `iterStats.txt`, restart files and the final mesh in the output folder.

Subdomain labels: `dotgpu_partition` (the reference's vendored METIS 5.1.0 behind the C ABI, bit-exact labels) when libdotmetis.so
has been built; otherwise give labels with `--labels file.npy|.npz` or fall back to a coordinate-bisection decomposition
(`--partition rcb`: valid, converges to the same frames, but not the reference's labels).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m dot_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("script")
    ap.add_argument("--frames", type=int, default=None, help="time steps to run (default: duration / dt of the script)")
    ap.add_argument("--labels", default=None, help=".npy/.npz (key `epart`) with one subdomain label per tet")
    ap.add_argument("--partition", default="rcb", choices=["rcb"], help="fallback partitioner when --labels is not given")
    ap.add_argument("--out", default="output", help="output folder")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--save-every", type=int, default=0, help="write a restart file every N frames (0: only at the end)")
    a = ap.parse_args(argv)

    import dot_b200 as D
    from dot_b200 import io, meshgen

    s = io.parse_script(a.script)
    if s.shape != "input":
        sys.exit("only `shape input <msh>` scripts are supported")
    msh = s.input_shape_path if os.path.isabs(s.input_shape_path) or os.path.exists(s.input_shape_path) \
        else os.path.join(os.path.dirname(os.path.abspath(a.script)), s.input_shape_path)
    V, T, SF = io.read_msh(msh)
    V = io.rotate_model(V, s.rot_axis, s.rot_deg)
    V = meshgen.normalise_like_loader(V, s.size)
    newton = s.time_stepper == "Newton"
    if s.warm_start != 2:
        sys.exit("warmStart %d is outside the GPU path: the resident stepper implements initX(2) (x^n + dt v^n + dt^2 g), Optimizer.cpp:472-493" % s.warm_start)
    if s.unknown:
        print("warning: script keys ignored by the GPU path: %s" % ", ".join(s.unknown), file=sys.stderr)
    node_part = None
    if newton or s.time_stepper == "LBFGSH":
        k, epart = 1, np.zeros(T.shape[0], dtype=np.int32)
    elif s.time_stepper == "LBFGSJH":
        k, epart = s.partitions, np.zeros(T.shape[0], dtype=np.int32)
        node_part = D.partition_nodes(V.shape[0], T, k)      # METIS<3>::partMesh_nodes (needs libdotmetis.so)
    else:
        k = s.partitions if s.block_size <= 0 else V.shape[0] // s.block_size + 1
        if a.labels:
            z = np.load(a.labels)
            epart = (z["epart"] if hasattr(z, "files") else z).astype(np.int32)
            if epart.shape[0] != T.shape[0] or epart.min() < 0 or epart.max() >= k:
                sys.exit("labels do not match the mesh / partition count")
        else:
            try:   # the reference's vendored METIS behind the C ABI (bit-exact labels); RCB only where libdotmetis.so is absent
                epart = D.partition(V.shape[0], T, k)
            except D.DotGpuError:
                epart = io.partition_rcb(V, T, k)
    anim = D.Anim(s.script, V, s.handle_ratio)
    fixed = anim.fixed_mask()
    if s.script == "null":
        fixed[:] = 0
        fixed[0] = 1                      # Mesh<3> default fixedVert = {0} (Mesh.cpp:593-599)
    gravity = (0.0, -9.80665, 0.0) if s.with_gravity else (0.0, 0.0, 0.0)
    t0 = time.time()
    stp = D.Stepper(V, T, epart, fixed, energy=s.energy, k=k, dt=s.dt, device=a.device, rel_tol=s.rel_tol(0), YM=s.YM, PR=s.PR, rho=s.rho,
                    newton=newton, gravity=gravity, method=s.time_stepper, node_part=node_part)
    x = V.copy()
    frame0 = 0
    if s.restart:
        st = io.read_status(s.restart)
        x = np.ascontiguousarray(st["position"])
        stp.set_state(x, st["velocity"])
        frame0 = st["timestep"]
    setup = time.time() - t0
    os.makedirs(a.out, exist_ok=True)
    stats = io.IterStatsWriter(os.path.join(a.out, "iterStats.txt"), s.time_stepper)
    if SF.shape[0] and s.time_stepper == "DOT":
        io.write_label_obj(os.path.join(a.out, "label.obj"), SF, io.surface_to_tet(T, SF), epart)
    nframes = a.frames if a.frames is not None else s.num_frames() - frame0
    iters = 0
    t0 = time.time()
    cur_tol = s.rel_tol(0)
    dxe = np.zeros_like(x)
    t_step = {"numericalFactorization": 0.0, "backSolve": 0.0, "lineSearch_eVal": 0.0}
    for f in range(frame0, frame0 + nframes):
        if s.rel_tol(f) != cur_tol:          # Optimizer::setRelGL2Tol only when the script changes it (recomputes targetGRes)
            cur_tol = s.rel_tol(f)
            stp.set_rel_tol(cur_tol)
        xt_prev = stp.get_state()[2]          # xTilta of this time step (dx_Elastic = x - xTilta, Optimizer.cpp:356)
        anim.step(x, s.dt)
        if anim.changed:                      # the script changed the Dirichlet set: updatePrecondMtrAndFactorize (Optimizer.cpp:334-336)
            stp.set_fixed(anim.fixed_mask(), x)
        fs = stp.frame(x)
        dxe = x - xt_prev
        iters += fs.iters
        # device times of this time step under the reference's timer_step names (the Hessian refresh = matrix computation +
        # assembly + numeric factorisation is one CUDA-event span; everything of the iteration besides the solves goes to eVal)
        t_step["numericalFactorization"] += fs.ms_refresh * 1e-3
        t_step["backSolve"] += fs.ms_precond * 1e-3
        t_step["lineSearch_eVal"] += max(0.0, fs.ms_solve - fs.ms_precond) * 1e-3
        stats.frame(f, stp.iter_log())
        if not fs.converged:
            print("frame %d did not converge (|g|^2 = %g > %g)" % (f, fs.grad_sqnorm, fs.target), file=sys.stderr)
        if a.save_every and (f + 1) % a.save_every == 0:
            xs, vs, _ = stp.get_state()
            io.write_status(os.path.join(a.out, "status%d" % (f + 1)), f + 1, xs, vs, dxe)
    wall = time.time() - t0
    stats.close()
    xs, vs, _ = stp.get_state()
    io.write_status(os.path.join(a.out, "status%d" % (frame0 + nframes)), frame0 + nframes, xs, vs, dxe)
    meshgen.write_msh(os.path.join(a.out, "finalResult_mesh.msh"), xs, T, SF if SF.shape[0] else None)
    info = {"frames": nframes, "inner_iters": iters, "fps": nframes / wall if wall > 0 else None, "setup_sec": setup, "nT": int(T.shape[0]),
            "nV": int(V.shape[0]), "parts": int(k), "energy": s.energy, "timeStepper": s.time_stepper, "sumV": float(xs.sum()),
            "sqnormV": float((xs ** 2).sum())}
    io.write_info_txt(os.path.join(a.out, "info.txt"), V.shape[0], T.shape[0], frame0 + nframes, iters, wall, t_step)  # main.cpp:338-358
    with open(os.path.join(a.out, "info.json"), "w") as f:
        json.dump(info, f)
    print(json.dumps(info))
    return 0


if __name__ == "__main__":
    sys.exit(main())
