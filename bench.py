#!/usr/bin/env python
"""bench.py - DOT time stepping on B200 (metric of BASELINE.json: simulated frames/s, Newton-converged).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one converged time step (frame) of the scripted boundary motion.  Default workload = BASELINE config C4:
the synthetic 1M-tet bar twist (1,029,000 tets / 182,736 nodes), Stable Neo-Hookean, 128 METIS subdomains, dt=0.025 -
the largest single-GPU configuration and the one north_star's scaling target names.  The same line carries a
`secondary` record for config C2 on the reference's own mesh (bar17K.msh: 86,058 tets, SNH, 8 subdomains; fixture
tests/golden/mesh_bar17K.npz) and `parity` objects: both arms (this library and the unmodified reference binary) run the
same script from rest at a tight tolerance and the positions are compared.  METIS labels of every workload were produced by
the reference's own wrapper and are committed under tests/golden/.  Other workloads: bunny5K (C1: FCR k=6 twistnsns, and
bunny5K_PN = Projected Newton), horse38K_tb4 (C5: SNH, dt=1/24, k=16, twistnsns_old), bar17K_like / bar5K_like /
bar136K_like (structured stand-ins of round 1).

JSON line: value = frames/s with positions resident on the device (only the scripted Dirichlet targets travel),
e2e = frames/s through dotgpu_stepper_frame with host positions in and out every frame, roofline = the dominant
kernel group of a frame, kernels = per-kernel device times with their algorithmic bytes,
cpu_baseline = the unmodified reference (oracle/_ref/dot_ref: TBB->OpenMP shim, CHOLMOD + OpenBLAS) on this box's cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (mesh preset or fixture, energy, subdomains, anim, dt)
    "bar1M": ("bar1M", "SNH", 128, "twist", 0.025),                      # C4 (default)
    "bar17K": ("bar17K", "SNH", 8, "twist", 0.025),                      # C2 on the reference's own mesh
    "bunny5K": ("bunny5K", "FCR", 6, "twistnsns", 0.025),                # C1 as shipped (DOT 6)
    "bunny5K_PN": ("bunny5K", "FCR", 1, "twistnsns", 0.025),             # C1 "1 subdomain" = `timeStepper Newton`
    "horse38K_tb4": ("horse38K", "SNH", 16, "twistnsns_old", 0.0416667), # C5 (extreme deformation, halving-bound)
    "bar17K_like": ("bar17K_like", "SNH", 8, "twist", 0.025),
    "bar5K_like": ("bar5K_like", "FCR", 6, "twist", 0.025),
    "bar136K_like": ("bar136K_like", "FCR", 64, "twist", 0.025),
    "bar5K_like_PN": ("bar5K_like", "FCR", 1, "twist", 0.025),
}
# the reference arm / CPU legs are bounded samples: at most this many (warm-up, timed) frames per workload size
CPU_FRAME_CAP = {"bar1M": (1, 3), "bar136K_like": (1, 4), "horse38K_tb4": (1, 4)}


def load_workload(name):
    from dot_b200 import meshgen
    preset, energy, k, anim, dt = WORKLOADS[name]
    if preset in meshgen.PRESETS:
        V, T = meshgen.preset(preset)
        kind = "structured Kuhn bar"
    else:
        V, T = meshgen.load_mesh_npz(os.path.join(ROOT, "tests", "golden", "mesh_%s.npz" % preset))
        kind = "the reference's input/tetMeshes/%s.msh" % preset
    Vn = meshgen.normalise_like_loader(V)
    newton = name.endswith("_PN")
    if newton:
        ep = np.zeros(T.shape[0], dtype=np.int32)
    else:
        ep = np.load(os.path.join(ROOT, "tests", "golden", "labels_%s_k%d.npz" % (preset, k)))["epart"].astype(np.int32)
    return dict(name=name, preset=preset, energy=energy, k=k, anim=anim, dt=dt, V_raw=V, V=Vn, T=T, epart=ep, newton=newton, kind=kind)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, device=0):
        super().__init__(daemon=True)
        self.device, self.rows, self.stop_flag = device, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def run_reference(wl, steps, warmup, threads=None, tol=None, want_V=False):
    """The unmodified reference CPU path (oracle/_ref/dot_ref) on the same mesh / script; times frames warmup+1..warmup+steps.
    tol: relative tolerance override (parity runs use a tight one); want_V: also return the final positions."""
    from dot_b200 import meshgen
    exe = os.path.join(ROOT, "oracle", "_ref", "dot_ref")
    if not os.path.exists(exe):
        return None
    tmp = tempfile.mkdtemp(prefix="dotbench_")
    msh = os.path.join(tmp, "mesh.msh")
    meshgen.write_msh(msh, wl["V_raw"], wl["T"])
    script = os.path.join(tmp, "script.txt")
    meshgen.write_script(script, msh, energy=wl["energy"], parts=wl["k"], anim=wl["anim"], dt=wl["dt"], stepper="Newton" if wl["newton"] else "DOT")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    blas = os.path.join(ROOT, "oracle", "_ref", "blasdir.txt")
    if os.path.exists(blas):
        env["LD_LIBRARY_PATH"] = open(blas).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    cores = threads or os.cpu_count() or 1
    env["OMP_NUM_THREADS"] = str(cores)
    t0 = time.time()
    cmd = [exe, "--script", script, "--frames", str(warmup + steps), "--quiet", "--threads", str(cores)]
    if tol is not None:
        cmd += ["--tol", repr(tol)]
    if want_V:
        cmd += ["--final-V", os.path.join(tmp, "finalV.npy")]
    out = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True)
    if out.returncode != 0:
        return {"error": out.stderr[-400:]}
    js = [l for l in out.stdout.splitlines() if l.startswith("{")]
    st = json.loads(js[-1])
    sec = st["frame_sec"][warmup:]
    r = {"fps": len(sec) / sum(sec), "sec_per_frame": sum(sec) / len(sec), "cores": int(st["threads"]), "frames": len(sec),
         "inner_iters": int(sum(st["frame_iters"][warmup:])), "setup_sec": st["setup_sec"], "wall_sec": time.time() - t0,
         "timers_sec": st["timers_sec"], "sumV": st["sumV"], "halvings": st["line_search_halvings"], "frame_iters": st["frame_iters"]}
    if want_V:
        r["V"] = np.load(os.path.join(tmp, "finalV.npy"))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return r


def _emit(line):
    """Exactly one JSON line on the real stdout (libraries such as NCCL print banners to fd 1: it is parked on stderr meanwhile)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def workload_config(name, wl, steps, warmup, world):
    nT, nV = wl["T"].shape[0], wl["V"].shape[0]
    method = "Projected Newton (timeStepper Newton)" if wl["newton"] else "DOT %d subdomains (reference METIS labels)" % wl["k"]
    return {"workload": "%s: %s, %d tets / %d nodes, %s, %s, script %s, dt %g, tol 1e-5"
                        % (name, wl["kind"], nT, nV, wl["energy"], method, wl["anim"], wl["dt"]),
            "l2": "inputs larger than L2, no explicit flush: every L-BFGS iteration streams the solve panels of all subdomains (2 x nnz(L) x 8 B "
                  "= 2.4 GB on bar1M, 0.2 GB on bar17K vs 126 MB of L2; the exact figure is roofline.algorithmic_bytes_per_launch) and every "
                  "frame rewrites ~4 x that in the Hessian refresh; positions / gradients (3 nV doubles) are L2-resident by design",
            "subdomains": wl["k"], "tets": nT, "nodes": nV, "energy": wl["energy"], "frames_timed": steps, "frames_warmup": warmup,
            "parallelism": "%d GPU(s): subdomains (factor + solves) and tets (energy / gradient / Hessians) sharded by subdomain, balanced by nnz(L); "
                           "search direction and gradient reduced across ranks every L-BFGS iteration" % world}


class Ctx:
    """Process-wide plumbing: torch.distributed for the rendezvous, barrier and max-over-ranks; everything timed runs in libdotgpu."""

    def __init__(self):
        import torch
        import dot_b200 as D
        self.torch, self.D = torch, D
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def fresh_nccl_id(self):
        """Every communicator needs its own ncclUniqueId: rank 0 creates one, torch.distributed broadcasts the 128 bytes."""
        if self.world == 1:
            return None
        torch = self.torch
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            buf.copy_(torch.frombuffer(bytearray(self.D.nccl_unique_id()), dtype=torch.uint8))
        torch.distributed.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.torch.distributed.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        tt = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.torch.distributed.all_reduce(tt, op=self.torch.distributed.ReduceOp.MAX)
        return [float(v) for v in tt]

    def make(self, wl, rel_tol=1e-5):
        return self.D.Stepper(wl["V"], wl["T"], wl["epart"], self.D.Anim(wl["anim"], wl["V"]).fixed_mask(), energy=wl["energy"], k=wl["k"],
                              dt=wl["dt"], device=self.local_rank, rank=self.rank, world=self.world, nccl_id=self.fresh_nccl_id(),
                              newton=wl["newton"], rel_tol=rel_tol)


def gpu_frames(ctx, wl, steps, warmup, profile_mode=False, with_kernels=True, with_e2e=True):
    """value (resident positions), e2e (host positions in/out every frame), per-kernel times, K5 roofline inputs."""
    D, torch = ctx.D, ctx.torch
    nT, nV = wl["T"].shape[0], wl["V"].shape[0]
    dt = wl["dt"]
    anim0 = D.Anim(wl["anim"], wl["V"])
    fixed_idx = np.nonzero(anim0.fixed_mask())[0].astype(np.int32)
    t_setup = time.time()
    stp = ctx.make(wl)
    t_setup = time.time() - t_setup
    an = D.Anim(wl["anim"], wl["V"])
    x = wl["V"].copy()
    for f in range(warmup):
        an.step(x, dt)
        stp.frame_resident(fixed_idx, x[fixed_idx])
        x[:] = stp.get_state()[0]
    ctx.barrier()
    l0 = stp.launch_count()
    t0 = time.perf_counter()
    acc = dict(iters=0, halv=0, dev_ms=0.0, solve_ms=0.0, refresh_ms=0.0, pc_ms=0.0, pc_calls=0, conv=True)
    for f in range(steps):
        # the scripted handle motion only needs the handle rows, which the solve leaves where the script put them
        an.step(x, dt)
        st = stp.frame_resident(fixed_idx, x[fixed_idx])
        acc["iters"] += st.iters
        acc["halv"] += st.halvings
        acc["dev_ms"] += st.ms_total
        acc["solve_ms"] += st.ms_solve
        acc["refresh_ms"] += st.ms_refresh
        acc["pc_ms"] += st.ms_precond
        acc["pc_calls"] += st.precond_calls
        acc["conv"] = acc["conv"] and bool(st.converged)
    ctx.barrier()
    t_value = time.perf_counter() - t0
    launches = stp.launch_count() - l0
    t_value, acc["dev_ms"], acc["pc_ms"] = ctx.max_over_ranks([t_value, acc["dev_ms"], acc["pc_ms"]])
    x_final = stp.get_state()[0]
    out = dict(acc, t_value=t_value, launches=launches, setup_sec=t_setup, nT=nT, nV=nV)
    owned = stp.owned()
    infos = [stp.solver_info(s) for s in owned]
    out["nnz_l"] = sum(i.nnz_l for i in infos)
    out["n_sub"] = sum(i.n for i in infos)
    out["flops"] = sum(i.flops for i in infos)
    out["owned_subdomains"] = len(owned)
    out["fill_stats"] = stp.fill_stats()
    if with_kernels:
        # per-kernel device times on the final state (CUDA events on the stepper's stream)
        names = ["energy", "gradient", "elem_hessians", "fill", "factorize", "precondition", "dot"]
        out["kms"] = {n: stp.time_kernels(i, 1 if profile_mode else 20) for i, n in enumerate(names)}
    del stp
    if with_e2e:
        # e2e: host positions in / out every frame through dotgpu_stepper_frame (pinned host buffer)
        stp = ctx.make(wl)
        an = D.Anim(wl["anim"], wl["V"])
        pinned = torch.empty((nV, 3), dtype=torch.float64).pin_memory()
        xe = pinned.numpy()
        xe[:] = wl["V"]
        for f in range(warmup):
            an.step(xe, dt)
            stp.frame(xe)
        ctx.barrier()
        t0 = time.perf_counter()
        for f in range(steps):
            an.step(xe, dt)
            st = stp.frame(xe)
        out["loss"] = float(st.E)       # device->host read of the step's result (energy) is part of frame()
        ctx.barrier()
        out["t_e2e"] = ctx.max_over_ranks([time.perf_counter() - t0])[0]
        out["e2e_vs_resident"] = float(np.abs(xe - x_final).max())
        del stp
    return out


def parity_run(ctx, wl, frames, tol):
    """Both arms from rest with the same script at a tight tolerance (with identical partition and start, tol 1e-9,
    max |dx| / bbox <= 1e-6 per frame); the longest bounding-box edge is 1 after the loader's normalisation.  The reference
    (oracle/_ref/dot_ref, the checker) runs on rank 0 only."""
    D = ctx.D
    stp = ctx.make(wl, rel_tol=tol)
    an = D.Anim(wl["anim"], wl["V"])
    x = wl["V"].copy()
    it = hv = 0
    conv = True
    for f in range(frames):
        an.step(x, wl["dt"])
        st = stp.frame(x)
        it += st.iters
        hv += st.halvings
        conv = conv and bool(st.converged)
    del stp
    ctx.barrier()
    if ctx.rank != 0:
        return None
    r = run_reference(wl, frames, 0, tol=tol, want_V=True)
    if not r or "error" in r:
        return {"error": (r or {}).get("error", "oracle/_ref/dot_ref missing")}
    return {"max_abs_dx_over_bbox": float(np.abs(x - r["V"]).max()), "frames": frames, "tol": tol, "iters_gpu": it, "iters_ref": r["inner_iters"],
            "halvings_gpu": hv, "halvings_ref": r["halvings"], "all_frames_converged_gpu": conv, "n_gpus": ctx.world,
            "reference": "unmodified reference binary, same script from rest, %d host threads" % r["cores"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="bar1M", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="dotgpu", choices=["dotgpu", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the bounded CPU-baseline sample (0: sized per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-C2 record (bar17K) carried next to the main workload")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="for runs under ncu: skip the per-kernel micro-timings (few launches)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = load_workload(a.workload)
    nT, nV = wl["T"].shape[0], wl["V"].shape[0]
    config = workload_config(a.workload, wl, a.steps, a.warmup, world)

    if a.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: the reference needs ~15 s per frame on the 1M-tet bar (+ ~30 s of std::map-heavy set-up)
        capw, caps = CPU_FRAME_CAP.get(a.workload, (a.warmup, a.steps))
        w, k = min(a.warmup, capw), min(a.steps, caps)
        r = run_reference(wl, k, w)
        if r is None or "error" in r:
            _emit({"impl": "reference", "unavailable": "oracle/_ref/dot_ref missing or failed: %s" % (r or {}).get("error", "not built")})
            return 0
        line = {"metric": "simulated frames/sec (Newton-converged)", "value": r["fps"], "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * r["sec_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "reference",
                                 "sample": "frames %d..%d of the same workload from rest (each step = one frame; bounded sample of the %d requested "
                                           "frames: %.0f s of CPU work incl. %.0f s set-up); unmodified reference (OpenMP shim for TBB, CHOLMOD 3.0.12, "
                                           "OpenBLAS 1 thread/solver)" % (w + 1, w + k, a.steps, r["wall_sec"], r["setup_sec"]),
                                 "inner_iters": r["inner_iters"], "timers_sec": r["timers_sec"]},
                "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        _emit(line)
        return 0

    ctx = Ctx()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    R = gpu_frames(ctx, wl, a.steps, a.warmup, a.profile_mode)
    sampler.stop_flag = True
    kms = R["kms"]
    names = list(kms)
    hbm, peak_src = peaks()
    bytes_per = {"energy": 117.0 * nT, "gradient": 129.0 * nT, "elem_hessians": 1000.0 * nT,
                 "precondition": 16.0 * R["nnz_l"] + 16.0 * R["n_sub"]}
    kernels = {}
    for n in names:
        kernels[n] = {"ms": kms[n]}
        if n in bytes_per:
            gbs = bytes_per[n] / (kms[n] * 1e-3) / 1e9
            kernels[n].update({"algorithmic_bytes": bytes_per[n], "GB/s": gbs, "frac_hbm": gbs / hbm})
    kernels["elem_hessians"]["note"] = "bytes moved by this layout: 280 B read + 720 B written per tet (10 unique 3x3 blocks); the reference's 12x12 layout would be 1432 B/tet"
    kernels["factorize"].update({"flops": R["flops"], "GFLOP/s": R["flops"] / (kms["factorize"] * 1e-3) / 1e9})
    pk = os.path.join(ROOT, "profiles", "r2", "dmma_peak.json")  # tools/dmma_peak.cu on a B200 of this pool: DMMA m8n8k4 = DFMA = 37.1 TFLOP/s
    if os.path.exists(pk):
        f64 = json.load(open(pk))["dmma_m8n8k4_f64_tflops"]
        kernels["factorize"].update({"fp64_peak_TFLOP/s": f64, "frac_fp64_peak": kernels["factorize"]["GFLOP/s"] / 1e3 / f64,
                                     "peak_source": "measured (tools/dmma_peak.cu, profiles/r2/dmma_peak.json)"})
    # K4: every stored matrix value is written once (8 B), every gathered elemental 3x3 block read once (72 B) + 4 B of gather list
    nnz_a, nblk, ngat = R["fill_stats"]
    fill_bytes = 8.0 * nnz_a + 76.0 * ngat
    kernels["fill"].update({"algorithmic_bytes": fill_bytes, "GB/s": fill_bytes / (kms["fill"] * 1e-3) / 1e9,
                            "frac_hbm": fill_bytes / (kms["fill"] * 1e-3) / 1e9 / hbm,
                            "note": "%d stored values written (8 B), %d elemental 3x3 blocks gathered (72 B + 4 B index) into %d matrix blocks" % (nnz_a, ngat, nblk)})
    # the reference's timer_step activities (main.cpp:867-880) for the GPU side, per time step: in-frame CUDA-event times where the stepper
    # records them (backSolve = K5 + scatter, the Hessian refresh), the isolated kernel times x their launch counts for the rest
    its = R["iters"] / max(a.steps, 1)
    timers_gpu = {"matrixComputation": kms["elem_hessians"], "matrixAssembly": kms["fill"], "symbolicFactorization": 0.0,
                  "numericalFactorization": kms["factorize"], "backSolve": R["pc_ms"] / max(a.steps, 1),
                  "lineSearch_eVal+updateHistory": its * kms["gradient"],
                  "modifyGrad+modifySearchDir+lineSearch_other": max(0.0, R["solve_ms"] / max(a.steps, 1) - R["pc_ms"] / max(a.steps, 1) - its * kms["gradient"])}

    # secondary record: config C2 on the reference's own mesh, same code path, short
    secondary = None
    if not a.no_secondary and a.workload != "bar17K" and not a.profile_mode:
        wl2 = load_workload("bar17K")
        R2 = gpu_frames(ctx, wl2, a.steps, a.warmup, with_kernels=False)
        secondary = {"config": workload_config("bar17K", wl2, a.steps, a.warmup, world), "value": a.steps / R2["t_value"], "unit": "frames/s",
                     "e2e": a.steps / R2["t_e2e"], "ms_per_step": 1e3 * R2["t_value"] / a.steps, "inner_iters": R2["iters"],
                     "line_search_halvings": R2["halv"], "all_frames_converged": R2["conv"],
                     "k5_ms_per_application": R2["pc_ms"] / max(R2["pc_calls"], 1),
                     "k5_frac_hbm": (16.0 * R2["nnz_l"] + 16.0 * R2["n_sub"]) / (R2["pc_ms"] / max(R2["pc_calls"], 1) * 1e-3) / 1e9 / hbm,
                     "solve_ms_per_step": R2["solve_ms"] / a.steps, "refresh_ms_per_step": R2["refresh_ms"] / a.steps}
        if not a.no_parity:
            secondary["parity"] = parity_run(ctx, wl2, 3, 1e-9)
    parity = None
    if not a.no_parity and not a.profile_mode and world == 1:
        parity = parity_run(ctx, wl, 1 if nT > 300000 else 3, 1e-9)

    if rank != 0:
        if world > 1:
            ctx.torch.distributed.destroy_process_group()
        return 0
    cpu = None
    if not a.no_cpu_baseline and world == 1 and not a.profile_mode:
        frames = a.cpu_frames or CPU_FRAME_CAP.get(a.workload, (0, 24))[1]
        r_cpu = run_reference(wl, frames, 0)
        if r_cpu and "error" not in r_cpu:
            one = None
            if nT < 300000:
                try:  # scaling context: the same reference on ONE host thread, short sample
                    r1 = run_reference(wl, max(2, frames // 6), 0, threads=1)
                    if r1 and "error" not in r1:
                        one = {"value": r1["fps"], "frames": r1["frames"], "inner_iters": r1["inner_iters"]}
                except Exception:
                    one = None
            cpu = {"value": r_cpu["fps"], "unit": "frames/s", "cores": r_cpu["cores"], "kind": "reference", "one_thread": one,
                   "sample": "first %d frames of the same workload from rest (%.1f s of CPU work incl. %.1f s set-up); unmodified reference, OpenMP shim "
                             "for TBB, CHOLMOD 3.0.12 + OpenBLAS (1 thread per solver)" % (r_cpu["frames"], r_cpu["wall_sec"], r_cpu["setup_sec"]),
                   "inner_iters": r_cpu["inner_iters"], "ms_per_iter": 1e3 * r_cpu["frames"] / r_cpu["fps"] / max(r_cpu["inner_iters"], 1),
                   "timers_sec": r_cpu["timers_sec"]}
    # roofline of the dominant kernel group (K5): algorithmic bytes per application = 2 sweeps x 8 B x nnz(L) + 2 x 8 B x n
    # (SURVEY.md 8(d)); duration = the average over the preconditioner applications INSIDE the timed frames (CUDA events on
    # the stepper's stream, recorded around every application), not a separate micro-benchmark.
    k5_bytes = bytes_per["precondition"]
    k5_ms = R["pc_ms"] / max(R["pc_calls"], 1)
    k5_gbs = k5_bytes / (k5_ms * 1e-3) / 1e9 if k5_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(a.workload, {}).get("k_solve_dram_bytes_per_launch")
    roof = {"bound": "hbm", "kernel": "K5 k_solve_stream: all per-subdomain supernodal triangular solves of one preconditioner application "
                                      "(forward + backward) as one persistent TMA-streamed dataflow kernel (the right-hand-side gather and the "
                                      "scatter/average [+ exchange at N > 1] kernels are included in the timed span); rank 0's share at N > 1",
            "achieved": k5_gbs, "peak": hbm, "unit": "GB/s", "frac": k5_gbs / hbm, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": k5_bytes, "avg_launch_ms_in_timed_region": k5_ms, "launches_in_timed_region": R["pc_calls"],
            "share_of_frame": R["pc_ms"] / max(R["dev_ms"], 1e-9), "isolated_ms": kms["precondition"]}
    line = {"metric": "simulated frames/sec (Newton-converged)", "value": a.steps / R["t_value"], "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * R["t_value"] / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": a.steps / R["t_e2e"], "unit": "frames/s", "h2d_bytes_per_step": 24 * nV, "d2h_bytes_per_step": 24 * nV + 8,
                    "max_abs_diff_vs_resident_run": R["e2e_vs_resident"]},
            "gpu_launches": R["launches"], "clocks": sampler.summary(), "roofline": roof, "kernels": kernels,
            "assembly_tets_per_s": {"energy+gradient": nT / ((kms["energy"] + kms["gradient"]) * 1e-3),
                                    "hessian+fill": nT / ((kms["elem_hessians"] + kms["fill"]) * 1e-3)},
            "inner_iters": R["iters"], "line_search_halvings": R["halv"], "all_frames_converged": R["conv"], "device_ms_per_step": R["dev_ms"] / a.steps,
            "solve_ms_per_step": R["solve_ms"] / a.steps, "refresh_ms_per_step": R["refresh_ms"] / a.steps, "setup_sec": R["setup_sec"],
            "nnz_L": R["nnz_l"], "factor_flops": R["flops"], "owned_subdomains_rank0": R["owned_subdomains"],
            "timers_ms_per_step": timers_gpu}
    if parity is not None:
        line["parity"] = parity
    if secondary is not None:
        line["secondary"] = secondary
    if cpu:
        line["cpu_baseline"] = cpu
    _emit(line)
    if world > 1:
        ctx.torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
