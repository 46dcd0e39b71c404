#!/usr/bin/env python
"""bench.py - DOT time stepping on B200 (metric of BASELINE.json: simulated frames/s, Newton-converged).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one converged time step (frame) of the scripted twist.  Default workload = BASELINE config C2:
the bar17K-sized synthetic bar (86,400 tets / 16,640 nodes), Stable Neo-Hookean, 8 METIS subdomains, dt=0.025
(the reference's bar17K.msh lives only under /root/reference/input, so the mesh is the structured Kuhn bar of
dot_b200/meshgen.py; its METIS labels were produced by the reference's own wrapper and are committed under
tests/golden/).  Other workloads: bar5K_like (C1 size, FCR k=6), bar136K_like (C3 size, FCR k=64), bar1M (C4,
SNH k=128).

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
    # name: (mesh preset, energy, subdomains, anim)
    "bar17K_like": ("bar17K_like", "SNH", 8, "twist"),
    "bar5K_like": ("bar5K_like", "FCR", 6, "twist"),
    "bar136K_like": ("bar136K_like", "FCR", 64, "twist"),
    "bar1M": ("bar1M", "SNH", 128, "twist"),
    # BASELINE config C1 ("1 subdomain" = the reference's `timeStepper Newton`): Projected Newton, FCR
    "bar5K_like_PN": ("bar5K_like", "FCR", 1, "twist"),
}
DT = 0.025


def load_workload(name):
    from dot_b200 import meshgen
    preset, energy, k, anim = WORKLOADS[name]
    V, T = meshgen.preset(preset)
    Vn = meshgen.normalise_like_loader(V)
    newton = name.endswith("_PN")
    if newton:
        ep = np.zeros(T.shape[0], dtype=np.int32)
    else:
        ep = np.load(os.path.join(ROOT, "tests", "golden", "labels_%s_k%d.npz" % (preset, k)))["epart"].astype(np.int32)
    return dict(name=name, preset=preset, energy=energy, k=k, anim=anim, V_raw=V, V=Vn, T=T, epart=ep, newton=newton)


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


def run_reference(wl, steps, warmup, threads=None):
    """The unmodified reference CPU path (oracle/_ref/dot_ref) on the same mesh / script; times frames warmup+1..warmup+steps."""
    from dot_b200 import meshgen
    exe = os.path.join(ROOT, "oracle", "_ref", "dot_ref")
    if not os.path.exists(exe):
        return None
    tmp = tempfile.mkdtemp(prefix="dotbench_")
    msh = os.path.join(tmp, "mesh.msh")
    meshgen.write_msh(msh, wl["V_raw"], wl["T"])
    script = os.path.join(tmp, "script.txt")
    meshgen.write_script(script, msh, energy=wl["energy"], parts=wl["k"], anim=wl["anim"], dt=DT, stepper="Newton" if wl["newton"] else "DOT")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    blas = os.path.join(ROOT, "oracle", "_ref", "blasdir.txt")
    if os.path.exists(blas):
        env["LD_LIBRARY_PATH"] = open(blas).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    cores = threads or os.cpu_count() or 1
    env["OMP_NUM_THREADS"] = str(cores)
    t0 = time.time()
    out = subprocess.run([exe, "--script", script, "--frames", str(warmup + steps), "--quiet", "--threads", str(cores)], cwd=tmp, env=env,
                         capture_output=True, text=True)
    if out.returncode != 0:
        return {"error": out.stderr[-400:]}
    js = [l for l in out.stdout.splitlines() if l.startswith("{")]
    st = json.loads(js[-1])
    sec = st["frame_sec"][warmup:]
    return {"fps": len(sec) / sum(sec), "sec_per_frame": sum(sec) / len(sec), "cores": int(st["threads"]), "frames": len(sec),
            "inner_iters": int(sum(st["frame_iters"][warmup:])), "setup_sec": st["setup_sec"], "wall_sec": time.time() - t0,
            "timers_sec": st["timers_sec"], "sumV": st["sumV"]}


def _emit(line):
    """Exactly one JSON line on the real stdout (libraries such as NCCL print banners to fd 1: it is parked on stderr meanwhile)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="bar17K_like", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="dotgpu", choices=["dotgpu", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames of the bounded CPU-baseline sample (~10 s of CPU work on bar17K_like)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="for runs under ncu: skip the per-kernel micro-timings (few launches)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = load_workload(a.workload)
    nT, nV = wl["T"].shape[0], wl["V"].shape[0]
    method = "Projected Newton (timeStepper Newton)" if wl["newton"] else "DOT %d subdomains (reference METIS labels)" % wl["k"]
    config = {"workload": "%s: structured Kuhn bar %d tets / %d nodes, %s, %s, script %s, dt %g, "
                          "tol 1e-5" % (a.workload, nT, nV, wl["energy"], method, wl["anim"], DT),
              "l2": "inputs larger than L2, no explicit flush: every L-BFGS iteration streams the solve panels of all subdomains (2 x nnz(L) x 8 B "
                    "= 271 MB on bar17K_like, 3.0 GB on bar1M vs 126 MB of L2) and every frame rewrites ~4 x that in the Hessian refresh; "
                    "positions / gradients (3 nV doubles) are L2-resident by design",
              "subdomains": wl["k"], "tets": nT, "nodes": nV, "energy": wl["energy"], "frames_timed": a.steps, "frames_warmup": a.warmup,
              "parallelism": "subdomains dealt round-robin to %d GPU(s); replicated per-tet kernels; one NCCL all-reduce of the search direction per L-BFGS iteration" % world}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference(wl, a.steps, a.warmup)
        if r is None or "error" in r:
            _emit({"impl": "reference", "unavailable": "oracle/_ref/dot_ref missing or failed: %s" % (r or {}).get("error", "not built")})
            return 0
        line = {"metric": "simulated frames/sec (Newton-converged)", "value": r["fps"], "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * r["sec_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "reference",
                                 "sample": "frames %d..%d of the same run, unmodified reference (OpenMP shim for TBB, CHOLMOD 3.0.12, OpenBLAS 1 thread/solver)" % (a.warmup + 1, a.warmup + a.steps),
                                 "inner_iters": r["inner_iters"], "timers_sec": r["timers_sec"]},
                "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        _emit(line)
        return 0

    import torch
    import dot_b200 as D
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def fresh_nccl_id():
        """Every communicator needs its own ncclUniqueId: rank 0 creates one, torch.distributed broadcasts the 128 bytes."""
        if world == 1:
            return None
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(D.nccl_unique_id()), dtype=torch.uint8))
        torch.distributed.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    anim = D.Anim(wl["anim"], wl["V"])
    fm = anim.fixed_mask()
    fixed_idx = np.nonzero(fm)[0].astype(np.int32)

    def make():
        return D.Stepper(wl["V"], wl["T"], wl["epart"], fm, energy=wl["energy"], k=wl["k"], dt=DT, device=local_rank, rank=rank, world=world,
                         nccl_id=fresh_nccl_id(), newton=wl["newton"])

    # ---------------- value: resident positions ----------------
    t_setup = time.time()
    stp = make()
    t_setup = time.time() - t_setup
    an = D.Anim(wl["anim"], wl["V"])
    x = wl["V"].copy()
    iters = halv = 0
    for f in range(a.warmup):
        an.step(x, DT)
        st = stp.frame_resident(fixed_idx, x[fixed_idx])
        xs = stp.get_state()[0]
        x[:] = xs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = stp.launch_count()
    t0 = time.perf_counter()
    dev_ms = solve_ms = refresh_ms = pc_ms = 0.0
    pc_calls = 0
    conv = True
    for f in range(a.steps):
        # the scripted handle motion only needs the handle rows, which the solve leaves where the script put them
        an.step(x, DT)
        st = stp.frame_resident(fixed_idx, x[fixed_idx])
        iters += st.iters
        halv += st.halvings
        dev_ms += st.ms_total
        solve_ms += st.ms_solve
        refresh_ms += st.ms_refresh
        pc_ms += st.ms_precond
        pc_calls += st.precond_calls
        conv = conv and bool(st.converged)
    barrier()
    t_value = time.perf_counter() - t0
    launches = stp.launch_count() - l0
    if world > 1:
        tt = torch.tensor([t_value, dev_ms, pc_ms], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_value, dev_ms, pc_ms = float(tt[0]), float(tt[1]), float(tt[2])
    x_final = stp.get_state()[0]

    # ---------------- per-kernel device times on the final state (CUDA events on the stepper's stream) ----------------
    names = ["energy", "gradient", "elem_hessians", "fill", "factorize", "precondition", "dot"]
    kms = {n: stp.time_kernels(i, 1 if a.profile_mode else 20) for i, n in enumerate(names)}
    infos = [stp.solver_info(s) for s in range(wl["k"]) if s % world == rank]
    nnz_l = sum(i.nnz_l for i in infos)
    n_sub = sum(i.n for i in infos)
    flops = sum(i.flops for i in infos)
    hbm, peak_src = peaks()
    r = nV / nT
    bytes_per = {"energy": 117.0 * nT, "gradient": 129.0 * nT, "elem_hessians": 1432.0 * nT, "precondition": 16.0 * nnz_l + 16.0 * n_sub}
    kernels = {}
    for n in names:
        kernels[n] = {"ms": kms[n]}
        if n in bytes_per:
            gbs = bytes_per[n] / (kms[n] * 1e-3) / 1e9
            kernels[n].update({"algorithmic_bytes": bytes_per[n], "GB/s": gbs, "frac_hbm": gbs / hbm})
    kernels["factorize"].update({"flops": flops, "GFLOP/s": flops / (kms["factorize"] * 1e-3) / 1e9})
    del stp

    # ---------------- e2e: host positions in / out every frame through dotgpu_stepper_frame ----------------
    stp = make()
    an = D.Anim(wl["anim"], wl["V"])
    x = wl["V"].copy()
    pinned = torch.empty((nV, 3), dtype=torch.float64).pin_memory()
    xe = pinned.numpy()
    xe[:] = x
    for f in range(a.warmup):
        an.step(xe, DT)
        stp.frame(xe)
    barrier()
    t0 = time.perf_counter()
    e_iters = 0
    for f in range(a.steps):
        an.step(xe, DT)
        st = stp.frame(xe)
        e_iters += st.iters
    loss = float(st.E)          # device->host read of the step's result (energy) is part of frame()
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        t_e2e = float(tt[0])
    same = float(np.abs(xe - x_final).max())
    sampler.stop_flag = True
    del stp

    if rank != 0:
        return 0
    cpu = None
    if not a.no_cpu_baseline and world == 1:
        r_cpu = run_reference(wl, a.cpu_frames, 0)
        if r_cpu and "error" not in r_cpu:
            one = None
            try:  # scaling context: the same reference on ONE host thread, short sample
                r1 = run_reference(wl, max(2, a.cpu_frames // 6), 0, threads=1)
                if r1 and "error" not in r1:
                    one = {"value": r1["fps"], "frames": r1["frames"], "inner_iters": r1["inner_iters"]}
            except Exception:
                one = None
            cpu = {"value": r_cpu["fps"], "unit": "frames/s", "cores": r_cpu["cores"], "kind": "reference", "one_thread": one,
                   "sample": "first %d frames of the same workload from rest (%.1f s of CPU work incl. %.1f s set-up); unmodified reference, OpenMP shim "
                             "for TBB, CHOLMOD 3.0.12 + OpenBLAS (1 thread per solver)" % (r_cpu["frames"], r_cpu["wall_sec"], r_cpu["setup_sec"]),
                   "inner_iters": r_cpu["inner_iters"], "ms_per_iter": 1e3 * r_cpu["frames"] / r_cpu["fps"] / max(r_cpu["inner_iters"], 1)}
    # roofline of the dominant kernel group (K5): algorithmic bytes per application = 2 sweeps x 8 B x nnz(L) + 2 x 8 B x n
    # (SURVEY.md 8(d)); duration = the average over the preconditioner applications INSIDE the timed frames (CUDA events on
    # the stepper's stream, recorded around every application), not a separate micro-benchmark.
    k5_bytes = bytes_per["precondition"]
    k5_ms = pc_ms / max(pc_calls, 1)
    k5_gbs = k5_bytes / (k5_ms * 1e-3) / 1e9 if k5_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(a.workload, {}).get("k_solve_stream_dram_bytes_per_launch")
    roof = {"bound": "hbm", "kernel": "K5 k_solve_stream: all per-subdomain supernodal triangular solves of one preconditioner application "
                                      "(forward + backward) as one persistent TMA-streamed dataflow kernel (+ the fused gather; the scatter/average kernel is included in the timed span)",
            "achieved": k5_gbs, "peak": hbm, "unit": "GB/s", "frac": k5_gbs / hbm, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": k5_bytes, "avg_launch_ms_in_timed_region": k5_ms, "launches_in_timed_region": pc_calls,
            "share_of_frame": pc_ms / max(dev_ms, 1e-9), "isolated_ms": kms["precondition"]}
    line = {"metric": "simulated frames/sec (Newton-converged)", "value": a.steps / t_value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t_value / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": a.steps / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": 24 * nV, "d2h_bytes_per_step": 24 * nV + 8,
                    "max_abs_diff_vs_resident_run": same},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "kernels": kernels,
            "assembly_tets_per_s": {"energy+gradient": nT / ((kms["energy"] + kms["gradient"]) * 1e-3),
                                    "hessian+fill": nT / ((kms["elem_hessians"] + kms["fill"]) * 1e-3)},
            "inner_iters": iters, "line_search_halvings": halv, "all_frames_converged": conv, "device_ms_per_step": dev_ms / a.steps,
            "solve_ms_per_step": solve_ms / a.steps, "refresh_ms_per_step": refresh_ms / a.steps, "setup_sec": t_setup,
            "nnz_L": nnz_l, "factor_flops": flops}
    if cpu:
        line["cpu_baseline"] = cpu
    _emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
