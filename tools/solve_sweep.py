"""GPU box: time one preconditioner application (K5) for several ring configurations."""
import os, sys, json, itertools, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dot_b200 as D
from bench import load_workload, DT
out = []
for wlname in sys.argv[1:] or ["bar17K_like", "bar1M"]:
    wl = load_workload(wlname)
    anim = D.Anim(wl["anim"], wl["V"]); fm = anim.fixed_mask()
    for stage, nst, grp, dbg in [(2560, 2, 4, 0), (1792, 2, 4, 0), (1920, 2, 4, 0), (2048, 2, 4, 0), (3072, 2, 4, 0), (1536, 3, 4, 0)]:
        os.environ["DOTGPU_SOLVE_STAGE_DBL"] = str(stage); os.environ["DOTGPU_SOLVE_NSTAGE"] = str(nst); os.environ["DOTGPU_SOLVE_GROUP"] = str(grp); os.environ["DOTGPU_SOLVE_DBG"] = str(dbg)
        try:
            stp = D.Stepper(wl["V"], wl["T"], wl["epart"], fm, energy=wl["energy"], k=wl["k"], dt=DT)
            ms = min(stp.time_kernels(5, 20) for _ in range(3))
            del stp
        except Exception as e:
            ms = str(e)
        out.append((wlname, stage, nst, grp, dbg, ms)); print(out[-1], flush=True)
json.dump(out, open("gpurun_out/solve_sweep.json", "w"))
