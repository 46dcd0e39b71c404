"""GPU box: per-phase device times of one numeric factorisation (DOTGPU_FACTOR_TIMING)."""
import os, sys
os.environ["DOTGPU_FACTOR_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dot_b200 as D
from bench import load_workload
for wlname in sys.argv[1:] or ["bar17K_like"]:
    wl = load_workload(wlname)
    anim = D.Anim(wl["anim"], wl["V"]); fm = anim.fixed_mask()
    stp = D.Stepper(wl["V"], wl["T"], wl["epart"], fm, energy=wl["energy"], k=wl["k"], dt=wl["dt"])
    print(wlname, "factorize ms (3 reps each):", [stp.time_kernels(4, 1) for _ in range(3)], flush=True)
