"""GPU box: build the stepper of a workload and apply the preconditioner (K5) a few times - the target of ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dot_b200 as D
from bench import load_workload
wl = load_workload(sys.argv[1] if len(sys.argv) > 1 else "bar17K")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
which = [int(w) for w in sys.argv[3].split(",")] if len(sys.argv) > 3 else [5]
fm = D.Anim(wl["anim"], wl["V"]).fixed_mask()
stp = D.Stepper(wl["V"], wl["T"], wl["epart"], fm, energy=wl["energy"], k=wl["k"], dt=wl["dt"])
for w in which:
    print("MS", w, stp.time_kernels(w, reps))
