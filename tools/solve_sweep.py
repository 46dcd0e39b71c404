"""GPU box: time one preconditioner application (K5) for several configurations (fresh process per configuration).
usage: python tools/solve_sweep.py <workload> ...   (DOTGPU_SWEEP_CONFIGS = JSON list of env dicts)"""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = json.loads(os.environ.get("DOTGPU_SWEEP_CONFIGS", '[{}, {"DOTGPU_SOLVE_NO_PREGATHER": 1}, {"DOTGPU_SOLVE_NSTAGE": 3}, {"DOTGPU_SOLVE_GROUP": 2}, {"DOTGPU_SOLVE_GROUP": 1}]'))
out = []
for wlname in sys.argv[1:] or ["bar17K", "bar1M"]:
    for cfg in CONFIGS:
        env = dict(os.environ)
        env.update({k: str(v) for k, v in cfg.items()})
        code = ("import sys; sys.path.insert(0, %r); import dot_b200 as D; from bench import load_workload; wl = load_workload(%r); "
                "fm = D.Anim(wl['anim'], wl['V']).fixed_mask(); "
                "stp = D.Stepper(wl['V'], wl['T'], wl['epart'], fm, energy=wl['energy'], k=wl['k'], dt=wl['dt']); "
                "print('MS', min(stp.time_kernels(5, 20) for _ in range(3)))" % (ROOT, wlname))
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
        ms = [l for l in r.stdout.splitlines() if l.startswith("MS")]
        ms = float(ms[-1].split()[1]) if ms else (r.stderr.strip().splitlines() or ["?"])[-1][-200:]
        out.append((wlname, cfg, ms))
        print(wlname, json.dumps(cfg), ms, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "solve_sweep.json"), "w"))
