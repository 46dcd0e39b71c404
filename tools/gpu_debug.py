"""Prints component-level errors of the CUDA path against golden/oracle (diagnostic aid, run on the GPU box)."""
import os, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import dot_b200 as D
from dot_b200 import meshgen
from golden_util import Golden, rel
from oracle import dot_oracle as O
import scipy.sparse.linalg as spla

def section(name, fn):
    t = time.time()
    try:
        fn()
        print("[ok  ] %s (%.2fs)" % (name, time.time() - t), flush=True)
    except Exception:
        print("[FAIL] %s" % name, flush=True)
        traceback.print_exc()

def energy_case(name):
    g = Golden(name); en, dt = g.meta["energy"], g.meta["dt"]
    V, T = g["setup/V_rest"], g["setup/F"]
    fm = np.zeros(V.shape[0], dtype=np.uint8); fm[g["setup/fixed"]] = 1
    e = D.Energy(en, T, g["setup/restTriInv"], g["setup/triArea"], g["setup/mu"], g["setup/lambda"], V.shape[0], fm)
    m = O.Mesh(V, T)
    st = g.states()[-1]; x = g[st + "/V"]
    fm[:] = 0; fm[g[st + "/fixed"]] = 1; e.set_fixed(fm)
    F, U, S, Vv = e.svd(x)
    print("  F", rel(F, g[st + "/F"]), "S", np.abs(S - g[st + "/Sigma"]).max(), "recon", np.abs(np.einsum("tia,ta,tja->tij", U, S, Vv) - F).max())
    print("  Epe", rel(e.energy_per_elem(x), g[st + "/E_per_elem"]), "E", e.compute_energy_val(x, dt*dt), g[st + "/E"][1])
    print("  g", rel(e.compute_gradient(x, dt*dt), g[st + "/g_elastic"]))
    He, vi = e.compute_elem_hessians(x, dt*dt, True)
    n = g[st + "/He"].shape[0]
    print("  He", rel(He[:n], g[st + "/He"]))

def solver_case(name, sub):
    g = Golden(name); st = g.states()[-1]
    pre = "global_" if sub < 0 else "sbd%d_" % sub
    ia, ja, a = g["setup/" + pre + "ia"], g["setup/" + pre + "ja"], g[st + "/" + pre + "a"]
    s = D.Solver(ia, ja); i = s.info()
    print("  n", i.n, "nsuper", i.nsuper, "levels", i.nlevels, "maxfront", i.max_front, "maxnscol", i.max_nscol, "nnzL", i.nnz_l, "flops", i.flops)
    s.set_values(a); s.factorize()
    A = O.csr_upper_to_full(ia, ja, a); b = np.random.default_rng(1).standard_normal(s.n)
    x = s.solve(b)
    print("  resid", np.linalg.norm(A @ x - b) / np.linalg.norm(b), "vs superlu", rel(x, spla.spsolve(A, b)), "spmv", rel(s.multiply(b), A @ b))

def stepper_case(name):
    g = Golden(name)
    V, T = g["setup/V_rest"], g["setup/F"]
    a = D.Anim(g.meta["anim"], V)
    stp = D.Stepper(V, T, g["setup/epart"], a.fixed_mask(), energy=g.meta["energy"], k=g.k, dt=g.meta["dt"])
    print("  target", stp.target, g.meta["stats"]["targetGRes"])
    st = g.states()[0]
    stp.set_state(g[st + "/V"], g[st + "/velocity"])
    print("  global_a", rel(stp.matrix(-1), g[st + "/global_a"]), [rel(stp.matrix(s), g[st + "/sbd%d_a" % s]) for s in range(g.k)])
    E, gr = stp.eval(g[st + "/V"]); print("  E", E, g[st + "/E"][0], "g", rel(gr, g[st + "/g"]))
    print("  p", rel(stp.precondition(-g[st + "/g"]), g[st + "/p"]))
    x = V.copy()
    stp2 = D.Stepper(V, T, g["setup/epart"], a.fixed_mask(), energy=g.meta["energy"], k=g.k, dt=g.meta["dt"])
    for f in range(1, g.meta["frames"] + 1):
        a.step(x, g.meta["dt"]); fs = stp2.frame(x)
        msg = "  frame %d iters %d (ref %d) conv %d halv %d ms %.2f/%.2f/%.2f" % (f, fs.iters, g.meta["stats"]["frame_iters"][f-1], fs.converged, fs.halvings, fs.ms_total, fs.ms_solve, fs.ms_refresh)
        if g.has("frame%d/V" % f): msg += " dx %.3e" % np.abs(x - g["frame%d/V" % f]).max()
        print(msg)

def big_case(preset, k, energy, frames=3):
    V, T = meshgen.preset(preset); V = meshgen.normalise_like_loader(V)
    ep = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "labels_%s_k%d.npz" % (preset, k)))["epart"].astype(np.int32)
    a = D.Anim("twist", V); t0 = time.time()
    stp = D.Stepper(V, T, ep, a.fixed_mask(), energy=energy, k=k)
    print("  setup %.2fs" % (time.time() - t0))
    for s in (0, k - 1):
        i = stp.solver_info(s); print("  sub", s, "n", i.n, "nsuper", i.nsuper, "levels", i.nlevels, "maxfront", i.max_front, "maxnscol", i.max_nscol, "nnzL", i.nnz_l, "GF", i.flops / 1e9, "dev MB", i.device_bytes / 1e6)
    x = V.copy()
    for f in range(frames):
        a.step(x, 0.025); t0 = time.time(); fs = stp.frame(x)
        print("  frame %d iters %d conv %d halv %d evals %d | ms total %.2f solve %.2f refresh %.2f | wall %.1f ms | launches %d" % (f + 1, fs.iters, fs.converged, fs.halvings, fs.energy_evals, fs.ms_total, fs.ms_solve, fs.ms_refresh, 1e3 * (time.time() - t0), stp.launch_count()))
    names = ["energy", "gradient", "elem_hessians", "fill", "factorize", "precondition", "dot"]
    print("  kernel ms:", {n: round(stp.time_kernels(i, 10), 4) for i, n in enumerate(names)})

if __name__ == "__main__":
    print("devices", D.device_count())
    section("energy tiny_snh", lambda: energy_case("tiny_snh_k4_twist"))
    section("energy tiny_fcr_inverted", lambda: energy_case("tiny_fcr_inverted"))
    section("solver tiny sub0", lambda: solver_case("tiny_snh_k4_twist", 0))
    section("solver small sub1", lambda: solver_case("small_snh_k4_twist", 1))
    section("solver small global", lambda: solver_case("small_snh_k4_twist", -1))
    section("stepper tiny_snh", lambda: stepper_case("tiny_snh_k4_twist"))
    section("stepper small_fcr", lambda: stepper_case("small_fcr_k3_stretch"))
    if "--big" in sys.argv:
        section("bar17K_like SNH k8", lambda: big_case("bar17K_like", 8, "SNH"))
