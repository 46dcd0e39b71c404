"""World-size-2 tests of the multi-GPU host logic on CPU (torch.distributed, gloo backend): the static subdomain -> rank map
of the C library and the reduction semantics of the sharded preconditioner (DOTTimeStepper.cpp:406-450 split over ranks):
  p = D^-1 all_reduce_sum_over_ranks( sum_{s owned by the rank} R_s^T H_s^-1 R_s q ).
The per-subdomain solves are done by the oracle here (scipy) - the point is the sharding/reduction logic, which is what
differs from the single-GPU path; the GPU solves themselves are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dot_b200 as D
from golden_util import Golden
from oracle import dot_oracle as O


def test_ownership_is_a_partition():
    for k in (1, 3, 8, 64, 128):
        for world in (1, 2, 4, 8):
            seen = np.concatenate([D.owned_subdomains(k, r, world) for r in range(world)])
            assert sorted(seen.tolist()) == list(range(k))
            sizes = [len(D.owned_subdomains(k, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(D.DotGpuError):
        D.owned_subdomains(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = Golden(name)
        st = g.states()[-1]
        k = g.k
        nV = g["setup/V_rest"].shape[0]
        q = -g[st + "/g"]
        part = np.zeros(3 * nV)
        for s in D.owned_subdomains(k, rank, world):
            pre = "sbd%d_" % s
            ia, ja, a = g["setup/" + pre + "ia"], g["setup/" + pre + "ja"], g[st + "/" + pre + "a"]
            A = O.csr_upper_to_full(ia, ja, a)
            l2g = g["setup/" + pre + "l2g"]
            dof = (3 * l2g[:, None] + np.arange(3)[None, :]).ravel()
            import scipy.sparse.linalg as spla
            part[dof] += spla.spsolve(A.tocsc(), q[dof])     # ascending subdomain order inside a rank
        t = torch.from_numpy(part)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)             # what ncclAllReduce does on the GPUs
        dup = np.repeat(g["setup/dup"].astype(float), 3)
        p = t.numpy() / np.maximum(dup, 1.0)
        np.save(os.path.join(out_dir, "p_rank%d.npy" % rank), p)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tiny_snh_k4_twist", "small_snh_k5_tsns_dt24"])
def test_sharded_preconditioner_world2_gloo(tmp_path, name):
    g = Golden(name)
    st = g.states()[-1]
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    p0, p1 = (np.load(tmp_path / ("p_rank%d.npy" % r)) for r in range(world))
    assert np.array_equal(p0, p1)                            # every rank ends with the same replica
    ref = g[st + "/p"]                                       # the reference's own p (CHOLMOD solves, serial scatter-add)
    assert np.linalg.norm(p0 - ref) <= 1e-9 * np.linalg.norm(ref)


def test_bench_reference_arm_rank1_is_silent():
    """Under torchrun only rank 0 runs / prints the reference arm; other ranks exit 0 without output."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_balanced_owner_is_a_deterministic_balanced_partition():
    """subdomains go to ranks balancing sum nnz(L_s).  The map is a partition, deterministic, within the LPT
    bound (max load <= 4/3 OPT, OPT >= max(mean, heaviest)) and never worse than round-robin on sorted weights."""
    rng = np.random.default_rng(5)
    for k, world in ((8, 2), (16, 8), (64, 8), (128, 8), (128, 4), (5, 8)):
        w = rng.uniform(0.7, 1.5, k) * 1e6
        own = D.balanced_owner(w, world)
        assert own.shape == (k,) and own.min() >= 0 and own.max() < world
        assert np.array_equal(own, D.balanced_owner(w.copy(), world))
        load = np.bincount(own, weights=w, minlength=world)
        opt_lb = max(w.sum() / world, w.max())
        assert load.max() <= 4.0 / 3.0 * opt_lb + 1e-9
        rr = np.bincount(np.arange(k) % world, weights=np.sort(w)[::-1], minlength=world)
        assert load.max() <= rr.max() + 1e-9
    # equal weights: ties go to the lower rank / lower subdomain id -> exactly the round-robin map
    assert np.array_equal(D.balanced_owner(np.ones(12), 4), np.arange(12) % 4)


def _worker_grad(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = Golden(name)
        st = g.states()[-1]
        m = O.Mesh(g["setup/V_rest"], g["setup/F"])
        en, dt = g.meta["energy"], g.meta["dt"]
        x, xt = g[st + "/V"], g[st + "/xTilta"]
        fm = np.zeros(m.nV, dtype=bool)
        fm[g[st + "/fixed"]] = True
        ep = g["setup/epart"]
        own = D.balanced_owner(np.bincount(ep, minlength=g.k).astype(float), world)   # any weights: the logic under test is the sharding
        mine = own[ep] == rank
        # what Stepper::eval_sharded does on rank `rank`: energy + gradient of the OWNED tets, inertia terms on rank 0 only
        F = O.deformation_gradient(m, x)
        U, s, V = O.svd_rot(F)
        e_t = O.elastic_energy_per_elem(en, m, s)
        ge = O.elem_gradient(m, O.first_piola(en, U, s, V, m.mu, m.lam), dt * dt)
        ge[~mine] = 0.0
        part = O.gather_gradient(m, ge, fm).reshape(-1, 3)
        E = dt * dt * e_t[mine].sum()
        if rank == 0:
            part[~fm] += m.mass[~fm, None] * (x[~fm] - xt[~fm])
            E += (((x - xt) ** 2).sum(axis=1) * m.mass / 2.0).sum()
        t = torch.from_numpy(np.concatenate([part.reshape(-1), [E]]))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)             # the one all-reduce of [g ; E] per evaluation
        np.save(os.path.join(out_dir, "gE_rank%d.npy" % rank), t.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tiny_snh_k4_twist", "small_fcr_k3_stretch"])
def test_tet_sharded_energy_gradient_world2_gloo(tmp_path, name):
    """SURVEY 8(e) collectives (2),(3): every tet is owned by exactly one rank (the element partition is disjoint), the inertia
    terms by rank 0; one all-reduce of [g ; E] gives every rank the reference's gradient and energy."""
    g = Golden(name)
    st = g.states()[-1]
    world = 2
    mp.spawn(_worker_grad, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / ("gE_rank%d.npy" % r)) for r in range(world))
    assert np.array_equal(r0, r1)
    assert np.linalg.norm(r0[:-1] - g[st + "/g"]) <= 2e-8 * np.linalg.norm(g[st + "/g"])
    assert abs(r0[-1] - g[st + "/E"][0]) <= 1e-10 * abs(g[st + "/E"][0])
