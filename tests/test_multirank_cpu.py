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
