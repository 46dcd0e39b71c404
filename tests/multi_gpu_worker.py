"""Worker of tests/test_gpu_parity.py::test_two_ranks_match_one_rank (launched by torch.distributed.run, one rank per GPU):
runs a few DOT time steps of a golden case with the subdomains and tets sharded over the ranks (NCCL) and saves the positions."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dot_b200 as D  # noqa: E402
from golden_util import Golden  # noqa: E402


def main():
    name, frames, tol, out = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(D.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    g = Golden(name)
    V, T = g["setup/V_rest"], g["setup/F"]
    a = D.Anim(g.meta["anim"], V)
    stp = D.Stepper(V, T, g["setup/epart"], a.fixed_mask(), energy=g.meta["energy"], k=g.k, dt=g.meta["dt"], rel_tol=tol, device=lr,
                    rank=rank, world=world, nccl_id=bytes(buf.cpu().numpy().tobytes()))
    owned = stp.owned()
    x = V.copy()
    iters = halv = 0
    for f in range(frames):
        a.step(x, g.meta["dt"])
        fs = stp.frame(x)
        assert fs.converged == 1, (rank, f)
        iters += fs.iters
        halv += fs.halvings
    np.savez(out + ".rank%d.npz" % rank, x=x, iters=iters, halvings=halv, owned=np.asarray(owned))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
