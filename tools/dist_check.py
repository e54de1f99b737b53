"""2+ GPUs, launched with torchrun: every rank integrates its contiguous slice of the orbit index on its own
GPU (leapfrog final state and DOP853 dense, device buffers), the results are gathered over NCCL, and rank 0
checks them bit for bit against the unsharded run on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import gala_b200 as gb
from gala_b200.dist import integrate_sharded
from bench import make_ic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pot = gb.MilkyWayPotential2022(); H = gb.Hamiltonian(pot)
N = 200_003
w0 = make_ic(N, 7, lambda q: pot.gradient(q))
w0d = torch.as_tensor(w0, device="cuda")
t = np.arange(501.0)
_, w = integrate_sharded(gb.leapfrog_integrate_hamiltonian, H, w0d, t, save_all=0)
td = np.linspace(0, 300, 31)
_, wd = integrate_sharded(gb.dop853_integrate_hamiltonian, H, w0d[:, :20_001].contiguous(), td, save_all=1)
if rank == 0:
    _, w1 = gb.leapfrog_integrate_hamiltonian(H, w0d, t, save_all=0)
    _, wd1 = gb.dop853_integrate_hamiltonian(H, w0d[:, :20_001].contiguous(), td, save_all=1)
    print(f"world={world}: leapfrog sharded == unsharded: {torch.equal(w, w1)}; shape {tuple(w.shape)}")
    print(f"world={world}: dop853 dense sharded == unsharded: {torch.equal(wd, wd1)}; shape {tuple(wd.shape)}")
dist.barrier()
dist.destroy_process_group()
