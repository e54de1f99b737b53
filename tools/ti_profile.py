"""One device-resident leapfrog call of `NFW + TimeInterpolated Plummer with a moving origin` (the case of
tools/composites_timing.py), for ncu: usage  ncu ... python tools/ti_profile.py [n_orbits] [n_steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gala_b200 as gb
from bench import make_ic

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3031040
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
k = np.linspace(0, 6, 41)
pot = gb.NFWPotential(m=6e11, r_s=16.0) + gb.TimeInterpolatedPotential(
    gb.PlummerPotential, np.linspace(0.0, 1000.0, 41), m=1e10 * np.linspace(1.0, 2.0, 41), b=1.0,
    origin=np.stack([8 * np.cos(k), 8 * np.sin(k), np.zeros(41)], axis=1))
H = gb.Hamiltonian(pot)
w0 = torch.as_tensor(make_ic(N, 1, lambda q: pot.gradient(q)), device="cuda")
t = np.arange(nsteps + 1.0)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0); e1.record()
    torch.cuda.synchronize()
    print(f"leapfrog {N} x {nsteps}: {e0.elapsed_time(e1):.2f} ms", flush=True)
