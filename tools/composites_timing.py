"""Leapfrog throughput (3.03e6 orbits x 1000 steps, device-resident) for the named composites: compile-time
signatures vs the analytic-only generic loop."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gala_b200 as gb
from bench import make_ic

cases = [("mw2022", gb.MilkyWayPotential2022()), ("mw_v1", gb.MilkyWayPotential()), ("lm10", gb.LM10Potential()),
         ("bovy2014", gb.BovyMWPotential2014()),
         ("plummer+nfw (light generic)", gb.PlummerPotential(m=1e10, b=1.0) + gb.NFWPotential(m=6e11, r_s=16.0)),
         ("nfw + time-interpolated plummer", gb.NFWPotential(m=6e11, r_s=16.0) + gb.TimeInterpolatedPotential(
             gb.PlummerPotential, np.linspace(0.0, 1000.0, 41), m=1e10 * np.linspace(1.0, 2.0, 41), b=1.0,
             origin=np.stack([8 * np.cos(np.linspace(0, 6, 41)), 8 * np.sin(np.linspace(0, 6, 41)), np.zeros(41)], axis=1)))]
for name, pot in cases:
    H = gb.Hamiltonian(pot)
    w0 = torch.as_tensor(make_ic(3031040, 1, lambda q: pot.gradient(q)), device="cuda")
    t = np.arange(1001.0)
    for _ in range(2):
        gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0)
    torch.cuda.synchronize()
    ev, wall = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0); e1.record()
        torch.cuda.synchronize(); wall.append((time.perf_counter() - t0) * 1e3); ev.append(e0.elapsed_time(e1))
    el = float(np.median(ev)) * 1e-3
    print(f"{name:32s} leapfrog 3.03e6 x 1000: {el * 1e3:7.2f} ms (CUDA events, median of 5; min {min(ev):.2f} max {max(ev):.2f}; "
          f"host wall median {np.median(wall):.2f})  {3031040 * 1000 / el:.3e} orbit-steps/s", flush=True)
