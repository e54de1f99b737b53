"""Repeated small calls through every entry point: device memory in use must not grow (scratch is cached,
temporaries are stream-ordered, per-call tables are freed)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gala_b200 as gb
from gala_b200.mockstream import DirectNBody

pots = [gb.MilkyWayPotential2022(), gb.BovyMWPotential2014(), gb.LM10Potential(),
        gb.PlummerPotential(m=1e10, b=1.0) + gb.NFWPotential(m=6e11, r_s=16.0, c=0.8)]
rng = np.random.default_rng(0)
w0 = np.vstack([rng.normal(0, 10, (3, 512)), rng.normal(0, 0.1, (3, 512))])
t = np.arange(21.0)
prog = np.array([13.0, 0.0, 20.0, 0.0, 0.13, 0.05])


def used():
    torch.cuda.synchronize()
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2**20


def one(k):
    pot = pots[k % len(pots)]
    H = gb.Hamiltonian(pot)
    gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=k % 2)
    gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=0)
    gb.dop853_integrate_hamiltonian(H, w0[:, :128], t, save_all=k % 2)
    pot.gradient(w0[:3]); pot.energy(w0[:3]); pot.hessian(w0[:3, :64])
    nb = DirectNBody(prog, [gb.PlummerPotential(m=1e8, b=0.1)], external_potential=pot)
    nb.integrate_orbit(t=t, Integrator="leapfrog" if k % 2 else "dopri853")
    if k % 10 == 0:
        gen = gb.MockStreamGenerator(gb.FardalStreamDF(random_state=np.random.RandomState(1)), H,
                                     progenitor_potential=gb.PlummerPotential(m=1e8, b=0.1) if k % 20 == 0 else None)
        gen.run(prog, 1e8, dt=1.0, n_steps=20, n_particles=1, Integrator="leapfrog" if k % 3 else "dopri853")
        gb.fast_lyapunov_max(w0[:, :4], H, dt=1.0, n_steps=40, return_orbit=False)


for k in range(40):
    one(k)
base = used()
t0 = time.perf_counter()
n = 600
for k in range(n):
    one(k)
    if k % 200 == 199:
        print(f"after {k + 1} rounds: {used():.1f} MiB in use (baseline {base:.1f})")
print(f"{n} rounds in {time.perf_counter() - t0:.1f} s; growth {used() - base:+.1f} MiB")
assert used() - base < 64.0
