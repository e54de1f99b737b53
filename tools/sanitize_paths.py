#!/usr/bin/env python
"""One tiny invocation of every kernel family through the public API, for `compute-sanitizer` (memcheck /
racecheck / synccheck): sizes are small because the sanitizer slows FP64 kernels by one to two orders of magnitude.
No oracle here -- results are only checked for finiteness; parity is the job of tests/.
usage: compute-sanitizer --tool memcheck python tools/sanitize_paths.py [family ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gala_b200 as gb                                     # noqa: E402
from gala_b200.mockstream import DirectNBody              # noqa: E402

KMS = gb.KMS_TO_KPC_MYR


def ics(pot, n, seed=0, rmin=6.0, rmax=30.0):
    rng = np.random.default_rng(seed)
    r = np.exp(rng.uniform(np.log(rmin), np.log(rmax), n))
    mu = rng.uniform(-1, 1, n); ph = rng.uniform(0, 2 * np.pi, n); s_ = np.sqrt(1 - mu * mu)
    rhat = np.vstack([s_ * np.cos(ph), s_ * np.sin(ph), mu])
    e1 = np.vstack([-np.sin(ph), np.cos(ph), np.zeros(n)])
    e2 = np.cross(rhat.T, e1.T).T
    psi = rng.uniform(0, 2 * np.pi, n)
    q = np.ascontiguousarray(r * rhat)
    g = pot.gradient(q)
    vc = np.sqrt(r * np.sqrt((g * g).sum(0)))
    v = rng.uniform(0.6, 1.0, n) * vc * (np.cos(psi) * e1 + np.sin(psi) * e2)
    return np.ascontiguousarray(np.vstack([q, v]))


def scf(nmax=3, lmax=2, seed=1):
    rng = np.random.default_rng(seed)
    S = rng.normal(0, 0.01, (nmax + 1, lmax + 1, lmax + 1)); T = rng.normal(0, 0.01, S.shape)
    S[0, 0, 0] = 1.0
    return gb.SCFPotential(m=1e12, r_s=20.0, Snlm=S, Tnlm=T)


def multipole(lmax=3, seed=2):
    rng = np.random.default_rng(seed)
    kw = {}
    for l in range(lmax + 1):
        for m in range(l + 1):
            kw[f"S{l}{m}"] = rng.normal()
            if m > 0:
                kw[f"T{l}{m}"] = rng.normal()
    return gb.MultipolePotential(lmax=lmax, inner=False, m=2e10, r_s=8.0, **kw)


def finite(*arrs):
    for a in arrs:
        assert np.isfinite(np.asarray(a)).all()


def fam_eval():
    mw = gb.MilkyWayPotential2022()
    q = ics(mw, 333)[:3].copy()
    for pot in (mw, gb.LM10Potential(), gb.BovyMWPotential2014(), scf(), multipole(),
                gb.MilkyWayPotential2022() + gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=0.4)):
        for strict in (True, False):
            pot.strict_math = strict
            finite(pot.gradient(q), pot.energy(q), pot.density(q))
            if not isinstance(pot, (gb.SCFPotential, gb.MultipolePotential)):     # no Hessian for expansions, as in the reference
                finite(pot.hessian(q))
    H = gb.Hamiltonian(mw, gb.ConstantRotatingFrame([0.0, 0.01, 0.03]))
    finite(H.energy(ics(mw, 77)))


def fam_fixed():
    mw = gb.MilkyWayPotential2022()
    bar = gb.MilkyWayPotential2022() + gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=0.4)
    t = np.arange(41.0)
    for n in (1, 31, 517):                         # below one warp, ragged, several CTAs
        w0 = ics(mw, n)
        for strict in (True, False):
            mw.strict_math = strict
            H = gb.Hamiltonian(mw)
            for sa in (0, 1):
                finite(gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=sa)[1])
                finite(gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=sa)[1])
        Hr = gb.Hamiltonian(bar, gb.ConstantRotatingFrame([0.0, 0.0, 0.030681]))
        finite(gb.ruth4_integrate_hamiltonian(Hr, w0, t, save_all=1, allow_rotating_frame=True)[1])
    for pot in (scf(), multipole(), gb.LM10Potential()):
        H = gb.Hamiltonian(pot)
        finite(gb.leapfrog_integrate_hamiltonian(H, ics(mw, 100), t, save_all=0)[1])


def fam_dop853():
    mw = gb.MilkyWayPotential2022()
    t = np.linspace(0.0, 600.0, 61)
    for n in (1, 31, 700):
        w0 = ics(mw, n)
        for strict in (True, False):
            mw.strict_math = strict
            for frame in (gb.StaticFrame(), gb.ConstantRotatingFrame([0.0, 0.0, 0.030681])):
                H = gb.Hamiltonian(mw, frame)
                finite(gb.dop853_integrate_hamiltonian(H, w0, t)[1])
                finite(gb.dop853_integrate_hamiltonian(H, w0, t, save_all=0)[1])
    # backward grid, irregular grid, and a failing run (nmax too small: NaN rows + status, no exception asked)
    w0 = ics(mw, 65)
    H = gb.Hamiltonian(mw)
    finite(gb.dop853_integrate_hamiltonian(H, w0, -t)[1])
    finite(gb.dop853_integrate_hamiltonian(H, w0, np.sort(np.random.default_rng(3).uniform(0, 500, 40)))[1])
    gb.dop853_integrate_hamiltonian(H, w0, t, nmax=3, err_if_fail=0)
    finite(gb.dop853_integrate_hamiltonian(gb.Hamiltonian(scf()), w0, t)[1])


def fam_extrema():
    mw = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(mw)
    w0 = ics(mw, 200)
    t = np.arange(301.0) * 2.0
    for integ in ("leapfrog", "ruth4"):
        for we in (False, True):
            gb.integrate_extrema(H, w0, t, Integrator=integ, with_energy=we)
    _, w = gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=1)
    gb.orbit_extrema(H, w, t, with_energy=True)
    for kind in ("peri", "apo", "zmax"):
        gb.orbit_extrema_list(w, t, kind=kind)


def fam_timeinterp():
    T = np.linspace(0.0, 400.0, 9)
    grow = 1.0 + 0.2 * T / 400.0
    ang = 0.04 * T
    R = np.array([[[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]] for a in ang])
    t = np.linspace(0.0, 400.0, 81)
    mw = gb.MilkyWayPotential2022()
    w0 = ics(mw, 150)
    for method in ("linear", "cspline", "akima", "steffen"):
        comp = gb.CCompositePotential()
        comp["halo"] = gb.NFWPotential(m=6e11, r_s=16.0)
        comp["disk"] = gb.MiyamotoNagaiPotential(m=6e10, a=3.0, b=0.3)
        comp["bar"] = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, T, interpolation_method=method,
                                                   m=1e10 * grow, a=4.0, b=0.8, c=0.25, R=R)
        H = gb.Hamiltonian(comp)
        finite(comp.gradient(w0[:3].copy(), t=100.0), comp.energy(w0[:3].copy(), t=100.0))
        finite(gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0)[1])
        finite(gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=1)[1])
        finite(gb.dop853_integrate_hamiltonian(H, w0, t[::8].copy())[1])
        finite(gb.PotentialBase.hessian(comp, w0[:3].copy(), 100.0))
    # a stream in a galaxy with an infalling, growing satellite (release + particles at their own times)
    Tk = np.linspace(-200.0, 10.0, 43)
    track = np.stack([60.0 + 0.2 * Tk, -20.0 - 0.15 * Tk, 10.0 + 0.05 * Tk], axis=1)
    gal = gb.CCompositePotential()
    gal["halo"] = gb.NFWPotential(m=6e11, r_s=16.0)
    gal["lmc"] = gb.TimeInterpolatedPotential(gb.HernquistPotential, Tk, m=1.5e11 * np.linspace(0.6, 1.0, Tk.size), c=10.0, origin=track)
    prog = gb.PhaseSpacePosition(pos=[13.0, 0.0, 20.0], vel=[0.0, 130.0 * KMS, 50.0 * KMS])
    for integ in ("dopri853", "leapfrog"):
        gen = gb.MockStreamGenerator(gb.FardalStreamDF(random_state=np.random.default_rng(1)), gb.Hamiltonian(gal))
        stream, p = gen.run(prog, 2.5e4, dt=-1.0, n_steps=100, n_particles=2, Integrator=integ)
        finite(stream.pos, stream.vel, p.pos)


def fam_nbody():
    mw = gb.MilkyWayPotential2022()
    rng = np.random.default_rng(3)
    b0 = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * KMS, 50.0 * KMS])
    b1 = np.array([-25.0, 8.0, 5.0, 0.02, -0.15, 0.03])
    tp = b0[None, :] + np.hstack([rng.normal(0, 1.5, (70, 3)), rng.normal(0, 0.004, (70, 3))])
    pps = [gb.HernquistPotential(m=2e9, c=0.8), gb.PlummerPotential(m=5e9, b=1.2)] + [None] * len(tp)
    nb = DirectNBody(gb.PhaseSpacePosition.from_w(np.ascontiguousarray(np.vstack([b0, b1, tp]).T)), pps, external_potential=mw)
    t = np.arange(0, 61.0)
    for integ in ("leapfrog", "ruth4", "dopri853"):
        for sa in (True, False):
            nb.save_all = sa
            o = nb.integrate_orbit(t=t, Integrator=integ)
            finite(o.pos, o.vel)


def fam_mockstream():
    mw = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(mw)
    prog = gb.PhaseSpacePosition(pos=[13.0, 0.0, 20.0], vel=[0.0, 130.0 * KMS, 50.0 * KMS])
    for df in (gb.FardalStreamDF(random_state=np.random.default_rng(1)),
               gb.ChenStreamDF(random_state=np.random.default_rng(1)),
               gb.StreaklineStreamDF(random_state=np.random.default_rng(1)),
               gb.LagrangeCloudStreamDF(0.001, random_state=np.random.default_rng(1))):
        for integ in ("dopri853", "leapfrog"):
            gen = gb.MockStreamGenerator(df, H)
            stream, p = gen.run(prog, 2.5e4, dt=-1.0, n_steps=150, n_particles=2, Integrator=integ)
            finite(stream.pos, stream.vel, p.pos)
    gen = gb.MockStreamGenerator(gb.FardalStreamDF(random_state=np.random.default_rng(2)), H,
                                 progenitor_potential=gb.PlummerPotential(m=2.5e4, b=0.004))
    for integ in ("dopri853", "leapfrog"):
        stream, p = gen.run(prog, 2.5e4, dt=-1.0, n_steps=120, n_particles=1, Integrator=integ)
        finite(stream.pos, p.pos)


def fam_lyapunov():
    pot = gb.LM10Potential()
    w0 = ics(pot, 12, seed=9)
    for frame in (gb.StaticFrame(), gb.ConstantRotatingFrame([0.0, 0.0, 0.02])):
        LEs = gb.fast_lyapunov_max(w0, gb.Hamiltonian(pot, frame), dt=2.0, n_steps=100, noffset_orbits=2, return_orbit=False)
        finite(LEs)


FAMILIES = {k[4:]: v for k, v in list(globals().items()) if k.startswith("fam_")}

if __name__ == "__main__":
    which = sys.argv[1:] or list(FAMILIES)
    n0 = gb._abi.launch_count()
    for name in which:
        t0 = time.time()
        FAMILIES[name]()
        print(f"[sanitize_paths] {name}: ok, {time.time() - t0:.1f} s, kernel launches so far {gb._abi.launch_count() - n0}", flush=True)
