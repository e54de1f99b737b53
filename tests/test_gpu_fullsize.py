"""BASELINE.json's FULL problem sizes on one B200, checked through size-independent properties (the oracle
cannot run 10^6-10^7 orbits in seconds): the initial row of a trajectory is the input, energy is conserved
the way the integrator conserves it, a leapfrog run reversed in time returns to its start, a sharded run
equals the unsharded one bit for bit, and a random SAMPLE of the orbits equals the compiled reference.
Everything big stays on the device (torch tensors are only buffers here)."""
import numpy as np
import pytest

import gala_b200 as gb
from conftest import make_ic, relnorm

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def _sample(N, k, seed=0):
    return np.sort(np.random.default_rng(seed).choice(N, k, replace=False))


def test_c2_full_million_orbits_dop853_dense(ref):
    """C2: 10^6 orbits, MW2022, DOP853 rtol=atol=1e-10, 1000 dense-output times (48 GB on the device)."""
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    N = 1_000_000
    w0 = make_ic(lambda q: pot.gradient(q), N, seed=2)
    t = np.linspace(0, 1000, 1000)
    w0d = _dev(w0)
    tt, w, st = gb.dop853_integrate_hamiltonian(H, w0d, t, save_all=1, return_status=True)
    assert tuple(w.shape) == (6, 1000, N)
    assert bool((st["status"] == 1).all())
    assert torch.equal(w[:, 0, :], w0d)                               # the first row is the input
    assert bool(torch.isfinite(w[:, -1, :]).all())
    # energy along every trajectory, evaluated on the device at 4 of the 1000 output times
    E0 = H.energy(w0d)
    for j in (1, 333, 666, 999):
        Ej = H.energy(w[:, j, :].contiguous())
        drift = ((Ej - E0) / E0).abs()
        assert float(drift.max()) < 1e-6 and float(drift.median()) < 1e-9, (j, float(drift.max()))
    # a random sample against the reference (nbatch=1 == per-orbit step control)
    idx = _sample(N, 300, seed=2)
    wr, sr, rc = ref.dop853(H, w0[:, idx], t, nbatch=1)
    assert rc >= 0
    got = w[:, :, torch.as_tensor(idx, device="cuda")].cpu().numpy()
    d = relnorm(got[:, -1], wr[:, -1])
    print(f"\n[C2 full] sample of 300 vs reference at t_end: q50/q99/max = {np.median(d):.2e} {np.quantile(d, .99):.2e} {d.max():.2e}")
    assert np.median(d) < 1e-11 and np.quantile(d, 0.9) < 1e-9
    mean_steps = float(st["nstep"].double().mean())
    assert 100 < mean_steps < 2000
    del w
    gb._abi.release_scratch()
    torch.cuda.empty_cache()


def test_c4_full_million_orbits_ruth4_rotating(ref):
    """C4: 10^6 orbits, LongMuraliBar + MW2022, ConstantRotatingFrame, Ruth4 dt=0.5 Myr x 1000."""
    pot = gb.CCompositePotential()
    pot["bar"] = gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=np.deg2rad(25.0))
    for k, v in gb.MilkyWayPotential2022().items():
        pot[k] = v
    H = gb.Hamiltonian(pot, gb.ConstantRotatingFrame([0.0, 0.0, 0.030681]))
    N = 1_000_000
    w0 = make_ic(lambda q: pot.gradient(q), N, seed=4)
    t = np.arange(1001) * 0.5
    w0d = _dev(w0)
    _, w = gb.ruth4_integrate_hamiltonian(H, w0d, t, save_all=0, allow_rotating_frame=True)
    assert tuple(w.shape) == (6, N) and bool(torch.isfinite(w).all())
    # sharding invariance: two halves integrated separately give the same bits
    h = N // 2
    _, wa = gb.ruth4_integrate_hamiltonian(H, w0d[:, :h].contiguous(), t, save_all=0, allow_rotating_frame=True)
    _, wb = gb.ruth4_integrate_hamiltonian(H, w0d[:, h:].contiguous(), t, save_all=0, allow_rotating_frame=True)
    assert torch.equal(torch.cat([wa, wb], dim=1), w)
    idx = _sample(N, 400, seed=4)
    wr = ref.ruth4(H, w0[:, idx], t, save_all=False)
    d = relnorm(w[:, torch.as_tensor(idx, device="cuda")].cpu().numpy(), wr)
    print(f"\n[C4 full] sample of 400 vs reference: q50/q99/max = {np.median(d):.2e} {np.quantile(d, .99):.2e} {d.max():.2e}")
    assert np.median(d) < 1e-12 and np.quantile(d, 0.9) < 1e-10


def test_c5_full_ten_million_orbits_scf(port_lib):
    """C5: 10^7 orbits in SCF(nmax=10, lmax=6), leapfrog dt=1 x 1000, final state only."""
    from test_gpu_parity import scf_c5
    pot = scf_c5()
    H = gb.Hamiltonian(pot)
    N = 10_000_000
    w0 = make_ic(lambda q: pot.gradient(q), N, seed=5)
    t = np.arange(1001, dtype=float)
    w0d = _dev(w0)
    _, w = gb.leapfrog_integrate_hamiltonian(H, w0d, t, save_all=0)
    assert tuple(w.shape) == (6, N) and bool(torch.isfinite(w).all())
    E0, E1 = H.energy(w0d), H.energy(w)
    drift = ((E1 - E0) / E0).abs()
    # leapfrog at dt = 1 Myr is coarse for the innermost orbits (r ~ 4 kpc): bounded, not tiny
    assert float(drift.median()) < 1e-4 and float(drift.quantile(0.999)) < 5e-2
    # time reversal: the leapfrog scheme retraces its steps
    wrev = w.clone(); wrev[3:] *= -1
    _, wback = gb.leapfrog_integrate_hamiltonian(H, wrev, t, save_all=0)
    wback[3:] *= -1
    back = (wback[:3] - w0d[:3]).norm(dim=0) / w0d[:3].norm(dim=0)
    assert float(back.median()) < 1e-10
    idx = _sample(N, 64, seed=5)
    wr = port_lib.leapfrog(pot, w0[:, idx], t, save_all=False)
    d = relnorm(w[:, torch.as_tensor(idx, device="cuda")].cpu().numpy(), wr)
    print(f"\n[C5 full] sample of 64 vs the port: q50/max = {np.median(d):.2e} {d.max():.2e}; "
          f"time-reversal median {float(back.median()):.2e}")
    assert np.median(d) < 1e-10


def test_headline_ten_thousand_steps_reversibility_and_energy(ref):
    """Headline potential at the parity length of the north star (10^4 leapfrog steps), 10^6 orbits."""
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    N = 1_000_000
    w0 = make_ic(lambda q: pot.gradient(q), N, seed=7, rmin=8.0)
    t = np.arange(10001, dtype=float)
    w0d = _dev(w0)
    _, w = gb.leapfrog_integrate_hamiltonian(H, w0d, t, save_all=0)
    E0, E1 = H.energy(w0d), H.energy(w)
    drift = ((E1 - E0) / E0).abs()
    assert float(drift.median()) < 1e-4
    wrev = w.clone(); wrev[3:] *= -1
    _, wback = gb.leapfrog_integrate_hamiltonian(H, wrev, t, save_all=0)
    back = (wback[:3] - w0d[:3]).norm(dim=0) / w0d[:3].norm(dim=0)
    print(f"\n[headline 1e4 steps] energy drift median {float(drift.median()):.2e}; "
          f"forward+backward returns to the start: median {float(back.median()):.2e}, q99 {float(back.quantile(0.99)):.2e}")
    assert float(back.median()) < 1e-10
    idx = _sample(N, 300, seed=7)
    wr = ref.leapfrog(pot, w0[:, idx], t, save_all=False)
    d = relnorm(w[:, torch.as_tensor(idx, device="cuda")].cpu().numpy(), wr)
    print(f"[headline 1e4 steps] sample of 300 vs reference: q50/q90/max = {np.median(d):.2e} {np.quantile(d, .9):.2e} {d.max():.2e}")
    assert np.median(d) < 1e-11


def test_c3_full_stream_hundred_thousand_particles(ref):
    """C3: 100 020-particle Fardal stream in MW2022 (dt=-1 Myr, 5000 steps, 10 particles per tail per step)."""
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    prog = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * gb.KMS_TO_KPC_MYR, 50.0 * gb.KMS_TO_KPC_MYR])
    out = {}
    for integ in ("leapfrog", "dopri853"):
        gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)), H)
        stream, p = gen.run(prog, 2.5e4, dt=-1.0, n_steps=5000, n_particles=10, release_every=1, Integrator=integ)
        assert stream.pos.shape == (3, 100020) and np.all(np.isfinite(stream.pos))
        # integrating back 5000 Myr and forward again: the progenitor ends where it started
        assert np.allclose(p.pos.ravel(), prog[:3], atol=1e-5) and np.allclose(p.vel.ravel(), prog[3:], atol=1e-7)
        # the particles released at the very end have not moved; the stream is long and thin
        last = np.asarray(stream.release_time) == 0.0
        assert last.sum() == 20
        d = np.sqrt(((stream.pos - p.pos.reshape(3, 1)) ** 2).sum(0))
        assert d[last].max() < 0.5 and d.max() > 10.0
        out[integ] = stream.pos
    # both integrators draw the same particles and agree on where the stream is
    sep = np.sqrt(((out["leapfrog"] - out["dopri853"]) ** 2).sum(0))
    assert np.median(sep) < 1e-2
