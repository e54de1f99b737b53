"""GPU parity of the mock-stream path (config C3 at reduced size) against the compiled reference +
the numpy restatement of the Python-level DF code (oracle/oracle.py)."""
import numpy as np
import pytest

import gala_b200 as gb
from gala_b200.mockstream import DirectNBody
from conftest import relnorm

pytestmark = pytest.mark.gpu

KMS = gb.KMS_TO_KPC_MYR
PROG_W0 = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * KMS, 50.0 * KMS])   # tests/dynamics/mockstream/test_mockstream.py:676-678


def _prog_orbit(H, n_steps=200, dt=-1.0):
    t = gb.parse_time_specification(None, dt=dt, n_steps=n_steps)
    nb = DirectNBody(PROG_W0, [None], external_potential=H.potential, frame=H.frame)
    return nb.integrate_orbit(t=t, Integrator="dopri853"), t


@pytest.mark.parametrize("rng_kind", ["RandomState", "Generator"])
@pytest.mark.parametrize("gala_modified", [True, False])
def test_fardal_release_parity(ref, rng_kind, gala_modified):
    from oracle import oracle
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    orb, t = _prog_orbit(H, 64)
    prog = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
    mk = (lambda: np.random.RandomState(42)) if rng_kind == "RandomState" else (lambda: np.random.default_rng(42))
    npart = np.zeros(65, dtype="i4"); npart[::2] = 3
    prog_m = np.full(65, 2.5e4); prog_m[10] = 0.0
    df = gb.FardalStreamDF(gala_modified=gala_modified, random_state=mk())
    pot.strict_math = True
    s = df.sample(prog, prog_m, n_particles=npart)
    x0, v0, t10 = oracle.fardal_release_numpy(ref, pot, prog.pos.T, prog.vel.T, prog.t, prog_m, npart, mk(),
                                              gala_modified=gala_modified)
    assert s.pos.shape == (3, x0.shape[0])
    assert np.array_equal(s.release_time, t10)
    # rj = (...)**(1/3) goes through libm pow on both sides: agreement to a few ulp of the offsets
    off = np.sqrt(((x0 - prog.pos.T[np.searchsorted(prog.t, t10)]) ** 2).sum(1))
    assert np.max(np.abs(s.pos.T - x0) / off[:, None]) < 1e-12
    assert np.max(np.abs(s.vel.T - v0)) / np.abs(v0).max() < 1e-13
    assert list(s.lead_trail[:6]) == ["t", "t", "t", "l", "l", "l"]


@pytest.mark.parametrize("kind", ["streakline", "lagrange", "chen"])
def test_other_stream_dfs_release_parity(ref, kind):
    """StreaklineStreamDF / LagrangeCloudStreamDF / ChenStreamDF (df.pyx:242-318, 460-552, 556-702) against the
    numpy-scalar restatement drawing the RNG call by call like the reference."""
    from oracle import oracle
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    orb, t = _prog_orbit(H, 48)
    prog = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
    npart = np.zeros(49, dtype="i4"); npart[::3] = 4
    prog_m = np.full(49, 2.5e4)
    for mk in (lambda: np.random.RandomState(5), lambda: np.random.default_rng(5)):
        df = {"streakline": lambda: gb.StreaklineStreamDF(random_state=mk()),
              "lagrange": lambda: gb.LagrangeCloudStreamDF(v_disp=1.2e-3, random_state=mk()),
              "chen": lambda: gb.ChenStreamDF(random_state=mk())}[kind]()
        pot.strict_math = True
        s = df.sample(prog, prog_m, n_particles=npart)
        x0, v0, t10 = oracle.stream_release_numpy(ref, pot, kind, prog.pos.T, prog.vel.T, prog.t, prog_m, npart, mk(),
                                                  v_disp=1.2e-3)
        pot.strict_math = False
        assert s.pos.shape == (3, x0.shape[0]) and np.array_equal(s.release_time, t10)
        off = np.sqrt(((x0 - prog.pos.T[np.searchsorted(prog.t, t10)]) ** 2).sum(1))
        assert np.max(np.abs(s.pos.T - x0) / off[:, None]) < 1e-11
        assert np.max(np.abs(s.vel.T - v0)) / np.abs(v0).max() < 1e-13
    # a whole stream through the generator with each DF
    gen = gb.MockStreamGenerator(df, H)
    stream, _ = gen.run(PROG_W0, 2.5e4, dt=-1.0, n_steps=60, n_particles=2, Integrator="leapfrog")
    assert stream.pos.shape == (3, 2 * 2 * 61) and np.all(np.isfinite(stream.pos))


def test_mockstream_dop853_parity(ref):
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)), H)
    stream, prog = gen.run(PROG_W0, 2.5e4, dt=-1.0, n_steps=120, n_particles=4, release_every=1)
    assert stream.pos.shape == (3, 2 * 4 * 121)
    # oracle: same ICs (re-sampled with the same seed), each particle its own n=6 dop853_step run
    orb, t = _prog_orbit(H, 120)
    prog_orb = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
    s0 = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)).sample(
        prog_orb, 2.5e4, n_particles=4)
    w0 = np.vstack([s0.pos, s0.vel]).T
    tf = prog_orb.t[-1]
    dt0 = prog_orb.t[1] - prog_orb.t[0]
    out = np.empty_like(w0)
    for t1 in np.unique(s0.release_time):
        m = s0.release_time == t1
        rows, st, rc = ref.dop853_step_rows(H, w0[m], t1, tf, dt0, group=True)
        assert rc >= 0
        out[m] = rows
    got = np.vstack([stream.pos, stream.vel])
    d = relnorm(got, out.T)
    print(f"\n[mockstream dop853] median={np.median(d):.2e} max={d.max():.2e}")
    assert d.max() < 1e-9
    # the reference's grouped integration (one shared step size per release group) -- measured, not asserted tight
    outg = np.empty_like(w0)
    for t1 in np.unique(s0.release_time):
        m = s0.release_time == t1
        rows, st, rc = ref.dop853_step_rows(H, w0[m], t1, tf, dt0, group=False)
        outg[m] = rows
    dg = relnorm(got, outg.T)
    print(f"[mockstream dop853 vs reference GROUPED stepping] median={np.median(dg):.2e} max={dg.max():.2e}")
    assert np.median(dg) < 1e-6
    # progenitor end state == direct orbit integration (tests/dynamics/mockstream/test_mockstream.py:663-805)
    direct = H.integrate_orbit(prog_orb[0].w(), t=prog_orb.t, Integrator="dopri853")
    assert np.allclose(prog.w()[:, 0], direct.w()[:, -1].reshape(6), rtol=1e-10)
    assert np.allclose(prog.w()[:, 0], PROG_W0, rtol=1e-7, atol=1e-6)      # integrated back and forth


def test_mockstream_leapfrog_parity(ref):
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(1)), H)
    stream, prog = gen.run(PROG_W0, 2.5e4, dt=-1.0, n_steps=150, n_particles=2, Integrator="leapfrog")
    t = gb.parse_time_specification(None, dt=-1.0, n_steps=150)
    nb = DirectNBody(PROG_W0, [None], external_potential=pot)
    orb = nb.integrate_orbit(t=t, Integrator="leapfrog")
    prog_orb = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
    s0 = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(1)).sample(prog_orb, 2.5e4, n_particles=2)
    w0 = np.vstack([s0.pos, s0.vel])
    tf = prog_orb.t[-1]
    out = np.empty_like(w0)
    for t1 in np.unique(s0.release_time):
        m = s0.release_time == t1
        n_steps = int((tf - t1) / 1.0 + 0.5)
        if n_steps == 0:
            out[:, m] = w0[:, m]
            continue
        tt = t1 + np.arange(n_steps + 1) * 1.0
        out[:, m] = ref.leapfrog(pot, np.ascontiguousarray(w0[:, m]), tt, save_all=False)
    got = np.vstack([stream.pos, stream.vel])
    d = relnorm(got, out)
    print(f"\n[mockstream leapfrog] median={np.median(d):.2e} max={d.max():.2e}")
    assert d.max() < 1e-12
    # leapfrog vs dop853 streams are close (tests/dynamics/mockstream/test_mockstream.py:361-412)
    gen2 = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(1)), H)
    s2, _ = gen2.run(PROG_W0, 2.5e4, dt=-1.0, n_steps=150, n_particles=2)
    assert np.mean(np.sqrt(((s2.pos - stream.pos) ** 2).sum(0))) < 2.0


def test_mockstream_animate_snapshots(ref, tmp_path):
    """mockstream_dop853_animate (mockstream.pyx:306-440) through MockStreamGenerator.run(output_every=...):
    snapshot bookkeeping (count, times, NaN before release, first appearance = the release state) and every
    particle's march -- one dop853_step call per interval -- against the compiled reference."""
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    fn = tmp_path / "snaps.hdf5"
    n_steps, oe = 40, 6
    gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)), H)
    stream, prog = gen.run(PROG_W0, 2.5e4, dt=1.0, n_steps=n_steps, n_particles=1, release_every=2,
                           output_every=oe, output_filename=str(fn))
    from gala_b200.mockstream import mockstream_dop853_animate
    snaps = mockstream_dop853_animate.last
    import os
    assert os.path.exists(mockstream_dop853_animate.last_file)
    out_i = [i for i in range(n_steps + 1) if i % oe == 0 or i == n_steps]
    assert snaps["stream"]["pos"].shape == (3, len(out_i), stream.pos.shape[1])
    assert np.allclose(snaps["stream"]["time"], np.array(out_i, dtype=float))
    t = np.arange(n_steps + 1.0)
    rel = np.asarray(stream.release_time)
    w_snap = np.concatenate([snaps["stream"]["pos"], snaps["stream"]["vel"]])        # (6, nout, Np)
    # same ICs again (same seed) for the reference march
    orb, _ = _prog_orbit(H, n_steps, dt=1.0)
    prog_orb = gb.Orbit(pos=orb.pos[:, :, 0], vel=orb.vel[:, :, 0], t=t, hamiltonian=H)
    s0 = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)).sample(
        prog_orb, 2.5e4, n_particles=1, release_every=2)
    w0 = np.vstack([s0.pos, s0.vel]).T
    for p in range(0, w0.shape[0], 3):
        k0 = int(round(rel[p]))
        y = w0[p:p + 1].copy()
        j = 0
        assert np.all(np.isnan(w_snap[:, 0, p])) == (k0 > 0)
        for i in range(1, n_steps + 1):
            if i > k0:
                y, st, rc = ref.dop853_step_rows(H, y, t[i - 1], t[i], 1.0, group=True)
                assert rc >= 0
            if i % oe == 0 or i == n_steps:
                j += 1
                if i < k0:
                    assert np.all(np.isnan(w_snap[:, j, p]))
                else:
                    assert np.allclose(w_snap[:, j, p], y[0], rtol=1e-11, atol=1e-13), (p, i)
        assert np.allclose(np.concatenate([stream.pos[:, p], stream.vel[:, p]]), y[0], rtol=1e-11, atol=1e-13)
    # the progenitor's snapshots follow its orbit
    assert np.allclose(snaps["nbody"]["pos"][:, -1, 0], prog.pos.ravel())
    with pytest.raises(IOError):
        gen.run(PROG_W0, 2.5e4, dt=1.0, n_steps=4, n_particles=1, output_every=2, output_filename=str(fn))
    with pytest.raises(NotImplementedError):
        gen.run(PROG_W0, 2.5e4, dt=1.0, n_steps=4, n_particles=1, output_every=2, output_filename=str(fn),
                Integrator="leapfrog", overwrite=True)


@pytest.mark.parametrize("integ", ["leapfrog", "dopri853"])
@pytest.mark.parametrize("case", ["backward_every_step", "forward_release_every_3", "streakline"])
def test_pipelined_run_equals_stepwise_run(monkeypatch, integ, case):
    """MockStreamGenerator.run's device-resident pipeline (deviates drawn chunk by chunk on a worker thread, chunks
    integrated on their own streams, the progenitor's second integration on a side stream) returns the bits of the
    step-by-step form (GB_NO_STREAM_PIPELINE=1), for both integrators, both time directions and release_every > 1."""
    H = gb.Hamiltonian(gb.MilkyWayPotential2022())
    prog = gb.PhaseSpacePosition(pos=[13.0, 0.0, 20.0], vel=np.array([0.0, 130.0, 50.0]) * gb.KMS_TO_KPC_MYR)

    def run():
        if case == "streakline":
            df = gb.StreaklineStreamDF()
        else:
            df = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(5))
        gen = gb.MockStreamGenerator(df, H)
        if case == "forward_release_every_3":
            s, p = gen.run(prog, 2.5e4, dt=1.0, n_steps=900, n_particles=7, release_every=3, Integrator=integ)
        else:
            s, p = gen.run(prog, 2.5e4, dt=-1.0, n_steps=1200, n_particles=6, release_every=1, Integrator=integ)
        return s.w(), p.w(), s.release_time, s.lead_trail

    monkeypatch.setenv("GB_STREAM_CHUNKS", "5")
    got = run()
    monkeypatch.setenv("GB_NO_STREAM_PIPELINE", "1")
    want = run()
    assert got[0].shape == want[0].shape and got[0].shape[1] > 4096
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
