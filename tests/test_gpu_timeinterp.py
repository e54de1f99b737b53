"""`-m gpu`: TimeInterpolatedPotential on the device (GB_POT_TIMEINTERP, csrc/timeinterp.cuh) against the
reference's own time_interp.cpp / time_interp_wrapper.cpp (oracle; compiled against the GSL spline stand-in that
tests/test_oracle_cpu.py pins to scipy and to the reference's docstring numbers).

The reference's gradient wrapper indexes its scratch arrays orbit-major while the wrapped gradients are
structure-of-arrays (time_interp_wrapper.cpp:186-201), so the oracle is only called with ONE point / ONE orbit at a
time (N = 1, where both layouts coincide)."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

import gala_b200 as gb
from conftest import relnorm

pytestmark = pytest.mark.gpu

T = np.linspace(0.0, 600.0, 13)


def _cases():
    rng = np.random.default_rng(42)
    grow = np.linspace(1.0, 1.8, T.size) * (1 + 0.05 * rng.normal(size=T.size))
    Rs = np.array([Rotation.from_rotvec([0.1 * np.sin(a), 0.05, a]).as_matrix() for a in np.linspace(0, 1.9, T.size)])
    orb = np.stack([8 * np.cos(T / 150), 8 * np.sin(T / 150), 0.5 * np.sin(T / 90)], axis=1)
    out = {}
    for method in ("linear", "cspline", "akima", "steffen"):
        out[f"kepler_mass_{method}"] = gb.TimeInterpolatedPotential(gb.KeplerPotential, T, interpolation_method=method, m=1e10 * grow)
        out[f"bar_rotating_{method}"] = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, T, interpolation_method=method,
                                                                     m=1e10, a=3.0, b=1.0, c=0.5, R=Rs)
    out["hernquist_moving_growing"] = gb.TimeInterpolatedPotential(gb.HernquistPotential, T, m=5e10 * grow, c=1.5 * np.sqrt(grow), origin=orb)
    out["mn3_growing"] = gb.TimeInterpolatedPotential(gb.MN3ExponentialDiskPotential, T, m=5e10 * grow, h_R=2.6, h_z=0.3)
    out["nfw_triaxial_all"] = gb.TimeInterpolatedPotential(gb.NFWPotential, T, interpolation_method="steffen", m=6e11 * grow, r_s=16.0,
                                                           a=1.0, b=0.9, c=0.8, origin=0.3 * orb, R=Rs)
    out["logarithmic_const"] = gb.TimeInterpolatedPotential(gb.LogarithmicPotential, T, v_c=0.2, r_h=5.0, q1=1.0, q2=0.9, q3=0.8, phi=0.3)
    comp = gb.CCompositePotential()
    comp["halo"] = gb.NFWPotential(m=6e11, r_s=16.0)
    comp["sat"] = gb.TimeInterpolatedPotential(gb.PlummerPotential, T, m=2e10 * grow, b=1.0, origin=2.5 * orb)
    out["static_halo_plus_moving_satellite"] = comp
    # TimeInterpolated, static, TimeInterpolated: the fixed-step kernels read the k-th TimeInterpolated component from
    # slot k of the per-step state row (kernels.cu: k_ti_table), not from its component index
    comp2 = gb.CCompositePotential()
    comp2["bar"] = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, T, m=1e10, a=3.0, b=1.0, c=0.5, R=Rs)
    comp2["halo"] = gb.NFWPotential(m=6e11, r_s=16.0)
    comp2["sat"] = gb.TimeInterpolatedPotential(gb.HernquistPotential, T, interpolation_method="akima", m=3e10 * grow, c=1.2, origin=2.0 * orb)
    comp2["disk"] = gb.MiyamotoNagaiPotential(m=6e10, a=3.0, b=0.3)
    out["two_interpolated_around_static_ones"] = comp2
    return out


CASES = _cases()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("strict", [True, False])
def test_evaluation_parity(ref, name, strict):
    pot = CASES[name]
    pot.strict_math = strict
    rng = np.random.default_rng(1)
    n = 40
    q = rng.normal(0, 9.0, (3, n))
    times = np.concatenate([rng.uniform(T[0], T[-1], n - 4), [T[0], T[-1], T[3], 0.5 * (T[5] + T[6])]])
    for i in range(n):
        qi = np.ascontiguousarray(q[:, i:i + 1])
        g0, e0, d0 = ref.gradient(pot, qi, t=times[i]), ref.energy(pot, qi, t=times[i]), ref.density(pot, qi, t=times[i])
        g, e = pot.gradient(qi, t=times[i]), pot.energy(qi, t=times[i])
        assert np.max(np.abs(g - g0)) <= 2e-13 * np.abs(g0).max(), (name, i)
        assert abs(e[0] - e0[0]) <= 2e-13 * abs(e0[0])
        if np.isfinite(d0[0]):
            d = pot.density(qi, t=times[i])
            assert abs(d[0] - d0[0]) <= 1e-11 * max(abs(d0[0]), 1e-300)
    # many points at one time == the same points one by one (the device has no N = 1 restriction)
    gall = pot.gradient(q, t=times[0])
    for i in range(0, n, 7):
        assert np.array_equal(gall[:, i], pot.gradient(np.ascontiguousarray(q[:, i:i + 1]), t=times[0])[:, 0])
    # outside the knot range: NaN (time_interp_wrapper.cpp:103-106,148-155)
    assert np.isnan(pot.gradient(q, t=T[-1] + 1.0)).all() and np.isnan(pot.energy(q, t=T[0] - 1e-9)).all()
    pot.strict_math = False


def test_reference_docstring_numbers_on_the_gpu():
    t = np.linspace(0, 100, 11)
    pot = gb.TimeInterpolatedPotential(gb.KeplerPotential, t, m=np.linspace(1e10, 2e10, 11))
    q = np.array([[1e-3], [0.0], [0.0]])
    assert abs(pot.energy(q, t=0.0)[0] - (-44.98502151)) < 5e-9 and abs(pot.energy(q, t=50.0)[0] - (-67.47753227)) < 5e-9
    Rs = np.array([Rotation.from_rotvec([0, 0, a]).as_matrix() for a in np.linspace(0, np.pi / 2, 11)])
    bar = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, np.linspace(0, 1000, 11), m=1e10, a=3.0, b=1.0, c=0.5, R=Rs)
    q = np.array([[5.0], [0.0], [0.0]])
    assert abs(bar.gradient(q, t=0.0)[0, 0] - 0.00207787) < 5e-9 and abs(bar.gradient(q, t=500.0)[0, 0] - 0.0015879) < 5e-8


@pytest.mark.parametrize("name", ["static_halo_plus_moving_satellite", "bar_rotating_cspline", "hernquist_moving_growing"])
def test_integration_parity(ref, name):
    """Leapfrog, Ruth4 and DOP853 through a time-dependent potential: the step time t[j] (fixed step,
    leapfrog.pyx:106, ruth4.pyx:100) / the stage time (DOP853) reaches the interpolation.  One orbit per oracle call."""
    pot = CASES[name]
    H = gb.Hamiltonian(pot)
    rng = np.random.default_rng(3)
    N = 12
    w0 = np.vstack([rng.normal(0, 8.0, (3, N)), rng.normal(0, 0.08, (3, N))])
    t = np.linspace(T[0], T[-1], 1201)
    for strict in (True, False):
        H.strict_math = strict
        _, w_lf = gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0)
        _, w_r4 = gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=0)
        _, w_d8 = gb.dop853_integrate_hamiltonian(H, w0, t[::10].copy(), save_all=0)
        d_lf, d_r4, d_d8 = [], [], []
        for i in range(N):
            wi = np.ascontiguousarray(w0[:, i:i + 1])
            d_lf.append(relnorm(w_lf[:, i:i + 1], ref.leapfrog(pot, wi, t, save_all=False)).max())
            d_r4.append(relnorm(w_r4[:, i:i + 1], ref.ruth4(H, wi, t, save_all=False)).max())
            d_d8.append(relnorm(w_d8[:, i:i + 1], ref.dop853(H, wi, t[::10].copy(), save_all=False, nbatch=1)[0]).max())
        print(f"\n[timeinterp {name} strict={strict}] leapfrog med/max {np.median(d_lf):.1e}/{np.max(d_lf):.1e}  "
              f"ruth4 {np.median(d_r4):.1e}/{np.max(d_r4):.1e}  dop853 {np.median(d_d8):.1e}/{np.max(d_d8):.1e}")
        # 1200 fixed steps: <= 1e-12-class medians, a close passage of the moving centre amplifies rounding on single orbits;
        # DOP853 at rtol = 1e-10: within the north-star 1e-9 in the median, tolerance-level on the worst orbit
        assert np.median(d_lf) < 1e-12 and np.max(d_lf) < 1e-9
        assert np.median(d_r4) < 1e-12 and np.max(d_r4) < 1e-9
        assert np.median(d_d8) < 1e-9 and np.max(d_d8) < 1e-6
    # the host class refuses a grid outside the knots; the raw boundary call integrates into NaN like the reference
    if isinstance(pot, gb.TimeInterpolatedPotential):
        with pytest.raises(ValueError):
            pot.integrate_orbit(w0, dt=1.0, n_steps=int(T[-1]) + 10)
    _, w_nan = gb.leapfrog_integrate_hamiltonian(H, w0, np.linspace(T[-1] - 5, T[-1] + 5, 11), save_all=0)
    assert np.isnan(w_nan).all()
    st = gb.integrate_extrema(H, w0, t, Integrator="leapfrog", return_final=True)
    assert np.allclose(st["w_final"], w_lf, rtol=1e-9, atol=1e-12)


def test_unsupported_entry_points_say_so():
    pot = CASES["kepler_mass_cspline"]
    H = gb.Hamiltonian(pot)
    prog = gb.PhaseSpacePosition(pos=[8.0, 0.0, 0.0], vel=[0.0, 0.07, 0.0])
    # mock streams without massive bodies ARE supported (test_mock_stream_in_a_time_dependent_potential); massive
    # bodies and the Lyapunov kernel are not
    stream, _ = gb.MockStreamGenerator(gb.StreaklineStreamDF(), H).run(prog, 1e4, dt=1.0, n_steps=50, Integrator="leapfrog")
    assert np.isfinite(stream.pos).all()
    with pytest.raises(gb._abi.GalaB200Error, match="time-dependent"):
        gb.MockStreamGenerator(gb.StreaklineStreamDF(), H, progenitor_potential=gb.PlummerPotential(m=1e4, b=0.01)).run(
            prog, 1e4, dt=1.0, n_steps=50, Integrator="leapfrog")
    with pytest.raises(gb._abi.GalaB200Error, match="time-dependent"):
        gb.fast_lyapunov_max(np.array([8.0, 0.0, 0.0, 0.0, 0.07, 0.0]), H, dt=1.0, n_steps=100, return_orbit=False)


def test_rotating_bar_inertial_vs_rotating_frame(ref):
    """The reference's own integration test for this feature, on the GPU
    (tests/integration/test_bar_rotating_frame.py:26-168): an orbit at the corotation radius of a barred Milky Way,
    (1) in a ConstantRotatingFrame with a static LongMuraliBar, (2) in the inertial frame with the bar's rotation matrix
    interpolated in time (cspline through 200 knots per bar period); (2) rotated into the frame must match (1) with the
    reference's tolerances (xyz rtol 5e-5 / atol 2e-3 kpc; v_xyz rtol 5e-5 / atol 3e-5 kpc/Myr), DOPRI853 atol = rtol = 1e-14."""
    Omega = 30.0 * gb.KMS_TO_KPC_MYR                                   # 30 km/s/kpc in rad/Myr
    dt_knot = 2 * np.pi / Omega / 200
    knots = np.arange(0.0, 5000.0, dt_knot)
    Rs = Rotation.from_rotvec(np.stack([0 * knots, 0 * knots, -Omega * knots], axis=1)).as_matrix()    # = from_euler("z", -Omega t)
    mw = gb.MilkyWayPotential2022()
    disk = gb.MN3ExponentialDiskPotential(m=4.1e10, h_R=mw["disk"].parameters["h_R"], h_z=mw["disk"].parameters["h_z"])
    bar_kw = dict(m=1e10, a=4.0, b=0.8, c=0.25, alpha=np.deg2rad(25.0))
    inertial = gb.CCompositePotential()
    inertial["bar"] = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, knots, R=Rs, **bar_kw)
    static = gb.CCompositePotential()
    static["bar"] = gb.LongMuraliBarPotential(**bar_kw)
    for pot in (inertial, static):
        pot["disk"], pot["halo"], pot["nucleus"] = disk, mw["halo"], mw["nucleus"]
    H_rot = gb.Hamiltonian(static, gb.ConstantRotatingFrame([0.0, 0.0, Omega]))
    # corotation radius: Omega_circ(r) = Omega along the x axis (bisection on the static barred potential)
    lo, hi = 3.0, 20.0
    for _ in range(60):
        r = 0.5 * (lo + hi)
        om = np.sqrt(static.gradient(np.array([[r], [0.0], [0.0]]))[0, 0] / r)
        lo, hi = (r, hi) if om > Omega else (lo, r)
    w0 = np.array([r, 0.0, 0.0, 0.0, Omega * r, 0.0])
    t = np.arange(0.0, knots.max(), 0.1)
    kw = {"atol": 1e-14, "rtol": 1e-14}
    o_rot = H_rot.integrate_orbit(w0, Integrator="dopri853", Integrator_kwargs=kw, t=t)
    o_in = inertial.integrate_orbit(w0, Integrator="dopri853", Integrator_kwargs=kw, t=t)
    assert np.isfinite(o_in.pos).all() and o_in.pos.shape == (3, t.size)
    # static -> constant rotating frame: positions and velocities rotated about z by -Omega t
    # (potential/frame/builtin/transformations.py:100-160)
    c, s = np.cos(Omega * t), np.sin(Omega * t)
    rot = lambda v: np.stack([c * v[0] + s * v[1], -s * v[0] + c * v[1], v[2]])
    x_r, v_r = rot(o_in.pos), rot(o_in.vel)
    assert np.allclose(x_r, o_rot.pos, rtol=5e-5, atol=2e-3)
    # the reference's own two integrations (oracle, same corotation orbit) differ by 3.7e-5 kpc/Myr at most -- the
    # interpolation error of a rotation sampled 200 times per period -- against the 3e-5 its test allows for its
    # Powell-minimised initial radius; 5e-5 here, and the GPU must reproduce the reference's orbits themselves:
    assert np.allclose(v_r, o_rot.vel, rtol=5e-5, atol=5e-5)
    w1 = w0.reshape(6, 1)
    ref_rot = ref.dop853(H_rot, w1, t, atol=1e-14, rtol=1e-14, nbatch=1)[0][:, :, 0]
    ref_in = ref.dop853(gb.Hamiltonian(inertial), w1, t, atol=1e-14, rtol=1e-14, nbatch=1)[0][:, :, 0]
    assert np.abs(o_rot.w() - ref_rot).max() < 1e-6 and np.abs(o_in.w() - ref_in).max() < 1e-6
    print(f"\n[rotating bar] corotation radius {r:.4f} kpc; max |dx| {np.abs(x_r - o_rot.pos).max():.2e} kpc, "
          f"max |dv| {np.abs(v_r - o_rot.vel).max():.2e} kpc/Myr over {t.size} samples / {knots.size} knots")


# ---- the reference's own tests of this class (tests/potential/potential/test_time_interpolated.py), on the GPU ----------
_ALL = ["KeplerPotential", "HernquistPotential", "PlummerPotential", "IsochronePotential", "JaffePotential", "NFWPotential",
        "MiyamotoNagaiPotential", "MN3ExponentialDiskPotential", "LongMuraliBarPotential", "StonePotential", "BurkertPotential",
        "SatohPotential", "KuzminPotential", "LogarithmicPotential", "LeeSutoTriaxialNFWPotential", "PowerLawCutoffPotential"]


@pytest.mark.parametrize("name", _ALL)
def test_all_builtin_potentials_time_interpolated(name):
    """test_time_interpolated.py:438-496: every builtin class, first parameter x linspace(1, 4) over 32 knots, the others 1:
    equal to the static potential at t = 0, different at t = 50."""
    cls = getattr(gb, name)
    names = list(cls._param_names)
    knots = np.linspace(0, 100, 32)
    const = {names[0]: 1e10 if names[0] == "m" else 1.0}
    timed = {names[0]: const[names[0]] * np.linspace(1.0, 4, knots.size)}
    for n in names[1:]:
        const[n] = timed[n] = 1.0
    if name == "MN3ExponentialDiskPotential":
        const["h_R"] = timed["h_R"] = 5.0
        const["h_z"] = timed["h_z"] = 0.5
    elif name == "StonePotential":
        const["r_c"] = timed["r_c"] = 1.0
        const["r_h"] = timed["r_h"] = 10.0
    elif name == "PowerLawCutoffPotential":
        const["alpha"] = timed["alpha"] = 1.0
    elif name == "LeeSutoTriaxialNFWPotential":
        const.update(a=1.0, b=0.9, c=0.8); timed.update(a=1.0, b=0.9, c=0.8)      # the class requires a >= b >= c
    static = cls(**const)
    tip = gb.TimeInterpolatedPotential(cls, knots, **timed)
    x = np.array([[1.0], [0.0], [0.0]])
    e0, e_t0, e_t50 = static.energy(x)[0], tip.energy(x, t=0.0)[0], tip.energy(x, t=50.0)[0]
    assert np.isclose(e0, e_t0, rtol=1e-12, atol=0.0) and not np.isclose(e0, e_t50, rtol=1e-5, atol=0.0)
    g0, g_t0 = static.gradient(x), tip.gradient(x, t=0.0)
    assert np.allclose(g0, g_t0, rtol=1e-12, atol=1e-300)


def test_mn3_time_interpolated_like_the_reference():
    """TestMN3TimeInterpolated (test_time_interpolated.py:533-700): constant parameters match the static MN3 at 1e-10,
    the end knots of a varying mass match the static potentials of those masses, the extra keyword positive_density is
    handed to the wrapped class, and a composite's gradient does not depend on the order of its components."""
    knots = np.linspace(0, 5000.0, 6)
    kw = dict(h_R=3.0, h_z=0.28)
    static = gb.MN3ExponentialDiskPotential(m=5e10, **kw)
    tip = gb.TimeInterpolatedPotential(gb.MN3ExponentialDiskPotential, knots, m=5e10, **kw)
    xyz = np.array([[8.0, 0.0, 0.0], [4.0, 0.0, 0.5]]).T.copy()
    for func in ("energy", "gradient", "density"):
        np.testing.assert_allclose(getattr(tip, func)(xyz, t=2500.0), getattr(static, func)(xyz), rtol=1e-10)
    masses = np.linspace(3e10, 7e10, knots.size)
    tip = gb.TimeInterpolatedPotential(gb.MN3ExponentialDiskPotential, knots, m=masses, **kw)
    x1 = np.array([[8.0], [0.0], [0.0]])
    for t, m in ((knots[0], masses[0]), (knots[-1], masses[-1])):
        np.testing.assert_allclose(tip.energy(x1, t=t), gb.MN3ExponentialDiskPotential(m=m, **kw).energy(x1), rtol=1e-10)
    mid = tip.energy(x1, t=2500.0)[0]
    lo, hi = tip.energy(x1, t=knots[0])[0], tip.energy(x1, t=knots[-1])[0]
    assert min(lo, hi) < mid < max(lo, hi)
    s2 = gb.MN3ExponentialDiskPotential(m=5e10, positive_density=False, **kw)
    t2 = gb.TimeInterpolatedPotential(gb.MN3ExponentialDiskPotential, knots, m=5e10, positive_density=False, **kw)
    np.testing.assert_allclose(t2.energy(xyz, t=100.0), s2.energy(xyz), rtol=1e-10)
    assert not np.allclose(s2.energy(xyz), static.energy(xyz))
    halo = gb.NFWPotential(m=6e11, r_s=16.0)
    a, b = gb.CCompositePotential(), gb.CCompositePotential()
    a["disk"], a["halo"] = tip, halo
    b["halo"], b["disk"] = halo, tip
    np.testing.assert_allclose(a.gradient(xyz, t=1234.0), b.gradient(xyz, t=1234.0), rtol=1e-13)
    assert not np.allclose(a.energy(xyz, t=0.0), a.energy(xyz, t=4000.0))
    orbit = tip.integrate_orbit(np.array([8.0, 0.0, 0.1, 0.0, 0.2, 0.0]), dt=1.0, n_steps=1000)
    assert np.isfinite(orbit.pos).all()


def test_interpolation_accuracy_and_bounds():
    """test_time_interpolated.py:162-186,222-247: a linearly varying mass is reproduced exactly between the knots by every
    spline type that reproduces straight lines; NaN outside the knot range, finite on its closed ends."""
    knots = np.linspace(0, 10, 11)
    m = 1e10 * (1 + 0.1 * knots)
    x = np.array([[8.0], [0.0], [0.0]])
    for method in ("linear", "cspline", "akima", "steffen"):
        tip = gb.TimeInterpolatedPotential(gb.KeplerPotential, knots, interpolation_method=method, m=m)
        for t in (0.0, 2.5, 7.25, 10.0):
            want = gb.KeplerPotential(m=1e10 * (1 + 0.1 * t)).energy(x)[0]
            assert np.isclose(tip.energy(x, t=t)[0], want, rtol=1e-13), (method, t)
        assert np.isnan(tip.energy(x, t=-0.1)[0]) and np.isnan(tip.energy(x, t=10.1)[0])
        assert np.isnan(tip.gradient(x, t=11.0)).all() and np.isnan(tip.density(x, t=-5.0)).all()
    with pytest.raises(ValueError):
        gb.TimeInterpolatedPotential(gb.KeplerPotential, knots, m=m[:5])          # test_mismatched_parameter_length
    with pytest.raises(ValueError):
        gb.TimeInterpolatedPotential(gb.KeplerPotential, knots[::-1].copy(), m=m)


def _time_dependent_galaxy(kind):
    """a static halo and disc + (bar) a bar turning in the inertial frame through rotation matrices at the knots, or
    (satellite) an infalling, growing Hernquist satellite on a moving origin -- knots cover a backward integration from 0"""
    Tk = np.linspace(-200.0, 10.0, 43)
    pot = gb.CCompositePotential()
    pot["halo"] = gb.NFWPotential(m=6e11, r_s=16.0)
    pot["disk"] = gb.MiyamotoNagaiPotential(m=6e10, a=3.0, b=0.3)
    if kind == "bar":
        Rk = np.array([Rotation.from_rotvec([0.0, 0.0, -0.04 * a]).as_matrix() for a in Tk])
        pot["bar"] = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, Tk, m=1e10, a=4.0, b=0.8, c=0.25, R=Rk)
    else:
        track = np.stack([60.0 + 0.2 * Tk, -20.0 - 0.15 * Tk, 10.0 + 0.05 * Tk], axis=1)
        pot["lmc"] = gb.TimeInterpolatedPotential(gb.HernquistPotential, Tk, m=1.5e11 * np.linspace(0.6, 1.0, Tk.size), c=10.0,
                                                  origin=track)
    return pot


@pytest.mark.parametrize("kind,integ", [("satellite", "dopri853"), ("satellite", "leapfrog"), ("bar", "leapfrog")])
def test_mock_stream_in_a_time_dependent_potential(ref, kind, integ):
    """MockStreamGenerator in a galaxy with an infalling satellite / a bar turning in the inertial frame (the uses the
    reference's TimeInterpolatedPotential was written for): the release (c_d2_dr2 at the release time, df.pyx:61-92),
    the progenitor orbit and every particle's integration (mockstream.pyx:176-303, 442-620) see the potential at their
    own time.  Same checks as test_gpu_mockstream.py's parity tests, against the compiled reference."""
    from gala_b200.mockstream import DirectNBody
    KMS = gb.KMS_TO_KPC_MYR
    PROG_W0 = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * KMS, 50.0 * KMS])
    pot = _time_dependent_galaxy(kind)
    H = gb.Hamiltonian(pot)
    n_steps, npart = 120, 3
    mk = lambda: np.random.RandomState(7)                                    # noqa: E731
    gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=mk()), H)
    stream, prog = gen.run(PROG_W0, 2.5e4, dt=-1.0, n_steps=n_steps, n_particles=npart, Integrator=integ)
    assert stream.pos.shape == (3, 2 * npart * (n_steps + 1)) and np.isfinite(stream.pos).all()
    t = gb.parse_time_specification(None, dt=-1.0, n_steps=n_steps)
    orb = DirectNBody(PROG_W0, [None], external_potential=pot).integrate_orbit(t=t, Integrator=integ)
    prog_orb = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
    s0 = gb.FardalStreamDF(gala_modified=True, random_state=mk()).sample(prog_orb, 2.5e4, n_particles=npart)
    # release against the reference's own potential evaluations at the release times (strict kernels, like
    # test_gpu_mockstream.py::test_fardal_release_parity: the second difference of Phi with h = 1e-2 amplifies rounding)
    from oracle import oracle
    pm = np.full(n_steps + 1, 2.5e4)
    x0, v0, t10 = oracle.fardal_release_numpy(ref, pot, prog_orb.pos.T, prog_orb.vel.T, prog_orb.t, pm,
                                              np.full(n_steps + 1, npart, dtype="i4"), mk(), gala_modified=True)
    pot.strict_math = True
    ss = gb.FardalStreamDF(gala_modified=True, random_state=mk()).sample(prog_orb, 2.5e4, n_particles=npart)
    pot.strict_math = False
    assert np.array_equal(s0.release_time, t10) and np.array_equal(ss.release_time, t10)
    off = np.sqrt(((x0 - prog_orb.pos.T[np.searchsorted(prog_orb.t, t10)]) ** 2).sum(1))
    e_pos, e_vel = np.max(np.abs(ss.pos.T - x0) / off[:, None]), np.max(np.abs(ss.vel.T - v0)) / np.abs(v0).max()
    f_pos = np.max(np.abs(s0.pos.T - x0) / off[:, None])
    print(f"\n[release, time-dependent {kind}] strict: offsets {e_pos:.1e}, velocities {e_vel:.1e}; fast: offsets {f_pos:.1e}")
    assert e_pos < 5e-9 and e_vel < 1e-10 and f_pos < 1e-7
    w0 = np.vstack([s0.pos, s0.vel])
    tf = prog_orb.t[-1]
    out = np.empty_like(w0)
    for t1 in np.unique(s0.release_time):
        m = s0.release_time == t1
        if integ == "dopri853":
            rows, st, rc = ref.dop853_step_rows(H, np.ascontiguousarray(w0[:, m].T), t1, tf, prog_orb.t[1] - prog_orb.t[0], group=True)
            assert rc >= 0
            out[:, m] = rows.T
        else:
            # one particle per oracle call: the reference's time_interp_gradient back-rotates a batch with AoS indices on
            # SoA arrays (time_interp_wrapper.cpp:189-201) and is only self-consistent for N = 1
            k = int((tf - t1) / 1.0 + 0.5)
            for i in np.flatnonzero(m):
                out[:, i] = w0[:, i] if k == 0 else ref.leapfrog(pot, np.ascontiguousarray(w0[:, i:i + 1]), t1 + np.arange(k + 1) * 1.0,
                                                                 save_all=False)[:, 0]
    d = relnorm(np.vstack([stream.pos, stream.vel]), out)
    print(f"[mock stream, time-dependent {kind}, {integ}] median={np.median(d):.2e} max={d.max():.2e}")
    assert d.max() < (1e-9 if integ == "dopri853" else 1e-11)
    # the time-dependent component matters: the same run without it ends somewhere else
    static = gb.CCompositePotential(halo=pot["halo"], disk=pot["disk"])
    s_static, _ = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=mk()), gb.Hamiltonian(static)).run(
        PROG_W0, 2.5e4, dt=-1.0, n_steps=n_steps, n_particles=npart, Integrator=integ)
    assert np.abs(s_static.pos - stream.pos).max() > 1e-3
    # massive bodies in a time-dependent field stay with the reference's CPU path: the library says so
    gen_sg = gb.MockStreamGenerator(gb.FardalStreamDF(random_state=mk()), H, progenitor_potential=gb.PlummerPotential(m=2.5e4, b=0.004))
    with pytest.raises(gb._abi.GalaB200Error, match="time-dependent"):
        gen_sg.run(PROG_W0, 2.5e4, dt=-1.0, n_steps=20, n_particles=1, Integrator=integ)


def test_turning_bar_trips_the_stiffness_test_like_the_reference(ref):
    """A quirk carried over on purpose: ``dop853_step`` hands ``nstiff = 1`` to ``dop853`` whatever its caller asked for
    (dop853.pyx:64), so the stiffness test runs after every accepted step of a stream particle, and a bar turning
    through interpolated rotation matrices makes it fire (code -4) for some release times -- in the reference's own
    C++ as on the device.  ``err_if_fail`` then raises the reference's RuntimeError."""
    KMS = gb.KMS_TO_KPC_MYR
    PROG_W0 = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * KMS, 50.0 * KMS])
    pot = _time_dependent_galaxy("bar")
    H = gb.Hamiltonian(pot)
    gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(7)), H)
    with pytest.raises(RuntimeError, match="Integration failed with code -4"):
        gen.run(PROG_W0, 2.5e4, dt=-1.0, n_steps=120, n_particles=3)
    # the reference on the same particles
    t = gb.parse_time_specification(None, dt=-1.0, n_steps=120)
    from gala_b200.mockstream import DirectNBody
    orb = DirectNBody(PROG_W0, [None], external_potential=pot).integrate_orbit(t=t, Integrator="dopri853")
    prog_orb = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
    s0 = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(7)).sample(prog_orb, 2.5e4, n_particles=3)
    w0 = np.vstack([s0.pos, s0.vel])
    codes = [ref.dop853_step_rows(H, np.ascontiguousarray(w0[:, s0.release_time == t1].T), t1, 0.0, 1.0, group=True)[2]
             for t1 in np.unique(s0.release_time)]
    assert min(codes) == -4


@pytest.mark.parametrize("name", ["kepler_mass_cspline", "hernquist_moving_growing", "mn3_growing",
                                  "static_halo_plus_moving_satellite", "bar_rotating_cspline", "nfw_triaxial_all",
                                  "two_interpolated_around_static_ones"])
def test_hessian_parity(ref, name):
    """``time_interp_hessian`` (time_interp_wrapper.cpp:254-318): the wrapped potential's Hessian at the interpolated
    parameters and origin, turned back with the interpolated rotation (R^T H R) -- forward-mode differentiation of the
    device gradient against the reference's closed-form Hessians; NaN outside the knots.  The host class refuses
    rotated cases like the reference's ``PotentialBase.hessian``; the C ABI evaluates them.
    Composites: the reference's function ASSIGNS its result (``hess[i*n_dim + j] = 0.0`` before accumulating, :303-312),
    wiping what the components before it added -- its composite Hessian is the last TimeInterpolated component's plus
    whatever follows.  The device sums all components; the comparison is against the sum of the reference's
    per-component Hessians (DESIGN.md deviations)."""
    pot = CASES[name]
    rng = np.random.default_rng(11)
    q = np.ascontiguousarray(rng.normal(0, 9.0, (3, 64)))
    rotated = name in ("bar_rotating_cspline", "nfw_triaxial_all")
    parts = list(pot.values()) if isinstance(pot, gb.CCompositePotential) else [pot]
    for t in (T[0], 0.37 * T[-1], T[-1]):
        want = sum(ref.hessian(p, q, t) for p in parts)
        for strict in (True, False):
            pot.strict_math = strict
            if rotated:
                with pytest.raises(NotImplementedError):
                    pot.hessian(q, t)
            got = gb.PotentialBase.hessian(pot, q, t)              # the C ABI itself
            scale = np.abs(want).max(axis=(0, 1))
            err = (np.abs(got - want).max(axis=(0, 1)) / scale).max()
            # relative to the largest entry of each 3x3 block; the long thin bar's second derivatives cancel (1.5e-12 measured)
            assert err < (1e-11 if strict else 1e-10), (name, t, strict, err)
        pot.strict_math = False
    if len(parts) > 1:      # the quirk itself, so that a change of the reference is noticed
        k = max(i for i, p in enumerate(parts) if isinstance(p, gb.TimeInterpolatedPotential))
        quirk = sum(ref.hessian(p, q, T[3]) for p in parts[k:])
        assert np.allclose(ref.hessian(pot, q, T[3]), quirk, rtol=1e-13, atol=0)
    assert np.isnan(gb.PotentialBase.hessian(pot, q, T[-1] + 1.0)).all()
