"""`-m gpu`: the per-point callers either side of the gradient / energy kernels -- ``PotentialBase.mass_enclosed``
(core.py:649-723), ``circular_velocity`` (core.py:725-784), ``PhaseSpacePosition.kinetic_energy / potential_energy /
angular_momentum`` (dynamics/core.py:652-740) -- against the compiled reference's ``c_potential`` / ``c_gradient`` put
through the same numpy arithmetic the reference's Python does, and against closed forms."""
import numpy as np
import pytest

import gala_b200 as gb
from conftest import make_ic

pytestmark = pytest.mark.gpu
G = gb.G_GALACTIC


def _pots():
    return {
        "hernquist": gb.HernquistPotential(m=1e11, c=10.0),
        "nfw": gb.NFWPotential(m=6e11, r_s=16.0),
        "plummer": gb.PlummerPotential(m=1e10, b=1.0),
        "mw2022": gb.MilkyWayPotential2022(),
        "lm10": gb.LM10Potential(),
        "kepler_negative": gb.KeplerPotential(m=-2e10),
    }


@pytest.mark.parametrize("name", list(_pots()))
def test_mass_enclosed_and_circular_velocity_match_reference_arithmetic(ref, name):
    pot = _pots()[name]
    pot.strict_math = True
    q = make_ic(lambda x: gb.MilkyWayPotential2022().gradient(x), 500, seed=4)[:3].copy()
    r = np.sqrt((q * q).sum(0))
    # core.py:703-716 with the reference's own c_potential
    h = 1e-3
    eps = h * q / r
    diff = ref.energy(pot, np.ascontiguousarray(q + eps)) - ref.energy(pot, np.ascontiguousarray(q - eps))
    want_m = np.abs(r * r * diff / pot.G / (2.0 * h)) * (-1.0 if name == "kepler_negative" else 1.0)
    got_m = pot.mass_enclosed(q)
    # a difference of two potentials 2e-3 kpc apart: relative rounding of Phi (1e-16) is amplified by Phi / (h dPhi/dr)
    assert np.allclose(got_m, want_m, rtol=1e-9, atol=0), np.abs(got_m / want_m - 1).max()
    # core.py:773-777 with the reference's own c_gradient
    want_v = np.sqrt(r * np.abs((ref.gradient(pot, q) * q / r).sum(0)))
    got_v = pot.circular_velocity(q)
    assert np.allclose(got_v, want_v, rtol=1e-13, atol=0), np.abs(got_v / want_v - 1).max()


def test_closed_forms_and_device_arrays():
    torch = pytest.importorskip("torch")
    m, c = 1e11, 10.0
    pot = gb.HernquistPotential(m=m, c=c)
    rng = np.random.default_rng(42)
    R = rng.uniform(4, 10, 128)                                   # tests/dynamics/test_dynamics_core.py:378-387
    q = R[None] * np.array([1.0, 0, 0])[:, None]
    vc = pot.circular_velocity(q)
    assert np.allclose(vc, np.sqrt(G * m * R) / (R + c), rtol=1e-13)
    assert np.allclose(pot.mass_enclosed(q), m * R ** 2 / (R + c) ** 2, rtol=1e-6)     # centred difference, h = 1e-3
    # torch.cuda in -> torch.cuda out, same numbers
    qd = torch.as_tensor(q, device="cuda")
    vd = pot.circular_velocity(qd)
    md = pot.mass_enclosed(qd)
    assert vd.is_cuda and md.is_cuda
    assert np.allclose(vd.cpu().numpy(), vc, rtol=1e-14) and np.allclose(md.cpu().numpy(), pot.mass_enclosed(q), rtol=1e-9)
    # NFWPotential.from_circular_velocity (builtin/core.py:743-775): v_c is the circular velocity AT r_ref = r_s
    nfw = gb.NFWPotential.from_circular_velocity(v_c=0.2, r_s=20.0)
    assert np.allclose(nfw.circular_velocity(np.array([[20.0], [0.0], [0.0]])), 0.2, rtol=1e-12)


def test_phase_space_point_quantities(ref):
    pot = gb.MilkyWayPotential2022()
    pot.strict_math = True
    w0 = make_ic(lambda x: pot.gradient(x), 128, seed=1)
    w = gb.PhaseSpacePosition.from_w(w0)
    L = w.angular_momentum()                                       # tests/dynamics/test_dynamics_core.py:366-376
    assert L.shape == (3, 128)
    assert np.allclose(L, np.cross(w0[:3].T, w0[3:].T).T, rtol=0, atol=1e-15)
    T = w.kinetic_energy()
    assert np.allclose(T, 0.5 * (w0[3:] ** 2).sum(0), rtol=1e-15)
    U = w.potential_energy(pot)
    assert np.allclose(U, ref.energy(pot, w0[:3].copy()), rtol=1e-14)
    # the docstring example of angular_momentum (dynamics/core.py:725-733): 1 au, 2 pi au / yr -> Lz = 6.28318531
    one = gb.PhaseSpacePosition(pos=[1.0, 0, 0], vel=[0, 2 * np.pi, 0])
    assert np.allclose(one.angular_momentum(), [0, 0, 6.28318531], atol=1e-8)
    # along an orbit: T + U is the Hamiltonian of the static frame, conserved by leapfrog to the scheme's order
    H = gb.Hamiltonian(pot)
    orbit = H.integrate_orbit(w0, dt=0.5, n_steps=400)
    E = orbit.kinetic_energy() + orbit.potential_energy()
    assert E.shape == (401, 128)
    assert np.allclose(E, orbit.energy(), rtol=1e-13)
    drift = np.abs(E / E[0] - 1).max(0)                           # per orbit; the worst ones pass the 70-pc nucleus at dt = 0.5
    assert np.median(drift) < 1e-3 and drift.max() < 5e-2
    with pytest.raises(ValueError):
        gb.Orbit(orbit.pos, orbit.vel).potential_energy()


def test_circulation_known_orbits_of_the_reference():
    """tests/dynamics/test_orbit.py:484-522 (Binney & Tremaine 2008 fig. 3.8 / 3.9): a loop and a box orbit of a
    flattened logarithmic potential at E = -0.337; also with the trajectory left on the device."""
    torch = pytest.importorskip("torch")
    pot = gb.LogarithmicPotential(v_c=1.0, r_h=0.14, q1=1.0, q2=0.9, q3=1.0)
    E = -0.337
    ws = []
    for x, vx in zip([0.5, 0.0], [0.0, 1.5]):
        vy = np.sqrt(2 * (E - pot.energy(np.array([[x], [0.0], [0.0]]))))[0]
        ws.append([x, 0.0, 0.0, vx, vy, 0.0])
    ws = np.ascontiguousarray(np.array(ws).T)
    H = gb.Hamiltonian(pot)
    orbit = H.integrate_orbit(ws, dt=0.05, n_steps=10000)
    c1 = orbit[:, 0].circulation()
    assert c1.shape == (3,) and c1.sum() == 1
    c2 = orbit[:, 1].circulation()
    assert c2.shape == (3,) and c2.sum() == 0
    circ = orbit.circulation()
    assert circ.shape == (3, 2) and np.array_equal(circ.sum(axis=0), [1, 0])
    dev = H.integrate_orbit(torch.as_tensor(ws, device="cuda"), dt=0.05, n_steps=10000)
    assert dev.pos.is_cuda and np.array_equal(dev.circulation(), circ)
    aligned = dev.align_circulation_with_z()
    assert aligned.pos.is_cuda and np.array_equal(aligned.circulation(), circ)       # already a z-tube and a box


def test_estimate_period_docstring_case_and_kepler(ref):
    """dynamics/orbit.py:697-712: MilkyWayPotential2022, w0 = [8, 0, 0, 0, 0.18, 0], dt = 1, 4000 leapfrog steps.  The
    periods are means of integer peak spacings: the GPU trajectory must give EXACTLY what the compiled reference's
    trajectory gives (175.86038961038963, 175.88419913419915, cylindrical 120.96875 Myr with this repo's G; the
    docstring prints 176.02 / 176.07 / 121.09 -- 0.1 % away, a different parameter / constant set than the source).
    tests/dynamics/test_orbit.py:472-482: three copies of a Kepler orbit in solar-system units through DOPRI853 --
    here also checked against Kepler's third law."""
    from gala_b200.dynamics import peak_to_peak_period
    pot = gb.MilkyWayPotential2022()
    w0 = np.array([8.0, 0, 0, 0, 0.18, 0])
    orbit = pot.integrate_orbit(w0, dt=1.0, n_steps=4000)
    T = orbit.estimate_period(components=("x", "y", "rho"))
    wr = ref.leapfrog(pot, w0.reshape(6, 1), np.arange(4001.0), save_all=True)[:, :, 0]
    want = {"x": peak_to_peak_period(orbit.t, wr[0]), "y": peak_to_peak_period(orbit.t, wr[1]),
            "rho": peak_to_peak_period(orbit.t, np.hypot(wr[0], wr[1]))}
    for k in want:
        assert T[k][0] == want[k], (k, T[k], want[k])
    assert np.allclose([want["x"], want["y"], want["rho"]], [175.86038961038963, 175.88419913419915, 120.96875], rtol=1e-14)
    assert np.allclose([want["x"], want["y"], want["rho"]], [176.02380952380952, 176.07034632034632, 121.09375], rtol=2e-3)
    kep = gb.KeplerPotential(m=1.0, units=gb.solarsystem)
    w0 = np.tile(np.array([1.0, 0, 0, 0, 1.5 * np.pi, 0])[:, None], (1, 3))
    w = gb.Hamiltonian(kep).integrate_orbit(w0, dt=0.01, n_steps=10000, Integrator=gb.DOPRI853Integrator)
    P = w.estimate_period()
    GM = kep.G * 1.0
    a = -GM / (2 * (0.5 * (1.5 * np.pi) ** 2 - GM))
    want = 2 * np.pi * np.sqrt(a ** 3 / GM)
    assert P["x"].shape == (3,) and np.allclose(P["x"], want, rtol=1e-3) and np.allclose(P["y"], want, rtol=1e-3)
    assert np.all(np.isnan(P["z"]))


def test_guiding_radius_like_the_reference():
    """tests/dynamics/test_dynamics_core.py:378-394 through the device gradient, numpy and torch.cuda; the root is
    checked against brentq on the closed-form Hernquist circular velocity, and in MilkyWayPotential2022 by
    substituting it back: R_g v_circ(R_g) = |L_z|."""
    torch = pytest.importorskip("torch")
    from scipy.optimize import brentq
    m, c = 1e11, 10.0
    p = gb.HernquistPotential(m=m, c=c)
    rng = np.random.default_rng(42)
    R = rng.uniform(4, 10, 128)
    xyz = R[None] * np.array([1.0, 0, 0])[:, None]
    vc = p.circular_velocity(xyz)
    vxyz = np.zeros((3, R.size))
    vxyz[1] = rng.normal(vc / gb.KMS_TO_KPC_MYR, 15.0) * gb.KMS_TO_KPC_MYR
    w0 = gb.PhaseSpacePosition(xyz, vxyz)
    Rg = w0.guiding_radius(p)
    assert np.all(Rg > 0) and np.all(Rg < 25)
    want = np.array([brentq(lambda x: L - x * np.sqrt(G * m * x) / (x + c), 1e-3, 1e3, xtol=1e-14, rtol=1e-14)
                     for L in np.abs(R * vxyz[1])])
    assert np.allclose(Rg, want, rtol=1e-10)
    Rg_d = gb.PhaseSpacePosition(torch.as_tensor(xyz, device="cuda"), torch.as_tensor(vxyz, device="cuda")).guiding_radius(p)
    assert Rg_d.is_cuda and np.allclose(Rg_d.cpu().numpy(), Rg, rtol=1e-12)
    mw = gb.MilkyWayPotential2022()
    w = gb.PhaseSpacePosition.from_w(make_ic(lambda x: mw.gradient(x), 300, seed=2))
    Rg = w.guiding_radius(mw)
    q = np.zeros((3, 300)); q[0] = Rg
    assert np.allclose(Rg * mw.circular_velocity(q), np.abs(w.angular_momentum()[2]), rtol=1e-10)


def test_rotating_frame_orbit_transforms_back_to_the_static_one():
    """Orbit.to_frame (dynamics/orbit.py:1256-1296, transformations.py:98-190) closes the loop on the rotating-frame
    kernels: the same initial conditions (frames coincide at t = 0; the momenta are inertial velocities) integrated
    with DOPRI853 in a rotating frame and turned back equal the static-frame orbit -- for potentials that are
    symmetric about the rotation axis (MilkyWayPotential2022 about z; a spherical composite about a tilted axis), so
    that "static in the rotating frame" is the same physical system; numpy and torch.cuda."""
    torch = pytest.importorskip("torch")
    sph = gb.CCompositePotential()
    sph["halo"] = gb.NFWPotential(m=6e11, r_s=16.0)
    sph["bulge"] = gb.HernquistPotential(m=3e10, c=1.0)
    mw = gb.MilkyWayPotential2022()
    static = gb.StaticFrame()
    t = np.linspace(0.0, 500.0, 101)
    for pot, Om in ((mw, [0.0, 0.0, 0.030681]), (sph, [0.004, -0.01, 0.030681])):
        rotating = gb.ConstantRotatingFrame(Om)
        w0 = make_ic(lambda x: pot.gradient(x), 200, seed=6, rmin=6.0, rmax=30.0)
        o_s = gb.Hamiltonian(pot, static).integrate_orbit(w0, t=t, Integrator=gb.DOPRI853Integrator)
        o_r = gb.Hamiltonian(pot, rotating).integrate_orbit(w0, t=t, Integrator=gb.DOPRI853Integrator)
        back = o_r.to_frame(static)
        assert isinstance(back.frame, gb.StaticFrame)
        dp = np.sqrt(((back.pos - o_s.pos) ** 2).sum(0)) / np.sqrt((o_s.pos ** 2).sum(0))
        dv = np.sqrt(((back.vel - o_s.vel) ** 2).sum(0)) / np.sqrt((o_s.vel ** 2).sum(0))
        print(f"\n[rotating -> static, Omega = {Om}] position q50/max = {np.median(dp):.2e} {dp.max():.2e}, "
              f"velocity {np.median(dv):.2e} {dv.max():.2e}")
        # two independent adaptive integrations at rtol = atol = 1e-10: the compiled reference's own pair differs by
        # 2.4e-9 (median) / 7.6e-7 (max) on these initial conditions
        assert np.median(dp) < 1e-8 and dp.max() < 1e-5 and np.median(dv) < 1e-8 and dv.max() < 1e-5
        fwd = o_s.to_frame(rotating)
        assert np.median(np.sqrt(((fwd.pos - o_r.pos) ** 2).sum(0)) / np.sqrt((o_r.pos ** 2).sum(0))) < 1e-8
    o_d = gb.Hamiltonian(pot, rotating).integrate_orbit(torch.as_tensor(w0, device="cuda"), t=t, Integrator=gb.DOPRI853Integrator)
    back_d = o_d.to_frame(static)
    assert back_d.pos.is_cuda and np.allclose(back_d.pos.cpu().numpy(), back.pos, rtol=0, atol=1e-12)


def test_estimate_dt_n_steps_like_the_reference():
    """tests/dynamics/test_dynamics_util.py:34-48: NFW from v_c = 1 at r_s = 10, w0 = [10, 0, 0, 0, 0.9, 0]; the
    recommended (dt, n_steps) covers exactly the requested 128 radial periods."""
    nperiods = 128
    pot = gb.NFWPotential.from_circular_velocity(v_c=1.0, r_s=10.0)
    w0 = [10.0, 0.0, 0.0, 0.0, 0.9, 0.0]
    H = gb.Hamiltonian(pot)
    dt, n_steps = gb.estimate_dt_n_steps(w0, H, n_periods=nperiods, n_steps_per_period=256, func=np.nanmin)
    assert n_steps == nperiods * 256
    orbit = H.integrate_orbit(np.array(w0), dt=dt, n_steps=n_steps)
    T = orbit.estimate_period(components=("r",))["r"]
    assert int(np.squeeze(np.round(orbit.t.max() / T))) == nperiods
    # a potential instead of a Hamiltonian, and no energy criterion (dt = 1 test orbit)
    dt2, n2 = gb.estimate_dt_n_steps(w0, pot, n_periods=4, n_steps_per_period=100, dE_threshold=None, func=np.nanmin)
    assert n2 == 400 and abs(dt2 * 100 / (dt * 256) - 1) < 2e-2          # the same period from either test orbit


def test_surface_of_section_like_the_reference():
    """tests/dynamics/test_nonlinear.py:312-320: triaxial logarithmic potential, one orbit, section y = 0 with
    p_y > 0; every recorded sample is the closest one to the plane on its crossing."""
    pot = gb.LogarithmicPotential(v_c=1.0, r_h=1.0, q1=1.0, q2=0.9, q3=0.8)
    w0 = np.array([0.0, 0.8, 0.0, 1.0, 0.0, 0.0])
    H = gb.Hamiltonian(pot)
    orbit = H.integrate_orbit(w0, dt=0.02, n_steps=100_000)
    sos = gb.surface_of_section(orbit, constant_idx=1)
    n = sos.pos.shape[1]
    assert sos.pos.shape == (3, n) and n > 100
    assert np.all(sos.vel[1] > 0) and np.abs(sos.pos[1]).max() < 0.02 * np.abs(orbit.vel[1]).max()
    many = H.integrate_orbit(np.stack([w0, w0 * np.array([1, 1.1, 1, 1, 1, 1])], axis=1), dt=0.02, n_steps=20_000)
    secs = gb.surface_of_section(many, constant_idx=1)
    assert len(secs) == 2 and np.array_equal(secs[0].pos, gb.surface_of_section(many[:, 0], constant_idx=1).pos)
