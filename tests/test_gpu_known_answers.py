"""The reference's own known answers that need no astropy, evaluated THROUGH THE GPU PATH (the same ones pin the
oracles on the CPU in test_oracle_cpu.py): the Plummer doctest numbers, a Kepler orbit closing after one
period, a circular orbit at rest in the co-rotating frame, conservation of the Jacobi constant, the Cython ==
Python integrator set-up, and the sympy closed forms of builtin/core.py."""
import numpy as np
import pytest

import gala_b200 as gb
from conftest import make_ic

pytestmark = pytest.mark.gpu
G = gb.G_GALACTIC


def test_plummer_doctest_numbers_gpu():
    """docs/dynamics/orbits-in-detail.rst:240-243,321-330: Plummer(m=1e10, b=1), w0 = [10,0,0] kpc, [0,75,0] km/s,
    default (leapfrog) integrator, dt = 0.1, 1e5 steps: <Lz> = 0.76703412, pericentre 10.00000005952518 kpc,
    apocentre 19.390916871970223 kpc, eccentricity 0.31951765618193967."""
    pot = gb.PlummerPotential(m=1e10, b=1.0)
    w0 = gb.PhaseSpacePosition(pos=[10.0, 0, 0], vel=[0, 75 * gb.KMS_TO_KPC_MYR, 0])
    for strict in (True, False):
        pot.strict_math = strict
        orbit = pot.integrate_orbit(w0, dt=0.1, n_steps=100000)
        w = np.vstack([orbit.pos, orbit.vel])
        r = np.sqrt((w[:3] ** 2).sum(0))

        def extremum(sign):          # parabolic refinement, like the interpolation of Orbit.apocenter()
            i = np.argmax(sign * r[1:-1]) + 1
            y0, y1, y2 = r[i - 1], r[i], r[i + 1]
            return y1 - 0.125 * (y2 - y0) ** 2 / (y2 - 2 * y1 + y0)
        peri, apo = r.min(), extremum(+1)
        assert abs(peri - 10.00000005952518) < 1e-6
        assert abs(apo - 19.390916871970223) < 2e-5
        assert abs((apo - peri) / (apo + peri) - 0.31951765618193967) < 2e-6
        Lz = w[0] * w[4] - w[1] * w[3]
        assert np.allclose(Lz, Lz[0], rtol=1e-11) and abs(Lz.mean() - 0.76703412) < 1e-8
    pot.strict_math = False


def test_kepler_orbit_closes_gpu():
    """tests/integrate/test_pyintegrators.py:96-104: a Kepler orbit returns to its start after one period."""
    M, a = 1e11, 10.0
    pot = gb.KeplerPotential(m=M)
    vc = np.sqrt(G * M / a)
    w0 = np.array([a, 0, 0, 0, 0.8 * vc, 0.0])
    E = 0.5 * (0.8 * vc) ** 2 - G * M / a
    T = 2 * np.pi * np.sqrt((-G * M / (2 * E)) ** 3 / (G * M))
    H = gb.Hamiltonian(pot)
    for integ, n, tol in (("leapfrog", 20000, 1e-5), ("ruth4", 4000, 1e-6), ("dopri853", 200, 1e-8)):
        end = H.integrate_orbit(w0, Integrator=integ, t=np.linspace(0, T, n + 1), save_all=False)
        assert np.allclose(end.pos.ravel(), w0[:3], atol=tol * a), integ


def test_rotating_frame_known_answers_gpu():
    """tests/potential/hamiltonian/test_with_frame_potential.py:130-177: a circular Kepler orbit is stationary in
    the co-rotating frame (atol 1e-7); the Jacobi constant is conserved to < 1e-9 at DOP853 rtol = atol = 1e-12."""
    M, r0 = 1e11, 8.0
    vc = np.sqrt(G * M / r0)
    H = gb.Hamiltonian(gb.KeplerPotential(m=M), gb.ConstantRotatingFrame([0.0, 0.0, vc / r0]))
    w0 = np.array([r0, 0, 0, 0, vc, 0.0])
    t = np.linspace(0, 1000, 201)
    orb = H.integrate_orbit(w0, Integrator="dopri853", t=t, Integrator_kwargs=dict(atol=1e-12, rtol=1e-12))
    assert np.allclose(orb.pos, w0[:3, None], atol=1e-7)
    pot2 = gb.MilkyWayPotential2022()
    H2 = gb.Hamiltonian(pot2, gb.ConstantRotatingFrame([0.0, 0.0, 0.030681]))
    w = make_ic(lambda q: pot2.gradient(q), 64, seed=3)
    orb = H2.integrate_orbit(w, Integrator="dopri853", t=t, Integrator_kwargs=dict(atol=1e-12, rtol=1e-12))
    EJ = orb.energy()
    assert np.max(np.abs(EJ / EJ[0] - 1)) < 1e-9


@pytest.mark.parametrize("name", ["hernquist", "mn", "nfw", "bar", "plummer", "isochrone", "jaffe", "kepler", "stone",
                                  "satoh", "kuzmin", "logarithmic"])
def test_against_sympy_closed_forms_gpu(name):
    """potential_helpers.py:409-503 (test_against_sympy): energy, gradient, density (Poisson) and Hessian of the
    DEVICE functions against the sympy expressions of builtin/core.py at 64 random points."""
    import sympy as sy
    from test_oracle_cpu import sympy_case
    pot, f, g, lap = sympy_case(name)
    q = np.random.default_rng(42).uniform(-10, 10, (3, 64)) + 0.1
    for strict in (True, False):
        pot.strict_math = strict
        assert np.allclose(pot.energy(q), f(*q), rtol=1e-11)
        assert np.allclose(pot.gradient(q), np.array(g(*q)), rtol=1e-9, atol=1e-30)
        if name not in ("nfw", "kepler", "kuzmin"):
            dens = lap(*q) / (4 * np.pi * G)
            assert np.allclose(pot.density(q), dens, rtol=2e-6, atol=1e-8 * np.abs(dens).max())
            H = pot.hessian(q)
            assert np.allclose(H[0, 0] + H[1, 1] + H[2, 2], lap(*q), rtol=1e-8, atol=1e-10 * np.abs(lap(*q)).max())
    pot.strict_math = False


def test_bar_rotating_frame_jacobi_energy_gpu():
    """tests/integration/test_bar_rotating_frame.py:171-219 (test_energy_conservation), the half that needs no
    TimeInterpolatedPotential: LongMuraliBar + MW(v2) disk (m = 4.1e10) + halo + nucleus, static in a frame rotating
    at 30 km/s/kpc, an orbit started at corotation, DOP853 atol = rtol = 1e-14, dt = 0.1 Myr over ~5 Gyr: the Jacobi
    energy is conserved to < 1e-12."""
    mw = gb.MilkyWayPotential(version="latest")
    pot = gb.CCompositePotential()
    pot["bar"] = gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=np.deg2rad(25.0))
    pot["disk"] = gb.MN3ExponentialDiskPotential(m=4.1e10, h_R=mw["disk"].parameters["h_R"], h_z=mw["disk"].parameters["h_z"])
    pot["halo"] = mw["halo"]
    pot["nucleus"] = mw["nucleus"]
    Omega = 30.0 * gb.KMS_TO_KPC_MYR                                  # rad / Myr
    H = gb.Hamiltonian(pot, gb.ConstantRotatingFrame([0.0, 0.0, Omega]))

    def om(r):
        g = pot.gradient(np.array([[r], [0.0], [0.0]]))[0, 0]
        return np.sqrt(g / r)
    lo, hi = 2.0, 30.0
    for _ in range(80):                                              # Omega_c(r) decreases outward
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if om(mid) > Omega else (lo, mid)
    r_c = 0.5 * (lo + hi)
    w0 = np.array([r_c, 0, 0, 0, Omega * r_c, 0.0])
    period = 2 * np.pi / Omega
    t_end = np.arange(0, 5000.0, period / 200)[-1]
    orb = H.integrate_orbit(w0, t1=0.0, t2=t_end, dt=0.1, Integrator="dopri853",
                            Integrator_kwargs={"atol": 1e-14, "rtol": 1e-14})
    E = orb.energy()
    frac = np.abs((E[1:] - E[0]) / E[0])
    print(f"\n[bar, rotating frame] corotation radius {r_c:.4f} kpc, {len(E)} outputs, max |dE_J/E_J| = {frac.max():.2e}")
    assert frac.max() < 1e-12


def _all_pots():
    from test_gpu_parity import POTS
    return POTS


@pytest.mark.parametrize("name", list(_all_pots()))
def test_gradient_is_derivative_of_energy_gpu(name):
    """PotentialTestBase (tests/potential/potential/potential_helpers.py:265-299): the device gradient against
    finite differences of the device energy for EVERY potential configuration of the parity suite (rtol 1e-5
    like the reference), and orbit integration smoke test of :366-396 (energy bounded along a 1000-step orbit)."""
    pot = _all_pots()[name]
    q = np.random.default_rng(21).normal(0, 9.0, (3, 257))
    if name == "kuzmin":
        q[2] = np.abs(q[2]) + 0.5
    g = pot.gradient(q)
    scale = np.sqrt((g ** 2).sum(0)).max()
    h = 1e-4
    for k in range(3):
        dq = np.zeros_like(q); dq[k] = h
        fd = (-pot.energy(q + 2 * dq) + 8 * pot.energy(q + dq) - 8 * pot.energy(q - dq) + pot.energy(q - 2 * dq)) / (12 * h)
        assert np.allclose(fd, g[k], rtol=1e-5, atol=1e-7 * scale), (name, k)
    H = gb.Hamiltonian(pot)
    w0 = make_ic(lambda qq: pot.gradient(qq), 64, seed=5, rmin=8.0, rmax=30.0)
    orb = H.integrate_orbit(w0, dt=0.5, n_steps=1000)
    E = orb.energy()
    # a bare multipole field binds nothing (an "inner" r^l expansion even diverges outward): orbits leave;
    # the razor-thin Kuzmin disc is not differentiable at z = 0, a fixed-step scheme jumps in energy there
    if not name.startswith("multipole") and name != "kuzmin":
        assert np.all(np.isfinite(E)) and np.all(np.isfinite(orb.pos))
        scale = 0.5 * (w0[3:] ** 2).sum(0) + np.abs(pot.energy(w0[:3]))
        assert np.max(np.abs(E[-1] - E[0]) / scale) < 5e-3
