"""`-m "not gpu"`: pins the ORACLES before anything is compared with them.

* oracle/_ref/libgala_ref.so  = the reference's own C++ (compiled unmodified) + restated Cython loops
* oracle/_ref/libgala_port.so = oracle/port.c, our plain-C restatement

against the reference's own known answers that need no astropy: the sympy closed forms of
builtin/core.py (tests/potential/potential/potential_helpers.py:409-503), the NFW enclosed-mass
identity (test_all_builtin.py:328-341), Kepler / co-rotating-frame / Jacobi-constant tests
(tests/potential/hamiltonian/test_with_frame_potential.py:130-177), SCF == Hernquist
(tests/potential/scf/test_class.py:27-57), the Fortran SCF golden vectors
(tests/potential/scf/test_accp_fortran.py) and the Plummer doctest numbers
(docs/dynamics/orbits-in-detail.rst:321-330).  The port is then pinned against the compiled reference.
"""
import os
import re

import numpy as np
import pytest

import gala_b200 as gb
from conftest import make_ic, relnorm
from oracle import oracle

G = gb.G_GALACTIC
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def port():
    if not os.path.exists(os.path.join(oracle._REF_DIR, "libgala_port.so")):
        oracle.build()
    return oracle.Port()


def checkers(ref, port):
    return [("ref", ref), ("port", port)]


# ---- closed forms (the reference's to_sympy expressions, builtin/core.py) -----------------------------
def sympy_case(name):
    import sympy as sy
    x, y, z = sy.symbols("x y z", real=True)
    if name == "hernquist":       # builtin/core.py:235-242
        pot = gb.HernquistPotential(m=1e11, c=0.5); m, c = 1e11, 0.5
        expr = -G * m / (sy.sqrt(x ** 2 + y ** 2 + z ** 2) + c)
    elif name == "mn":            # :554-563
        pot = gb.MiyamotoNagaiPotential(m=6.8e10, a=3.0, b=0.28); m, a, b = 6.8e10, 3.0, 0.28
        expr = -G * m / sy.sqrt(x ** 2 + y ** 2 + (a + sy.sqrt(z ** 2 + b ** 2)) ** 2)
    elif name == "nfw":           # :743-756
        pot = gb.NFWPotential(m=6e11, r_s=15.0, a=1.0, b=0.9, c=0.8); m, rs = 6e11, 15.0
        uu = sy.sqrt((x / 1.0) ** 2 + (y / 0.9) ** 2 + (z / 0.8) ** 2) / rs
        expr = -G * m / rs * sy.log(1 + uu) / uu
    elif name == "bar":           # :1049-1072
        al = 0.4
        pot = gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=al); m, a, b, c = 1e10, 4.0, 0.8, 0.25
        xx = x * sy.cos(al) + y * sy.sin(al); yy = -x * sy.sin(al) + y * sy.cos(al)
        Tm = sy.sqrt((a - xx) ** 2 + yy ** 2 + (b + sy.sqrt(c ** 2 + z ** 2)) ** 2)
        Tp = sy.sqrt((a + xx) ** 2 + yy ** 2 + (b + sy.sqrt(c ** 2 + z ** 2)) ** 2)
        expr = G * m / (2 * a) * sy.log((xx - a + Tm) / (xx + a + Tp))
    elif name == "plummer":
        pot = gb.PlummerPotential(m=1e11, b=1.5)
        expr = -G * 1e11 / sy.sqrt(x ** 2 + y ** 2 + z ** 2 + 1.5 ** 2)
    elif name == "isochrone":
        pot = gb.IsochronePotential(m=1e11, b=1.5)
        expr = -G * 1e11 / (1.5 + sy.sqrt(x ** 2 + y ** 2 + z ** 2 + 1.5 ** 2))
    elif name == "jaffe":
        pot = gb.JaffePotential(m=1e11, c=2.0)
        r = sy.sqrt(x ** 2 + y ** 2 + z ** 2)
        expr = -G * 1e11 / 2.0 * sy.log(1 + 2.0 / r)
    elif name == "kepler":
        pot = gb.KeplerPotential(m=1e11)
        expr = -G * 1e11 / sy.sqrt(x ** 2 + y ** 2 + z ** 2)
    elif name == "stone":         # builtin/core.py:333-343
        pot = gb.StonePotential(m=1e11, r_c=0.5, r_h=20.0); m, rc, rh = 1e11, 0.5, 20.0
        r = sy.sqrt(x ** 2 + y ** 2 + z ** 2)
        A = -2 * G * m / (np.pi * (rh - rc))
        expr = A * (rh / r * sy.atan(r / rh) - rc / r * sy.atan(r / rc) + sy.log((r ** 2 + rh ** 2) / (r ** 2 + rc ** 2)) / 2)
    elif name == "satoh":         # :485-492
        pot = gb.SatohPotential(m=8e10, a=3.0, b=0.4); m, a, b = 8e10, 3.0, 0.4
        expr = -G * m / sy.sqrt(x ** 2 + y ** 2 + z ** 2 + a * (a + 2 * sy.sqrt(z ** 2 + b ** 2)))
    elif name == "kuzmin":        # :519-524 (test points stay off the z = 0 sheet)
        pot = gb.KuzminPotential(m=8e10, a=3.0); m, a = 8e10, 3.0
        expr = -G * m / sy.sqrt(x ** 2 + y ** 2 + (a + sy.Abs(z)) ** 2)
    elif name == "logarithmic":   # :974-979 (the sympy form has no phi rotation)
        pot = gb.LogarithmicPotential(v_c=0.2, r_h=12.0, q1=1.38, q2=1.0, q3=1.36)
        expr = 0.2 ** 2 / 2 * sy.log(12.0 ** 2 + (x / 1.38) ** 2 + y ** 2 + (z / 1.36) ** 2)
    f = sy.lambdify((x, y, z), expr, "numpy")
    g = sy.lambdify((x, y, z), [sy.diff(expr, v) for v in (x, y, z)], "numpy")
    lap = sy.lambdify((x, y, z), sum(sy.diff(expr, v, 2) for v in (x, y, z)), "numpy")
    return pot, f, g, lap


@pytest.mark.parametrize("name", ["hernquist", "mn", "nfw", "bar", "plummer", "isochrone", "jaffe", "kepler",
                                  "stone", "satoh", "kuzmin", "logarithmic"])
def test_against_sympy_closed_forms(ref, port, name):
    """Energy, gradient and density (via Poisson) of both oracles against the reference's sympy
    definitions at 64 random points (potential_helpers.py:409-503 uses rtol 1e-5; here 1e-9)."""
    pot, f, g, lap = sympy_case(name)
    q = np.random.default_rng(42).uniform(-10, 10, (3, 64)) + 0.1
    for label, chk in checkers(ref, port):
        assert np.allclose(chk.energy(pot, q), f(*q), rtol=1e-11), (label, name)
        assert np.allclose(chk.gradient(pot, q), np.array(g(*q)), rtol=1e-9, atol=1e-30), (label, name)
        if name not in ("nfw", "kepler", "kuzmin"):      # triaxial NFW has no density in the reference (nan_density)
            dens = lap(*q) / (4 * np.pi * G)
            assert np.allclose(chk.density(pot, q), dens, rtol=2e-6, atol=1e-8 * np.abs(dens).max()), (label, name)


def test_burkert_powerlawcutoff_leesuto_known_answers(ref, port):
    """Potentials without a to_sympy form in the reference, pinned through textbook identities:
    Burkert and PowerLawCutoff are spherical, so dPhi/dr = G M(<r) / r^2 with the analytic enclosed mass
    (Mori & Burkert 2000 eq. 3; M(<r) = m P((3-alpha)/2, r^2/r_c^2) with scipy's regularised incomplete
    gamma function, which also pins the GSL stand-in); Lee & Suto: the gradient is the derivative of the
    value (central differences) and the spherical limit a=b=c is an NFW profile."""
    from scipy.special import gammainc
    r = np.geomspace(0.05, 200, 60)
    q = np.vstack([r * 0.6, r * 0.0, r * 0.8])
    rho, r0 = 2.3e7, 4.0
    bk = gb.BurkertPotential(rho=rho, r0=r0)
    xx = r / r0
    m_bk = np.pi * rho * r0 ** 3 * (np.log(1 + xx ** 2) + 2 * np.log(1 + xx) - 2 * np.arctan(xx))
    m, al, rc = 4.5e9, 1.8, 1.9
    pl = gb.PowerLawCutoffPotential(m=m, alpha=al, r_c=rc)
    m_pl = m * gammainc(0.5 * (3 - al), (r / rc) ** 2)
    for label, chk in checkers(ref, port):
        for pot, menc, tol in ((bk, m_bk, 1e-9), (pl, m_pl, 1e-12)):
            g = chk.gradient(pot, q)
            dphi_dr = (g * q).sum(0) / r
            assert np.allclose(dphi_dr * r ** 2 / G, menc, rtol=tol), (label, type(pot).__name__)
            # value <-> gradient consistency
            h = 1e-5 * r
            num = (chk.energy(pot, q * (1 + h / r)) - chk.energy(pot, q * (1 - h / r))) / (2 * h)
            assert np.allclose(num, dphi_dr, rtol=2e-6), (label, type(pot).__name__)
        assert np.allclose(chk.density(bk, q), rho / ((1 + xx) * (1 + xx ** 2)), rtol=1e-13)
        from scipy.special import gamma as Gam
        assert np.allclose(chk.density(pl, q), m / (2 * np.pi) * rc ** (al - 3) / Gam(0.5 * (3 - al)) * r ** -al * np.exp(-(r / rc) ** 2), rtol=1e-12)
        ls = gb.LeeSutoTriaxialNFWPotential(v_c=0.2, r_s=15.0, a=1.0, b=0.85, c=0.7)
        qq = np.random.default_rng(3).normal(0, 12.0, (3, 64))
        g = chk.gradient(ls, qq)
        for k in range(3):
            dq = np.zeros_like(qq); dq[k] = 1e-4
            num = (chk.energy(ls, qq + dq) - chk.energy(ls, qq - dq)) / 2e-4
            assert np.allclose(num, g[k], rtol=1e-5, atol=1e-7 * np.abs(g).max()), (label, k)
        sph = gb.LeeSutoTriaxialNFWPotential(v_c=0.2, r_s=15.0)
        # spherical limit: Phi = -v_h^2 ln(1+u)/u with v_h^2 = v_c^2 / (ln 2 - 1/2)
        u = np.sqrt((qq ** 2).sum(0)) / 15.0
        assert np.allclose(chk.energy(sph, qq), -(0.2 ** 2 / (np.log(2.) - 0.5)) * np.log(1 + u) / u, rtol=1e-13), label


def test_nfw_enclosed_mass_identity(ref, port):
    """tests/potential/potential/test_all_builtin.py:328-341: M(<r) from dPhi/dr vs the analytic NFW mass."""
    m, rs = 6e11, 15.0
    pot = gb.NFWPotential(m=m, r_s=rs)
    r = np.geomspace(0.1, 300, 50)
    q = np.vstack([r, 0 * r, 0 * r])
    analytic = m * (np.log(1 + r / rs) - (r / rs) / (1 + r / rs))
    for label, chk in checkers(ref, port):
        menc = r ** 2 * chk.gradient(pot, q)[0] / G
        assert np.allclose(menc, analytic, rtol=1e-12), label


def test_mw2022_is_sum_of_components(ref, port):
    pot = gb.MilkyWayPotential2022()
    q = np.random.default_rng(1).normal(0, 10, (3, 100))
    for label, chk in checkers(ref, port):
        tot = chk.gradient(pot, q)
        parts = sum(chk.gradient(p, q) for p in pot.values())
        assert np.allclose(tot, parts, rtol=1e-13)
        three = sum(chk.gradient(p, q) for p in pot["disk"].get_three_potentials().values())
        assert np.allclose(chk.gradient(pot["disk"], q), three, rtol=1e-13)


def test_kepler_orbit_closes(ref, port):
    """tests/integrate/test_pyintegrators.py:96-104: a Kepler orbit returns to its start after one period."""
    M = 1e11
    pot = gb.KeplerPotential(m=M)
    H = gb.Hamiltonian(pot)
    a = 10.0
    vc = np.sqrt(G * M / a)
    w0 = np.array([[a, 0, 0, 0, 0.8 * vc, 0.0]]).T          # eccentric, still bound
    E = 0.5 * (0.8 * vc) ** 2 - G * M / a
    a_orb = -G * M / (2 * E)
    T = 2 * np.pi * np.sqrt(a_orb ** 3 / (G * M))
    for label, chk in checkers(ref, port):
        t = np.linspace(0, T, 20001)
        w = chk.leapfrog(pot, w0, t, save_all=False)
        assert np.allclose(w[:, 0], w0[:, 0], atol=1e-5 * a), label
        w = chk.ruth4(H, w0, t[::10].copy(), save_all=False)
        assert np.allclose(w[:, 0], w0[:, 0], atol=1e-6 * a), label
        wd, st, rc = chk.dop853(H, w0, np.array([0.0, T / 2, T]), atol=1e-12, rtol=1e-12, nbatch=1)
        assert rc >= 0 and np.all(st == 1)
        assert np.allclose(wd[:, -1, 0], w0[:, 0], atol=1e-8 * a), label


def test_rotating_frame_known_answers(ref, port):
    """tests/potential/hamiltonian/test_with_frame_potential.py:130-177: a circular Kepler orbit is
    stationary in the co-rotating frame (atol 1e-7) and the Jacobi constant is conserved (< 1e-9 at
    DOP853 rtol=atol=1e-12)."""
    M = 1e11
    pot = gb.KeplerPotential(m=M)
    r0 = 8.0
    vc = np.sqrt(G * M / r0)
    Om = vc / r0
    H = gb.Hamiltonian(pot, gb.ConstantRotatingFrame([0.0, 0.0, Om]))
    w0 = np.array([[r0, 0, 0, 0, vc, 0.0]]).T            # canonical momentum p = v_inertial
    t = np.linspace(0, 1000, 201)
    for label, chk in checkers(ref, port):
        w, st, rc = chk.dop853(H, w0, t, atol=1e-12, rtol=1e-12, nbatch=1)
        assert rc >= 0
        assert np.allclose(w[:3, :, 0], w0[:3], atol=1e-7), label
    pot2 = gb.MilkyWayPotential2022()
    H2 = gb.Hamiltonian(pot2, gb.ConstantRotatingFrame([0.0, 0.0, 0.030681]))
    w0 = make_ic(lambda q: ref.gradient(pot2, q), 16, seed=3)
    for label, chk in checkers(ref, port):
        w, st, rc = chk.dop853(H2, w0, t, atol=1e-12, rtol=1e-12, nbatch=1)
        EJ = chk.hamiltonian_energy(H2, w.reshape(6, -1)).reshape(len(t), -1)
        assert np.max(np.abs(EJ / EJ[0] - 1)) < 1e-9, label


def test_plummer_doctest_numbers(ref, port):
    """docs/dynamics/orbits-in-detail.rst:240-243,321-330: Plummer(m=1e10, b=1), w0 = [10,0,0] kpc,
    [0,75,0] km/s, default (leapfrog) integrator, dt=0.1, 1e5 steps: <Lz> = 0.76703412 kpc^2/Myr,
    pericentre 10.00000005952518 kpc, apocentre 19.390916871970223 kpc, e = 0.31951765618193967."""
    pot = gb.PlummerPotential(m=1e10, b=1.0)
    w0 = np.array([[10.0, 0, 0, 0, 75 * gb.KMS_TO_KPC_MYR, 0]]).T
    t = 0.1 * np.arange(100001)
    for label, chk in checkers(ref, port):
        w = chk.leapfrog(pot, w0, t, save_all=True)
        r = np.sqrt((w[:3, :, 0] ** 2).sum(0))
        # parabolic refinement of the extrema like Orbit.pericenter()/apocenter() do by interpolation
        def extremum(sign):
            i = np.argmax(sign * r[1:-1]) + 1
            y0, y1, y2 = r[i - 1], r[i], r[i + 1]
            return y1 - 0.125 * (y2 - y0) ** 2 / (y2 - 2 * y1 + y0)
        peri, apo = r.min(), extremum(+1)
        assert abs(peri - 10.00000005952518) < 1e-6, (label, peri)
        assert abs(apo - 19.390916871970223) < 2e-5, (label, apo)
        assert abs((apo - peri) / (apo + peri) - 0.31951765618193967) < 2e-6, label
        Lz = w[0, :, 0] * w[4, :, 0] - w[1, :, 0] * w[3, :, 0]
        assert np.allclose(Lz, Lz[0], rtol=1e-12)
        assert abs(Lz.mean() - 0.76703412) < 1e-8


# ---- SCF ------------------------------------------------------------------------------------------------
SCF_SETS = ["simple_hernquist", "multi_hernquist", "simple_nonsph", "random", "wang_zhao"]


@pytest.mark.parametrize("name", SCF_SETS)
def test_scf_fortran_golden_vectors(ref, port, name):
    """tests/potential/scf/test_accp_fortran.py:77-156 on the committed fixture
    (tests/golden/scf_fortran.npz, made by tests/golden/make_scf_golden.py): potential and gradient at
    rtol 1e-6 with G = M = r_s = 1."""
    d = np.load(os.path.join(GOLD, "scf_fortran.npz"))
    pot = gb.SCFPotential(m=1.0, r_s=1.0, Snlm=d[name + "_S"], Tnlm=d[name + "_T"], units=gb.dimensionless)
    assert pot.G == 1.0
    q = np.ascontiguousarray(d["xyz"].T)
    for label, chk in checkers(ref, port):
        np.testing.assert_allclose(chk.energy(pot, q), d[name + "_pot"], rtol=1e-6, err_msg=label)
        np.testing.assert_allclose(chk.gradient(pot, q).T, d[name + "_grad"], rtol=1e-6, err_msg=label)


def test_scf_s000_is_hernquist(ref, port):
    """tests/potential/scf/test_class.py:27-57: SCF with S000 = 1 is the Hernquist sphere."""
    S = np.zeros((4, 3, 3)); S[0, 0, 0] = 1.0
    scf = gb.SCFPotential(m=1e10, r_s=3.0, Snlm=S)
    hern = gb.HernquistPotential(m=1e10, c=3.0)
    r = np.geomspace(0.05, 100, 128)
    rng = np.random.default_rng(0)
    n = rng.normal(size=(3, 128)); n /= np.sqrt((n ** 2).sum(0))
    q = r * n
    for label, chk in checkers(ref, port):
        np.testing.assert_allclose(chk.energy(scf, q), chk.energy(hern, q), rtol=1e-12, err_msg=label)
        np.testing.assert_allclose(chk.gradient(scf, q), chk.gradient(hern, q), rtol=1e-9, atol=1e-25, err_msg=label)
        np.testing.assert_allclose(chk.density(scf, q), chk.density(hern, q), rtol=1e-11, err_msg=label)


# ---- the port against the compiled reference -------------------------------------------------------------
def all_potentials():
    from test_gpu_parity import potentials
    return potentials()


def test_port_equals_reference_evaluation(ref, port):
    rng = np.random.default_rng(7)
    q = rng.normal(0, 10.0, (3, 2000))
    for name, pot in all_potentials().items():
        g, g0 = port.gradient(pot, q), ref.gradient(pot, q)
        loose = name.startswith(('scf', 'multipole')) or name in ("powerlawcutoff", "bovy2014")   # series on both sides
        assert np.max(np.sqrt(((g - g0) ** 2).sum(0)) / np.sqrt((g0 ** 2).sum(0))) < (1e-13 if loose else 4e-16), name
        assert np.allclose(port.energy(pot, q), ref.energy(pot, q), rtol=(1e-12 if loose else 1e-15), atol=0), name
        d, d0 = port.density(pot, q), ref.density(pot, q)
        ok = np.isfinite(d0)
        assert np.array_equal(np.isfinite(d), ok)
        if ok.any():
            tol = 1e-8 if "bar" in name else (1e-12 if name.startswith("scf") else 1e-14)
            assert np.max(np.abs(d[ok] - d0[ok])) <= tol * np.abs(d0[ok]).max(), name
    H = gb.Hamiltonian(all_potentials()["bar_mw2022"], gb.ConstantRotatingFrame([0.001, 0.002, 0.03]))
    w = rng.normal(0, 8.0, (6, 500)); w[3:] *= 0.02
    assert np.allclose(port.hamiltonian_energy(H, w), ref.hamiltonian_energy(H, w), rtol=1e-15)
    assert np.allclose(port.hamiltonian_gradient(H, w), ref.hamiltonian_gradient(H, w), rtol=1e-13, atol=1e-18)
    assert np.isclose(port.d2_dr2(H.potential, w[:3, 0]), ref.d2_dr2(H.potential, w[:3, 0]), rtol=1e-12)


def test_port_equals_reference_integrators(ref, port):
    pots = all_potentials()
    pot = pots["mw2022"]
    w0 = make_ic(lambda q: ref.gradient(pot, q), 300, seed=1, rmin=15.0)
    t = np.arange(1001, dtype=float)
    d = relnorm(port.leapfrog(pot, w0, t), ref.leapfrog(pot, w0, t))
    assert d.max() < 1e-12 and np.median(d) < 1e-15
    for frame in (gb.StaticFrame(), gb.ConstantRotatingFrame([0, 0, 0.030681])):
        H = gb.Hamiltonian(pots["bar_mw2022"], frame)
        d = relnorm(port.ruth4(H, w0, t * 0.5, save_all=False), ref.ruth4(H, w0, t * 0.5, save_all=False))
        assert d.max() < 1e-12
        H = gb.Hamiltonian(pot, frame)
        tt = np.linspace(0, 1000, 200)
        wp, sp, rcp = port.dop853(H, w0[:, :100], tt, nbatch=1)
        wr, sr, rcr = ref.dop853(H, w0[:, :100], tt, nbatch=1)
        assert rcp >= 0 and rcr >= 0 and np.array_equal(sp, sr)
        assert relnorm(wp, wr).max() < 1e-10
        rows = np.ascontiguousarray(w0[:, :50].T)
        a, sa, _ = port.dop853_step_rows(H, rows, 0.0, 800.0, 1.0)
        b, sb, _ = ref.dop853_step_rows(H, rows, 0.0, 800.0, 1.0, group=True)
        assert np.array_equal(sa, sb) and relnorm(a.T, b.T).max() < 1e-10


def test_long_double_truth_run(ref):
    """The long-double build of the port arbitrates FP64 differences: the strict reference stays
    within ~1e-12 of it over 1000 MW2022 leapfrog steps on regular orbits."""
    ld = oracle.Port(long_double=True)
    pot = gb.MilkyWayPotential2022()
    w0 = make_ic(lambda q: ref.gradient(pot, q), 200, seed=5, rmin=15.0)
    t = np.arange(1001, dtype=float)
    d = relnorm(ref.leapfrog(pot, w0, t, save_all=False), ld.leapfrog(pot, w0, t, save_all=False))
    assert np.median(d) < 1e-13 and d.max() < 1e-10


def test_dop853_coefficients_match_reference():
    """Our device / port coefficient tables against the numbers in the reference source, numerically."""
    ref_src = "/root/reference/src/gala/integrate/cyintegrators/dopri/dop853.cpp"
    if not os.path.exists(ref_src):
        pytest.skip("reference tree not present")
    txt = open(ref_src).read()
    txt = txt[txt.index("case 1:"):txt.index("facold = 1.0E-4")]
    refc = {k: float(v) for k, v in re.findall(r"\b([a-z]+[0-9]+)\s*=\s*([-+0-9.Ee]+);", txt)}
    root = os.path.dirname(GOLD.rstrip("/")).rsplit("/tests", 1)[0]
    for path in ("gala_b200/csrc/dop853_coeffs.cuh", "oracle/port_dop853_coeffs.h"):
        mine = {k: float(v) for k, v in re.findall(r"\b([a-z]+[0-9]+)\s*=\s*([-+0-9.Ee]+)", open(os.path.join(root, path)).read())}
        assert set(mine) == set(refc) and len(refc) == 154
        assert all(mine[k] == refc[k] for k in refc), path


# ---- multipole expansion (builtin/multipole.cpp) ---------------------------------------------------
def _mp_random(lmax, inner, seed):
    rng = np.random.default_rng(seed)
    kw = {}
    for l in range(lmax + 1):
        for m in range(l + 1):
            kw[f"S{l}{m}"] = rng.normal()
            if m > 0:
                kw[f"T{l}{m}"] = rng.normal()
    return gb.MultipolePotential(lmax=lmax, inner=inner, m=3e10, r_s=7.0, **kw)


def test_multipole_closed_forms(ref, port):
    """Known answers for the lowest orders, derived from Y_lm with the Condon-Shortley phase:
    outer l=0 with S00 = -sqrt(4 pi) is a Kepler potential; inner (l,m)=(1,0) is a uniform field along z;
    inner (2,2) is proportional to x^2 - y^2."""
    rng = np.random.default_rng(3)
    q = rng.normal(0, 9.0, (3, 257))
    M, rs = 3e10, 7.0
    kep = gb.KeplerPotential(m=M)
    mp0 = gb.MultipolePotential(lmax=0, inner=False, m=M, r_s=rs, S00=-np.sqrt(4 * np.pi))
    mp10 = gb.MultipolePotential(lmax=1, inner=True, m=M, r_s=rs, S10=1.0)
    mp22 = gb.MultipolePotential(lmax=2, inner=True, m=M, r_s=rs, S22=1.0)
    for _, chk in checkers(ref, port):
        assert np.allclose(chk.energy(mp0, q), chk.energy(kep, q), rtol=1e-13)
        assert np.allclose(chk.gradient(mp0, q), chk.gradient(kep, q), rtol=1e-12, atol=1e-18)
        g = chk.gradient(mp10, q)
        assert np.allclose(g[2], G * M / rs ** 2 * np.sqrt(3 / (4 * np.pi)), rtol=1e-12)
        assert np.abs(g[:2]).max() < 1e-12 * np.abs(g[2]).max()
        c22 = G * M / rs ** 3 * np.sqrt(5 / (4 * np.pi) / 24.0) * 3.0
        assert np.allclose(chk.energy(mp22, q), c22 * (q[0] ** 2 - q[1] ** 2), rtol=1e-11, atol=1e-12 * c22 * 81)
        g = chk.gradient(mp22, q)
        scale = np.abs(g).max()
        assert np.allclose(g[0], 2 * c22 * q[0], atol=1e-11 * scale)
        assert np.allclose(g[1], -2 * c22 * q[1], atol=1e-11 * scale)
        assert np.abs(g[2]).max() < 1e-11 * scale
        assert np.all(chk.density(mp22, q) == 0.0)      # mp_density returns 0 (multipole.cpp:404-420)


@pytest.mark.parametrize("inner", [True, False])
def test_multipole_port_equals_reference_and_is_a_gradient(ref, port, inner):
    pot = _mp_random(5, inner, 40 + inner)
    rng = np.random.default_rng(8)
    q = rng.normal(0, 9.0, (3, 513))
    g_ref, g_port = ref.gradient(pot, q), port.gradient(pot, q)
    scale = np.sqrt((g_ref ** 2).sum(0))
    assert np.max(np.sqrt(((g_ref - g_port) ** 2).sum(0)) / scale) < 1e-12
    assert np.allclose(ref.energy(pot, q), port.energy(pot, q), rtol=1e-12)
    # the gradient is the derivative of the value (the reference's own generic potential test,
    # tests/potential/potential/potential_helpers.py numerical-gradient check)
    h = 1e-5
    for k in range(3):
        dq = np.zeros((3, 1)); dq[k] = h
        num = (port.energy(pot, q + dq) - port.energy(pot, q - dq)) / (2 * h)
        assert np.max(np.abs(num - g_port[k]) / scale) < 1e-7
    # on the z axis the theta and phi components are dropped (multipole.cpp:117-121, 281-283)
    qz = np.array([[0.0], [0.0], [5.0]])
    assert np.allclose(ref.gradient(pot, qz), port.gradient(pot, qz), rtol=1e-12, atol=1e-30)


# ---- the GSL stand-ins of oracle/gsl_shim against scipy.special ----------------------
def _shim():
    import ctypes as C
    from oracle import oracle
    path = os.path.join(oracle._REF_DIR, "libgsl_shim_probe.so")
    if not os.path.exists(path):
        oracle.build()
    return C.CDLL(path)


def _shim_call(fn, ints, dbls):
    import ctypes as C
    n = len(dbls[-1])
    out = np.empty(n)
    args = [np.ascontiguousarray(a, dtype=np.int32) for a in ints] + [np.ascontiguousarray(a, dtype=np.float64) for a in dbls]
    fn(*[a.ctypes.data_as(C.c_void_p) for a in args], C.c_int(n), out.ctypes.data_as(C.c_void_p))
    return out


def test_gsl_shim_gegenbauer_matches_scipy():
    """gsl_sf_gegenpoly_n(n, 2l + 3/2, xi) as bfe_helper.cpp:17,25,56-57 calls it: n <= 12, l <= 8, xi in [-1, 1]."""
    from scipy.special import eval_gegenbauer
    L = _shim()
    xi = np.concatenate([np.linspace(-1, 1, 41), np.random.default_rng(0).uniform(-1, 1, 60)])
    n, l, x = np.meshgrid(np.arange(13), np.arange(9), xi, indexing="ij")
    n, l, x = n.ravel(), l.ravel(), x.ravel()
    lam = 2.0 * l + 1.5
    got = _shim_call(L.shim_gegenpoly_n, [n], [lam, x])
    want = eval_gegenbauer(n, lam, x)
    scale = np.maximum(np.abs(want), eval_gegenbauer(n, lam, 1.0) * 1e-3)     # relative to the polynomial's own size
    assert np.max(np.abs(got - want) / scale) < 1e-13
    # arbitrary-precision arbiter on a sub-sample: the three-term recurrence in exact rational arithmetic
    from fractions import Fraction
    for nn, ll, xx in [(12, 8, 0.3), (7, 3, -0.9), (10, 0, 0.999), (4, 6, -1.0)]:
        lamf, xf = Fraction(2 * ll) + Fraction(3, 2), Fraction(xx)
        c0, c1 = Fraction(1), 2 * lamf * xf
        for k in range(2, nn + 1):
            c0, c1 = c1, (2 * (k + lamf - 1) * xf * c1 - (k + 2 * lamf - 2) * c0) / k
        g = _shim_call(L.shim_gegenpoly_n, [[nn]], [[float(lamf)], [xx]])[0]
        assert abs(g - float(c1)) <= 1e-13 * abs(float(c1))


def test_gsl_shim_legendre_matches_scipy():
    """gsl_sf_legendre_Plm / sphPlm (bfe_helper.cpp:21,28,42,46,66; multipole.cpp:40-111) with the Condon-Shortley
    phase, l <= 8 (and up to the multipole kernel's lmax = 15), x = cos(theta) over [-1, 1]."""
    from scipy.special import lpmv, gammaln
    L = _shim()
    xs = np.concatenate([np.linspace(-1, 1, 41), np.random.default_rng(1).uniform(-1, 1, 60)])
    ls, ms, x = [], [], []
    for l in range(16):
        for m in range(l + 1):
            ls += [l] * xs.size; ms += [m] * xs.size; x += list(xs)
    ls, ms, x = np.array(ls), np.array(ms), np.array(x)
    want = lpmv(ms, ls, x)
    got = _shim_call(L.shim_legendre_Plm, [ls, ms], [x])
    # size of P_l^m over [-1,1] grows like (l+m)!/(l-m)!: compare relative to the per-(l,m) maximum
    key = ls * 100 + ms
    scale = np.zeros_like(want)
    for k in np.unique(key):
        sel = key == k
        scale[sel] = np.abs(want[sel]).max()
    lo = ls <= 8
    assert np.max(np.abs(got - want)[lo] / scale[lo]) < 1e-13
    assert np.max(np.abs(got - want) / scale) < 5e-13
    norm = np.sqrt((2 * ls + 1) / (4 * np.pi) * np.exp(gammaln(ls - ms + 1) - gammaln(ls + ms + 1)))
    got_s = _shim_call(L.shim_legendre_sphPlm, [ls, ms], [x])
    assert np.max(np.abs(got_s - norm * want)[lo]) < 1e-13       # |Y_lm| <= sqrt((2l+1)/4pi) ~ O(1)
    assert np.max(np.abs(got_s - norm * want)) < 5e-13
    try:                                                             # scipy >= 1.15: the spherical harmonic itself
        from scipy.special import sph_harm_y
        th = np.arccos(np.clip(x, -1, 1))
        y = sph_harm_y(ls, ms, th, np.zeros_like(th)).real
        assert np.max(np.abs(got_s - y)) < 5e-13
    except ImportError:
        pass
    # Gamma at the integer arguments of bfe_helper.cpp:72-74 / multipole.cpp:106-111
    from scipy.special import gamma
    k = np.arange(1.0, 35.0)
    assert np.max(np.abs(_shim_call(L.shim_gamma, [], [k]) / gamma(k) - 1)) < 1e-14


def test_multipole_closed_form_lmax5(ref):
    """MultipolePotential through the reference's own multipole.cpp (compiled against the shim) equals the closed
    form written with scipy: Phi = (G M / r_s) sum_lm s^-(l+1) [s^l inner] Y_lm(cos theta) (S_lm cos m phi +
    T_lm sin m phi), s = r / r_s (multipole.cpp:50-66,182-224); gradient against 4th-order differences of that form."""
    from scipy.special import lpmv, gammaln
    rng = np.random.default_rng(55)
    for inner in (False, True):
        lmax = 5
        kw = {}
        for l in range(lmax + 1):
            for m in range(l + 1):
                kw[f"S{l}{m}"] = rng.normal()
                if m:
                    kw[f"T{l}{m}"] = rng.normal()
        pot = gb.MultipolePotential(lmax=lmax, m=3e10, r_s=7.0, inner=inner, **kw)

        def phi(q):
            r = np.sqrt((q * q).sum(0)); s = r / 7.0
            X = q[2] / r; az = np.arctan2(q[1], q[0])
            out = np.zeros_like(r)
            for l in range(lmax + 1):
                for m in range(l + 1):
                    nlm = np.sqrt((2 * l + 1) / (4 * np.pi) * np.exp(gammaln(l - m + 1) - gammaln(l + m + 1)))
                    rad = s ** l if inner else s ** (-(l + 1.0))
                    out += rad * nlm * lpmv(m, l, X) * (kw[f"S{l}{m}"] * np.cos(m * az) + kw.get(f"T{l}{m}", 0.0) * np.sin(m * az))
            return pot.G * 3e10 / 7.0 * out

        q = rng.normal(0, 6.0, (3, 400))
        q = q[:, np.sqrt((q * q).sum(0)) > 1.0]
        e = ref.energy(pot, q)
        assert np.max(np.abs(e - phi(q)) / np.abs(phi(q)).max()) < 1e-13
        g = ref.gradient(pot, q)
        h = 1e-3
        gfd = np.empty_like(g)
        for k in range(3):
            d = np.zeros((3, 1)); d[k] = h
            gfd[k] = (-phi(q + 2 * d) + 8 * phi(q + d) - 8 * phi(q - d) + phi(q - 2 * d)) / (12 * h)
        assert np.max(np.abs(g - gfd)) / np.abs(g).max() < 1e-8


# ---- GSL spline stand-in (oracle/gsl_shim/gsl/gsl_spline.h) and the reference's TimeInterpolatedPotential ----------
def _shim_spline(kind, xk, yk, t):
    import ctypes as C
    L = _shim()
    xk, yk, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (xk, yk, t))
    out = np.empty_like(t)
    rc = L.shim_spline_eval(kind, xk.ctypes.data_as(C.c_void_p), yk.ctypes.data_as(C.c_void_p), xk.size,
                            t.ctypes.data_as(C.c_void_p), t.size, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def test_gsl_shim_splines_match_scipy():
    """linear / cspline (natural) / akima against scipy.interpolate on irregular knots; steffen against an independent
    numpy statement of Steffen (1990) eqs. 2-11 plus its defining property (monotone between knots)."""
    from scipy.interpolate import Akima1DInterpolator, CubicSpline
    rng = np.random.default_rng(7)
    for n in (5, 6, 12, 40):
        xk = np.sort(rng.uniform(-50, 150, n)); yk = rng.normal(0, 3, n)
        t = np.concatenate([np.linspace(xk[0], xk[-1], 801), xk])
        scale = np.abs(yk).max()
        assert np.max(np.abs(_shim_spline(0, xk, yk, t) - np.interp(t, xk, yk))) <= 1e-14 * scale
        assert np.max(np.abs(_shim_spline(1, xk, yk, t) - CubicSpline(xk, yk, bc_type="natural")(t))) <= 1e-12 * scale
        assert np.max(np.abs(_shim_spline(2, xk, yk, t) - Akima1DInterpolator(xk, yk)(t))) <= 1e-12 * scale
        # Steffen: slopes at the knots, then the cubic Hermite form
        h = np.diff(xk); s = np.diff(yk) / h
        yp = np.empty(n); yp[0] = s[0]; yp[-1] = s[-1]
        p = (s[:-1] * h[1:] + s[1:] * h[:-1]) / (h[:-1] + h[1:])
        yp[1:-1] = (np.sign(s[:-1]) + np.sign(s[1:])) * np.minimum(np.minimum(np.abs(s[:-1]), np.abs(s[1:])), 0.5 * np.abs(p))
        i = np.clip(np.searchsorted(xk, t, side="right") - 1, 0, n - 2)
        u = (t - xk[i]) / h[i]
        herm = ((2 * u ** 3 - 3 * u ** 2 + 1) * yk[i] + (u ** 3 - 2 * u ** 2 + u) * h[i] * yp[i]
                + (-2 * u ** 3 + 3 * u ** 2) * yk[i + 1] + (u ** 3 - u ** 2) * h[i] * yp[i + 1])
        got = _shim_spline(3, xk, yk, t)
        assert np.max(np.abs(got - herm)) <= 1e-12 * scale
        for k in range(n - 1):           # monotone on every interval
            seg = got[:801][(t[:801] >= xk[k]) & (t[:801] <= xk[k + 1])]
            d = np.diff(seg)
            assert np.all(d * np.sign(yk[k + 1] - yk[k]) >= -1e-12 * scale)
    assert np.isnan(_shim_spline(1, [0, 1, 2.0], [1, 2, 3.0], [-0.1, 2.1])).all()       # GSL_EDOM outside the knots


def test_reference_time_interpolated_doctest_numbers(ref):
    """The reference's own TimeInterpolatedPotential (time_interp.cpp / time_interp_wrapper.cpp compiled against the
    spline stand-in, wired like cytimeinterp.pyx:157-300) reproduces the numbers of its class docstring
    (time_interpolated.py:77-110): Kepler with a linearly growing mass, and a LongMuraliBar rotating by 90 deg in 1 Gyr."""
    t = np.linspace(0, 100, 11)
    pot = gb.TimeInterpolatedPotential(gb.KeplerPotential, t, m=np.linspace(1e10, 2e10, 11))
    q = np.array([[1e-3], [0.0], [0.0]])
    assert abs(ref.energy(pot, q, t=0.0)[0] - (-44.98502151)) < 5e-9
    assert abs(ref.energy(pot, q, t=50.0)[0] - (-67.47753227)) < 5e-9
    assert np.isnan(ref.energy(pot, q, t=100.5)[0]) and np.isnan(ref.gradient(pot, q, t=-1.0)).all()
    from scipy.spatial.transform import Rotation
    Rs = np.array([Rotation.from_rotvec([0, 0, a]).as_matrix() for a in np.linspace(0, np.pi / 2, 11)])
    bar = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, np.linspace(0, 1000, 11), m=1e10, a=3.0, b=1.0, c=0.5, R=Rs)
    q = np.array([[5.0], [0.0], [0.0]])
    assert abs(ref.gradient(bar, q, t=0.0)[0, 0] - 0.00207787) < 5e-9
    assert abs(ref.gradient(bar, q, t=500.0)[0, 0] - 0.0015879) < 5e-8
    with pytest.raises(ValueError):
        gb.TimeInterpolatedPotential(gb.KeplerPotential, t[:4], interpolation_method="akima", m=np.ones(4))
    with pytest.raises(NotImplementedError):
        gb.TimeInterpolatedPotential(gb.NullPotential, t)
    with pytest.raises(ValueError):
        pot.integrate_orbit(np.ones(6), dt=1.0, n_steps=200)


def test_reference_time_interpolated_quirks_are_what_the_gpu_tests_assume(ref):
    """Three properties of the reference's own C/C++ that the GPU tests lean on, pinned here on the CPU so that a change of
    the reference (or of how the oracle drives it) is noticed where it is cheap:
    (1) ``time_interp_gradient`` back-rotates a batch with AoS indices on SoA arrays (time_interp_wrapper.cpp:189-201):
        a multi-point call with a rotating component differs from per-point calls -- the GPU tests call it per point;
    (2) ``time_interp_hessian`` zeroes its output before accumulating (:303-312): the Hessian of a composite is the last
        TimeInterpolated component's plus what follows it;
    (3) ``dop853_step`` runs the stiffness test after every accepted step (nstiff = 1, dop853.pyx:64): a stream particle
        in a potential with a turning bar can stop with code -4, while the same call in the static part does not."""
    from scipy.spatial.transform import Rotation
    Tk = np.linspace(-200.0, 10.0, 43)
    Rk = np.array([Rotation.from_rotvec([0.0, 0.0, -0.04 * a]).as_matrix() for a in Tk])
    halo = gb.NFWPotential(m=6e11, r_s=16.0)
    disk = gb.MiyamotoNagaiPotential(m=6e10, a=3.0, b=0.3)
    bar = gb.TimeInterpolatedPotential(gb.LongMuraliBarPotential, Tk, m=1e10, a=4.0, b=0.8, c=0.25, R=Rk)
    gal = gb.CCompositePotential(halo=halo, disk=disk, bar=bar)
    q = np.ascontiguousarray(np.random.default_rng(2).normal(0, 6.0, (3, 6)))
    # (1)
    batch = ref.gradient(bar, q, -50.0)
    single = np.concatenate([ref.gradient(bar, np.ascontiguousarray(q[:, i:i + 1]), -50.0) for i in range(6)], axis=1)
    assert np.abs(batch - single).max() > 1e-3 * np.abs(single).max()           # not even the first point survives the batch
    h = 1e-5                                       # the per-point gradient IS the derivative of the reference's potential
    for j in range(3):
        dq = np.zeros((3, 1)); dq[j] = h
        num = (ref.energy(bar, q[:, :1] + dq, -50.0) - ref.energy(bar, q[:, :1] - dq, -50.0)) / (2 * h)
        assert np.isclose(num[0], single[j, 0], rtol=1e-6)
    # (2)
    Hc = ref.hessian(gal, q, -50.0)
    assert np.allclose(Hc, ref.hessian(bar, q, -50.0), rtol=1e-13, atol=0)
    assert np.abs(Hc - sum(ref.hessian(p, q, -50.0) for p in (halo, disk, bar))).max() > 0.1 * np.abs(Hc).max()
    # (3)
    KMS = gb.KMS_TO_KPC_MYR
    H_bar, H_static = gb.Hamiltonian(gal), gb.Hamiltonian(gb.CCompositePotential(halo=halo, disk=disk))
    w0 = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * KMS, 50.0 * KMS]).reshape(6, 1)
    t = -np.arange(121.0)
    back, _, rc = ref.dop853(H_bar, w0, t, nbatch=1)
    assert rc == 0
    px, pv, pt = back[:3, ::-1, 0].T.copy(), back[3:, ::-1, 0].T.copy(), t[::-1].copy()
    x0, v0, t10 = oracle.fardal_release_numpy(ref, gal, px, pv, pt, np.full(121, 2.5e4), np.full(121, 3, dtype="i4"),
                                              np.random.RandomState(7), gala_modified=True)
    rows = np.hstack([x0, v0])
    codes_bar, codes_static = [], []
    for t1 in np.unique(t10):
        grp = np.ascontiguousarray(rows[t10 == t1])
        codes_bar.append(ref.dop853_step_rows(H_bar, grp, t1, 0.0, 1.0, group=True)[2])
        codes_static.append(ref.dop853_step_rows(H_static, grp, t1, 0.0, 1.0, group=True)[2])
    assert min(codes_static) >= 0 and min(codes_bar) == -4 and set(codes_bar) <= {0, 1, -4}
    print(f"\n[reference dop853_step in a turning bar] {codes_bar.count(-4)} of {len(codes_bar)} release times stop with code -4")
