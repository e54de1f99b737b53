"""`-m gpu`: on-device trajectory reductions (gb_orbit_extrema / gb_integrate_extrema) against a numpy
restatement of the reference's Orbit._max_helper / pericenter / apocenter (dynamics/orbit.py:391-553) applied to the
same trajectories: scipy.signal.argrelmax(mode="wrap") with the two edge samples removed, then
np.polynomial.polynomial.polyfit(t[j-1:j+2], r[j-1:j+2], 2) and the vertex of that parabola."""
import numpy as np
import pytest
from scipy.signal import argrelmax

import gala_b200 as gb
from conftest import make_ic

pytestmark = pytest.mark.gpu


def reference_extrema(t, r, sign):
    """orbit.py:391-422 for one orbit; sign = +1 apocentres, -1 pericentres.  Returns (values, times)."""
    arr = sign * r
    ix = argrelmax(arr, mode="wrap")[0]
    ix = ix[(ix != 0) & (ix != (len(arr) - 1))]
    vals, times = np.zeros(len(ix)), np.zeros(len(ix))
    for i, j in enumerate(ix):
        co = np.polynomial.polynomial.polyfit(t[j - 1:j + 2], arr[j - 1:j + 2], 2)
        times[i] = -co[1] / (2 * co[2])
        vals[i] = co[2] * times[i] ** 2 + co[1] * times[i] + co[0]
    return sign * vals, times


def check(stats, t, w, N):
    r = np.sqrt((w[:3] ** 2).sum(0))                    # (ntimes, N)
    if t[-1] < t[0]:
        t, r = t[::-1], r[::-1]
    for kind, sign in (("peri", -1.0), ("apo", 1.0)):
        for i in range(N):
            v, tt = reference_extrema(t, r[:, i], sign)
            assert stats[f"n_{kind}"][i] == len(v), (kind, i)
            if len(v) == 0:
                assert np.isnan(stats[f"{kind}_mean"][i])
                continue
            # the reference's polyfit works in absolute t (conditioning ~ t^2 / dt^2); centred coordinates here
            assert abs(stats[f"{kind}_mean"][i] - v.mean()) <= 1e-8 * v.mean(), (kind, i)
            assert abs(stats[f"{kind}_min"][i] - v.min()) <= 1e-8 * v.min()
            assert abs(stats[f"{kind}_max"][i] - v.max()) <= 1e-8 * v.max()
            assert abs(stats[f"{kind}_t_first"][i] - tt[0]) <= 1e-5 * abs(t[1] - t[0])
            assert abs(stats[f"{kind}_t_last"][i] - tt[-1]) <= 1e-5 * abs(t[1] - t[0])
    assert np.allclose(stats["abs_z_max"], np.abs(w[2]).max(0), rtol=0, atol=0)
    az = np.abs(w[2])
    if t[-1] < t[0]:
        az = az[::-1]
    for i in range(N):                                   # Orbit.zmax: local maxima of |z| (orbit.py:600-656)
        v, tt = reference_extrema(t, az[:, i], 1.0)
        assert stats["n_zmax"][i] == len(v)
        if len(v):
            assert abs(stats["zmax_mean"][i] - v.mean()) <= 1e-8 * max(v.mean(), 1e-3) and abs(stats["zmax_max"][i] - v.max()) <= 1e-8 * max(v.max(), 1e-3)


@pytest.mark.parametrize("integ", ["leapfrog", "ruth4"])
@pytest.mark.parametrize("strict", [True, False])
def test_integrate_extrema_matches_reference_helper(integ, strict):
    H = gb.Hamiltonian(gb.MilkyWayPotential2022())
    H.strict_math = strict
    N = 300
    w0 = make_ic(lambda q: H.potential.gradient(q), N, 31)
    t = np.arange(3001) * 0.5
    fn = gb.leapfrog_integrate_hamiltonian if integ == "leapfrog" else gb.ruth4_integrate_hamiltonian
    _, w = fn(H, w0, t, save_all=1)
    st = gb.integrate_extrema(H, w0, t, Integrator=integ, with_energy=True, return_final=True)
    if strict:
        assert np.array_equal(st["w_final"], w[:, -1])            # same statements, same order as the plain kernel
    else:
        assert np.max(np.abs(st["w_final"] - w[:, -1])) <= 1e-9 * np.abs(w[:, -1]).max()
    if strict:
        check(st, t, w, N)
    E = H.energy(w.reshape(6, -1)).reshape(t.size, N)
    assert np.allclose(st["E_first"], E[0], rtol=1e-13)
    assert np.allclose(st["E_last"], E[-1], rtol=1e-9)
    assert np.allclose(st["dE_max"], np.abs(E - E[0]).max(0), rtol=1e-6, atol=1e-16)
    # the plummer doctest numbers of the reference (docs/dynamics/orbits-in-detail.rst:321-330) through the reduction
    assert (st["n_peri"] >= 1).mean() > 0.9


@pytest.mark.parametrize("backward", [False, True])
def test_orbit_extrema_of_existing_trajectory(backward):
    """gb_orbit_extrema on a dense DOP853 output (host arrays and device tensors), and Orbit.pericenter/apocenter."""
    import torch
    pot = gb.PlummerPotential(m=1e10, b=1.0)
    H = gb.Hamiltonian(pot)
    N = 64
    w0 = make_ic(lambda q: pot.gradient(q), N, 32, rmin=1.0, rmax=10.0)
    t = np.linspace(0, -2000 if backward else 2000, 2001)
    _, w = gb.dop853_integrate_hamiltonian(H, w0, t)
    st = gb.orbit_extrema(H, w, t, with_energy=True)
    check(st, t, w, N)
    std = gb.orbit_extrema(H, torch.as_tensor(w, device="cuda"), t, with_energy=True)
    for k in st:
        assert np.array_equal(st[k], std[k].cpu().numpy(), equal_nan=True), k
    orbit = H.integrate_orbit(w0, Integrator="dopri853", t=t)
    peri, apo = orbit.pericenter(), orbit.apocenter(func=np.max)
    assert np.array_equal(peri, st["peri_mean"], equal_nan=True) and np.array_equal(apo, st["apo_max"], equal_nan=True)
    one = H.integrate_orbit(w0[:, 0], Integrator="dopri853", t=t)
    assert one.pericenter() == st["peri_mean"][0]
    with pytest.raises(ValueError):
        orbit.pericenter(return_times=True)
    # func=None: every extremum and its time (orbit.py:473-480), one array per orbit; other reductions apply to the list
    r = np.sqrt((w[:3] ** 2).sum(0))
    tt = t[::-1] if backward else t
    rr = r[::-1] if backward else r
    vals, times = orbit.apocenter(func=None, return_times=True)
    for i in (0, 7, N - 1):
        v, tv = reference_extrema(tt, rr[:, i], 1.0)
        assert len(vals[i]) == len(v) and np.allclose(vals[i], v, rtol=1e-8) and np.allclose(times[i], tv, rtol=0, atol=1e-5 * abs(t[1] - t[0]))
    v1, t1 = one.pericenter(func=None, return_times=True)
    assert np.array_equal(v1, orbit.pericenter(func=None)[0]) and len(v1) == st["n_peri"][0]
    assert np.allclose(orbit.pericenter(func=np.median), [np.median(v) if len(v) else np.nan for v in orbit.pericenter(func=None)], equal_nan=True)
    ecc = orbit.eccentricity()
    assert np.allclose(ecc, (st["apo_mean"] - st["peri_mean"]) / (st["apo_mean"] + st["peri_mean"]), equal_nan=True)
    assert np.all((ecc[np.isfinite(ecc)] >= 0) & (ecc[np.isfinite(ecc)] < 1))
    zm = orbit.zmax()
    assert zm.shape == (N,) and np.array_equal(zm, st["zmax_mean"], equal_nan=True)


def test_integrate_extrema_errors_and_devices():
    H = gb.Hamiltonian(gb.MilkyWayPotential2022(), gb.ConstantRotatingFrame([0, 0, 0.03]))
    w0 = np.ones((6, 4)); t = np.arange(5.0)
    with pytest.raises(TypeError):
        gb.integrate_extrema(H, w0, t, Integrator="leapfrog")
    with pytest.raises(ValueError):
        gb.integrate_extrema(H, w0, t, Integrator="dopri853")
    st = gb.integrate_extrema(H, w0 * np.array([[8.0], [0.0], [1.0], [0.0], [0.2], [0.0]]), np.arange(400) * 1.0, Integrator="ruth4")
    assert st["n_apo"].shape == (4,)


def test_reference_orbit_tests_kepler():
    """The reference's own tests of these methods (tests/dynamics/test_orbit.py:352-446): Kepler m = 1 in solar-system
    units -- a circular orbit has |e| < 1e-3; for v = 1.5 pi au/yr the mean apocentre / pericentre sit at the roots of
    2 (E - Phi(r)) - L^2 / r^2 to 1 %, and the individual values scatter by < 1e-4."""
    import scipy.optimize as so
    pot = gb.KeplerPotential(m=1.0, units=gb.solarsystem)
    H = gb.Hamiltonian(pot)
    w = H.integrate_orbit(np.array([1.0, 0.0, 0.0, 0.0, 2 * np.pi, 0.0]), dt=0.01, n_steps=10000, Integrator="dopri853")
    assert abs(w.eccentricity()) < 1e-3
    w = H.integrate_orbit(np.array([1.0, 0.0, 0.0, 0.0, 1.5 * np.pi, 0.0]), dt=0.01, n_steps=10000, Integrator="dopri853")
    apo, per, zmax = w.apocenter(), w.pericenter(), w.zmax()
    assert np.shape(apo) == () and np.shape(per) == () and np.shape(zmax) == () and apo > per
    E = np.mean(w.energy())
    L = np.mean(np.sqrt((np.cross(w.pos.T, w.vel.T) ** 2).sum(1)))
    f = lambda r: 2 * (E - pot.energy(np.array([[r], [0.0], [0.0]]))[0]) - L ** 2 / r ** 2
    assert np.isclose(apo, so.brentq(f, 0.9, 1.0), rtol=1e-2) and np.isclose(per, so.brentq(f, 0.3, 0.5), rtol=1e-2)
    apos, pers = w.apocenter(func=None), w.pericenter(func=None)
    for v in (apos, pers):
        d = np.std(v) / np.mean(v)
        assert 0 < d < 1e-4
    # several orbits at once
    w0 = np.array([[1.0, 0, 0, 0, 1.5 * np.pi, 0], [1.1, 0, 0, 0, 1.5 * np.pi, 0]]).T
    w2 = H.integrate_orbit(w0, dt=0.01, n_steps=10000)
    per2, apo2, ecc2 = w2.pericenter(), w2.apocenter(), w2.eccentricity()
    assert per2.shape == (2,) and np.all(apo2 > per2) and np.all((ecc2 > 0) & (ecc2 < 1))
