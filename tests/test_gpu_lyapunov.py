"""fast_lyapunov_max on the device (one lane per parent orbit) against dop853_lyapunov_max restated around the
compiled reference's dop853() + Fwrapper (oracle/ref_driver.cpp:ref_lyapunov; reference
dynamics/lyapunov/dop853_lyapunov.pyx:22-118, dynamics/nonlinear.py:11-152)."""
import numpy as np
import pytest

import gala_b200 as gb
from conftest import make_ic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rotating", [False, True])
def test_lyapunov_batch_matches_reference(ref, rotating):
    pot = gb.LM10Potential()                                        # triaxial halo: some orbits are chaotic
    frame = gb.ConstantRotatingFrame([0.0, 0.0, 0.02]) if rotating else gb.StaticFrame()
    H = gb.Hamiltonian(pot, frame)
    N = 24
    w0 = make_ic(lambda q: pot.gradient(q), N, seed=9, rmin=6.0, rmax=30.0)
    dt, n_steps, d0, pull, noff = 2.0, 400, 1e-5, 10, 2
    np.random.seed(123)
    LEs, orbit = gb.fast_lyapunov_max(w0, H, dt=dt, n_steps=n_steps, d0=d0, n_steps_per_pullback=pull,
                                      noffset_orbits=noff, return_orbit=True)
    nst = n_steps + 1
    niter = nst // pull
    assert LEs.shape == (niter - 1, noff, N)
    assert orbit.pos.shape == (3, nst, (1 + noff) * N)
    # the same offset vectors again
    np.random.seed(123)
    d0_vec = np.random.uniform(size=(N, noff, 6))
    d0_vec *= d0 / np.linalg.norm(d0_vec, axis=2, keepdims=True)
    t = np.linspace(0.0, nst * dt, nst)
    worst_le, worst_w = 0.0, 0.0
    for p in range(N):
        raw, allw, rc = ref.lyapunov(H, w0[:, p], d0_vec[p], t, d0, pull, save=True)
        assert rc >= 0
        want = np.array([raw[:j].sum(0) / t[j * pull] for j in range(1, niter)])
        # the exponents are running means of ln(|d1|/d0) ~ O(1..10): absolute agreement
        worst_le = max(worst_le, np.abs(LEs[:, :, p] - want).max() / max(np.abs(want).max(), 1e-30))
        got_w = np.vstack([orbit.pos, orbit.vel]).reshape(6, nst, 1 + noff, N)[:, :, :, p].transpose(1, 2, 0)
        worst_w = max(worst_w, np.abs(got_w[-1, 0] - allw[-1, 0]).max() / np.abs(allw[-1, 0]).max())
    print(f"\n[lyapunov rotating={rotating}] worst relative difference of the exponents {worst_le:.2e}, "
          f"of the parent orbit's final state {worst_w:.2e}")
    # offsets of 1e-5 amplify rounding by ~1e5 / pullback: the exponents agree to ~1e-6 of their scale
    assert worst_le < 1e-5 and worst_w < 1e-8
    # single-orbit call has the reference's shapes
    np.random.seed(5)
    l1 = gb.fast_lyapunov_max(w0[:, 0], H, dt=dt, n_steps=100, return_orbit=False)
    assert l1.shape == (101 // 10 - 1, 2)


def test_lyapunov_separates_regular_from_chaotic():
    """A circular orbit in a spherical potential is regular (exponent -> 0 like ln(t)/t); the exponent of a
    box-like orbit in the triaxial LM10 halo stays well above it."""
    sph = gb.Hamiltonian(gb.HernquistPotential(m=1e11, c=5.0))
    G = gb.G_GALACTIC
    r = 10.0
    vc = np.sqrt(G * 1e11 * r / (r + 5.0) ** 2)
    np.random.seed(1)
    reg = gb.fast_lyapunov_max(np.array([r, 0, 0, 0, vc, 0.0]), sph, dt=2.0, n_steps=4000, return_orbit=False)
    assert np.all(np.isfinite(reg)) and reg[-1].max() < 2e-3
    assert reg[-1].max() < 0.5 * reg[len(reg) // 4].max()          # still decaying
