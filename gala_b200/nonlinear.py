"""Chaos indicators on the device: ``fast_lyapunov_max`` (reference ``dynamics/nonlinear.py:11-152`` ->
``dynamics/lyapunov/dop853_lyapunov.pyx``).  The reference accepts ONE orbit per call (ValueError otherwise);
here ``w0`` may also be (6, N): every parent orbit is one device lane."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .dynamics import Orbit, PhaseSpacePosition
from .hamiltonian import Hamiltonian
from .units import strip

__all__ = ["fast_lyapunov_max", "surface_of_section"]


def fast_lyapunov_max(w0, hamiltonian, dt, n_steps, d0=1e-5, n_steps_per_pullback=10, noffset_orbits=2, t1=0.0,
                      atol=1e-10, rtol=1e-10, nmax=0, return_orbit=True, err_if_fail=True):
    """Maximum Lyapunov exponent from ``noffset_orbits`` offset orbits per parent orbit.

    Returns ``LEs`` of shape ``(niter - 1, noffset_orbits)`` -- or ``(niter - 1, noffset_orbits, N)`` for N
    parent orbits -- exactly as the reference builds them (running mean of ln(|d1|/d0) over the elapsed time,
    ``dop853_lyapunov.pyx:102-104``), and with ``return_orbit`` the ``Orbit`` of the parent + offset orbits.
    The initial offset directions come from ``np.random.uniform`` (the global numpy RNG, ``:47``), one
    ``(noffset_orbits, 6)`` block per parent orbit in order."""
    H = hamiltonian if isinstance(hamiltonian, Hamiltonian) else Hamiltonian(hamiltonian)
    if not H.c_enabled:
        raise TypeError("Input Hamiltonian must contain a C-implemented potential and frame.")
    if isinstance(w0, PhaseSpacePosition):
        w0 = w0.w()
    w = np.asarray(strip(w0), dtype=np.float64)
    single = w.ndim == 1
    w = np.ascontiguousarray(w.reshape(6, -1).T)                  # (N, 6) rows
    N = w.shape[0]
    nst = int(n_steps) + 1                                        # nonlinear.py:117: n_steps + 1 is passed down
    t_end = float(nst) * float(dt)
    t = np.linspace(float(t1), t_end, nst)                        # dop853_lyapunov.pyx:37-38 (as coded)
    niter = nst // int(n_steps_per_pullback)
    d0_vec = np.random.uniform(size=(N, int(noffset_orbits), 6))
    d0_vec *= float(d0) / np.linalg.norm(d0_vec, axis=2, keepdims=True)          # :60-63
    d0_vec = np.ascontiguousarray(d0_vec)
    LEs_raw = np.zeros((N, niter, int(noffset_orbits)))
    traj = np.empty((N, nst, 1 + int(noffset_orbits), 6)) if return_orbit else None
    status = np.empty(max(N, 1), dtype=np.int32)
    opt = _abi.launch_opts(False, bool(getattr(H, "strict_math", False) or H.potential.strict_math))
    fr = H.frame.spec()
    rc = _abi.lib().gb_lyapunov_max(H.potential.spec().ptr(), C.byref(fr), w.ctypes.data, d0_vec.ctypes.data, N,
                                    t.ctypes.data, nst, float(d0), int(n_steps_per_pullback), int(noffset_orbits),
                                    float(atol), float(rtol), int(nmax), LEs_raw.ctypes.data,
                                    None if traj is None else traj.ctypes.data, status.ctypes.data, C.byref(opt))
    if rc in (-1, -2, -3, -4):
        if err_if_fail:
            _abi.check(rc)
    else:
        _abi.check(rc)
    # LEs = [sum(LEs[:j]) / t[j * pullback] for j in 1..niter-1]   (:102-104)
    csum = np.cumsum(LEs_raw, axis=1)                             # (N, niter, noff)
    js = np.arange(1, niter)
    LEs = csum[:, js - 1, :] / t[js * int(n_steps_per_pullback)][None, :, None]
    LEs = np.ascontiguousarray(LEs.transpose(1, 2, 0))            # (niter-1, noff, N)
    if single:
        LEs = LEs[:, :, 0]
    if not return_orbit:
        return LEs
    ww = traj.transpose(3, 1, 2, 0)                               # (6, nst, norbits, N)
    if single:
        ww = ww[:, :, :, 0]
    else:
        ww = ww.reshape(6, nst, -1)
    return LEs, Orbit.from_w(np.ascontiguousarray(ww), t=t, hamiltonian=H)


def surface_of_section(orbit, constant_idx, constant_val=0.0):
    """``gala.dynamics.nonlinear.surface_of_section`` (``dynamics/nonlinear.py:293-340``): the samples of the orbit
    closest to the plane ``w[constant_idx] = constant_val`` (local minima of the squared distance) that cross it with
    positive conjugate momentum, as an ``Orbit`` of those samples.  One orbit per call like the reference; for an
    orbit array a list with one section per orbit is returned (the sections have different lengths).  Host-side
    analysis of a trajectory (a device-resident one is copied to the host first)."""
    from scipy.signal import argrelmin
    w = orbit.w()
    if type(w).__module__.startswith("torch"):
        w = w.cpu().numpy()
    ndim = w.shape[0] // 2

    def one(wn):
        cross = argrelmin((wn[constant_idx] - constant_val) ** 2)[0]
        cross = cross[wn[constant_idx + ndim][cross] > 0.0]
        return Orbit(pos=wn[:ndim, cross], vel=wn[ndim:, cross])

    if w.ndim == 2:
        return one(w)
    return [one(w[:, :, n]) for n in range(w.shape[2])]
