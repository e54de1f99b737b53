"""``Hamiltonian`` = potential + frame (reference ``potential/hamiltonian/chamiltonian.pyx``), with
``integrate_orbit`` dispatching exactly like ``chamiltonian.pyx:225-380``."""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _abi
from .dynamics import Orbit, PhaseSpacePosition
from .frame import ConstantRotatingFrame, StaticFrame
from .integrate import (DOPRI853Integrator, LeapfrogIntegrator, Ruth4Integrator, dop853_integrate_hamiltonian,
                        get_integrator, leapfrog_integrate_hamiltonian, parse_time_specification,
                        ruth4_integrate_hamiltonian)
from .potential import PotentialBase, _alloc_like, _stream_of
from .units import strip

__all__ = ["Hamiltonian"]


class Hamiltonian:
    def __init__(self, potential, frame=None):
        if isinstance(potential, Hamiltonian):
            frame = potential.frame if frame is None else frame
            potential = potential.potential
        if not isinstance(potential, PotentialBase):
            raise TypeError("potential must be a gala_b200 potential")
        self.potential = potential
        self.frame = StaticFrame(potential.units) if frame is None else frame
        self.units = potential.units
        self.c_enabled = bool(potential.c_enabled and self.frame.c_enabled)
        self.strict_math = False

    # -- value / gradient (chamiltonian.pyx:80-128; src/chamiltonian.cpp) ---------------------------
    def _call6(self, fn_name, w, t, rows):
        if _abi._is_torch_cuda(w):
            buf = _abi.Buf(w.reshape(6, -1).contiguous())
        else:
            buf = _abi.Buf(np.ascontiguousarray(np.asarray(strip(w), dtype=np.float64).reshape(6, -1)))
        N = buf.arr.shape[1]
        out = _alloc_like(buf, (rows, N) if rows > 1 else (N,))
        stream, dev = _stream_of(buf)
        opt = _abi.launch_opts(buf.device, self.strict_math or self.potential.strict_math, stream, device=dev,
                               devices=None)
        fr = self.frame.spec()
        fn = getattr(_abi.lib(), fn_name)
        _abi.check(fn(self.potential.spec().ptr(), C.byref(fr), buf.ptr, float(strip(t)), N, _abi.Buf(out).ptr,
                      C.byref(opt)))
        return out

    def energy(self, w, t=0.0):
        """Value of the Hamiltonian at w (6,...) -> (...)."""
        if isinstance(w, PhaseSpacePosition):
            w = w.w()
        shape = tuple(w.shape)
        return self._call6("gb_hamiltonian_energy", w, t, 1).reshape(shape[1:])

    def gradient(self, w, t=0.0):
        if isinstance(w, PhaseSpacePosition):
            w = w.w()
        shape = tuple(w.shape)
        return self._call6("gb_hamiltonian_gradient", w, t, 6).reshape(shape)

    # -- orbit integration ------------------------------------------------------------------------
    def integrate_orbit(self, w0, Integrator=None, Integrator_kwargs=None, cython_if_possible=True,
                        save_all=True, **time_spec):
        """Same dispatch as ``chamiltonian.pyx:225-380``.  Returns an ``Orbit`` (save_all) or a
        ``PhaseSpacePosition``.  ``w0`` is a ``PhaseSpacePosition``, a (6,) / (6,N) array, or a
        (6,N) float64 torch.cuda tensor (result then stays on the device)."""
        if Integrator_kwargs is None:
            Integrator_kwargs = {}
        if Integrator is None:          # chamiltonian.pyx:274-277
            Integrator = LeapfrogIntegrator if isinstance(self.frame, StaticFrame) else DOPRI853Integrator
        Integrator = get_integrator(Integrator)
        if Integrator in (LeapfrogIntegrator, Ruth4Integrator) and not isinstance(self.frame, StaticFrame):
            # chamiltonian.pyx:282-288
            warnings.warn("Using a symplectic integrator with a non-static frame can lead to wildly incorrect "
                          "orbits. It is recommended that you use DOPRI853Integrator instead.", RuntimeWarning)
        if isinstance(w0, PhaseSpacePosition):
            w0 = w0.w()
        if _abi._is_torch_cuda(w0):
            arr = w0 if w0.ndim == 2 else w0.reshape(6, 1)
            single = w0.ndim == 1
        else:
            arr = np.asarray(strip(w0), dtype=np.float64)
            single = arr.ndim == 1
            arr = np.ascontiguousarray(arr.reshape(6, -1))
        if arr.shape[0] != 6:
            raise ValueError("Initial conditions must have shape (6,) or (6, N)")
        t = parse_time_specification(self.units, **time_spec)
        if not self.c_enabled:
            raise TypeError("Input Hamiltonian object does not support C-level access.")
        if Integrator is LeapfrogIntegrator:
            if not isinstance(self.frame, StaticFrame):
                # cython path raises (leapfrog.pyx:64-68); the Python leapfrog of the reference
                # is not part of the accelerated path
                raise TypeError("Leapfrog integration is currently only supported for StaticFrame, "
                                f"not {self.frame.__class__.__name__}")
            tt, w = leapfrog_integrate_hamiltonian(self, arr, t, save_all=int(save_all))
        elif Integrator is Ruth4Integrator:
            rot = not isinstance(self.frame, StaticFrame)
            if rot and cython_if_possible:
                raise TypeError("Leapfrog integration is currently only supported for StaticFrame, not "
                                f"{self.frame.__class__.__name__}.")
            tt, w = ruth4_integrate_hamiltonian(self, arr, t, save_all=int(save_all), allow_rotating_frame=rot)
        elif Integrator is DOPRI853Integrator:
            kw = {k: Integrator_kwargs[k] for k in ("atol", "rtol", "nmax", "err_if_fail", "log_output", "nbatch")
                  if k in Integrator_kwargs}      # dt_max / nstiff are not forwarded (chamiltonian.pyx:338-349)
            tt, w = dop853_integrate_hamiltonian(self, arr, t, save_all=int(save_all), **kw)
        else:
            raise ValueError(f"Integrator {Integrator} is not supported")
        if save_all:
            if single:
                w = w[:, :, 0]
            return Orbit.from_w(w, units=self.units, t=tt, hamiltonian=self)
        if single:
            w = w[:, 0]
        return PhaseSpacePosition.from_w(w, units=self.units, frame=self.frame)

    def __call__(self, w, t=0.0):
        return self.energy(w, t)

    def __repr__(self):
        return f"<Hamiltonian {self.potential!r} in {self.frame!r}>"
