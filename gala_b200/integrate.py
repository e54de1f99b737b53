"""Integrator front-ends and the three functions at the reference's Cython boundary.

``leapfrog_integrate_hamiltonian``, ``ruth4_integrate_hamiltonian`` and
``dop853_integrate_hamiltonian`` keep the reference's names, argument meaning, return shapes and
error behaviour (``integrate/cyintegrators/leapfrog.pyx:54-121``, ``ruth4.pyx:37-113``,
``dop853.pyx:196-250``) and forward to the C ABI.  ``w0`` may be a numpy array (host buffers,
copies inside the call) or a float64 ``torch.cuda`` tensor (device buffers, no copies, result
stays on the device).
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _abi
from .frame import StaticFrame
from .units import strip

__all__ = ["pinned_empty", "parse_time_specification", "LeapfrogIntegrator", "Ruth4Integrator", "DOPRI853Integrator",
           "get_integrator", "leapfrog_integrate_hamiltonian", "ruth4_integrate_hamiltonian",
           "dop853_integrate_hamiltonian", "integrate_extrema", "orbit_extrema", "orbit_extrema_list"]


def parse_time_specification(units=None, dt=None, n_steps=None, t1=None, t2=None, t=None):
    """Restates ``integrate/timespec.py:10-147`` (same accepted combinations, same arithmetic:
    ``dt, n_steps`` -> ``t1 + cumsum`` of n_steps+1 entries)."""
    if n_steps is not None:
        n_steps = int(n_steps)
    dt, t1, t2, t = strip(dt), strip(t1), strip(t2), strip(t)
    if t is not None:
        return np.asarray(t).astype(np.float64)
    if dt is None and (t1 is None or t2 is None or n_steps is None):
        raise ValueError("Invalid specification of integration time. See docstring for more information.")
    if dt is not None and n_steps is not None:
        if t1 is None:
            t1 = 0.0
        times = parse_time_specification(units, dt=np.ones(n_steps + 1) * dt, t1=t1)
    elif dt is not None and t1 is not None and t2 is not None:
        if t2 < t1 and dt < 0:
            t_i, times = t1, []
            while t_i > t2 and len(times) < 1e6:
                times.append(t_i)
                t_i += dt
            if times[-1] != t2:
                times.append(t2)
            return np.array(times, dtype=np.float64)
        if t2 > t1 and dt > 0:
            t_i, times = t1, []
            while t_i < t2 and len(times) < 1e6:
                times.append(t_i)
                t_i += dt
            return np.array(times, dtype=np.float64)
        if dt == 0:
            raise ValueError("dt must be non-zero.")
        raise ValueError("If t2 < t1, dt must be negative. If t1 < t2, dt must be positive.")
    elif isinstance(dt, np.ndarray) and t1 is not None:
        times = np.cumsum(np.append([0.0], dt)) + t1
        times = times[:-1]
    elif dt is None and not (t1 is None or t2 is None or n_steps is None):
        times = np.linspace(t1, t2, n_steps, endpoint=True)
    else:
        raise ValueError("Invalid options. See docstring.")
    return times.astype(np.float64)


# ---- buffers ------------------------------------------------------------------------------------
def pinned_empty(shape, dtype=np.float64):
    """Page-locked host array (numpy view of a pinned torch tensor).  Passing pinned ``w0`` / ``out``
    arrays lets the chunked host pipeline of the C ABI overlap its H2D/D2H copies with the kernels;
    pageable arrays work too, the copies are then staged by the driver."""
    import torch
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    return torch.empty(tuple(np.atleast_1d(shape)), dtype=tdt).pin_memory().numpy()


def _check_out(out, like, shape):
    if like.device:
        ok = _abi._is_torch_cuda(out) and tuple(out.shape) == tuple(shape) and out.is_contiguous() \
            and str(out.dtype) == "torch.float64"
    else:
        ok = isinstance(out, np.ndarray) and out.shape == tuple(shape) and out.dtype == np.float64 \
            and out.flags.c_contiguous and out.flags.writeable
    if not ok:
        raise ValueError(f"out must be a C-contiguous float64 array of shape {tuple(shape)} on the same side "
                         "(host/device) as w0")
    return out


def _prep_w0(w0):
    if _abi._is_torch_cuda(w0):
        import torch
        if w0.dtype != torch.float64 or w0.ndim != 2 or w0.shape[0] != 6:
            raise ValueError("w0 must be a float64 tensor of shape (6, N)")
        return _abi.Buf(w0.contiguous())
    a = np.ascontiguousarray(w0, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] != 6:
        raise ValueError(f"w0 must have shape (6, N), got {a.shape}")
    return _abi.Buf(a)


def _prep_t(t, like):
    th = np.ascontiguousarray(strip(t) if not _abi._is_torch_cuda(t) else t.cpu().numpy(), dtype=np.float64)
    if th.ndim != 1:
        raise ValueError("t must be one-dimensional")
    if like.device:
        import torch
        return th, _abi.Buf(torch.as_tensor(th, device=like.arr.device))
    return th, _abi.Buf(th)


def _alloc(like, shape, dtype="f8"):
    if like.device:
        import torch
        return torch.empty(shape, dtype=torch.float64 if dtype == "f8" else torch.int32, device=like.arr.device)
    return np.empty(shape, dtype=np.float64 if dtype == "f8" else np.int32)


def _opts(like, hamiltonian, block=0):
    stream, dev = None, -1
    if like.device:
        import torch
        stream = torch.cuda.current_stream(like.arr.device).cuda_stream
        dev = like.arr.device.index
    strict = bool(getattr(hamiltonian, "strict_math", False) or getattr(hamiltonian.potential, "strict_math", False))
    return _abi.launch_opts(like.device, strict, stream, block=block, device=dev, devices=None)


def _check_c_enabled(hamiltonian):
    if not getattr(hamiltonian, "c_enabled", False):
        raise TypeError("Input Hamiltonian object does not support C-level access.")


# ---- the boundary functions -----------------------------------------------------------------------
def _fixed_step(fn_name, hamiltonian, w0, t, save_all, out=None):
    _check_c_enabled(hamiltonian)
    w = _prep_w0(w0)
    th, tb = _prep_t(t, w)
    N, ntimes = w.arr.shape[1], th.size
    shape = (6, ntimes, N) if save_all else (6, N)
    out = _alloc(w, shape) if out is None else _check_out(out, w, shape)
    opt = _opts(w, hamiltonian)
    fr = hamiltonian.frame.spec()
    fn = getattr(_abi.lib(), fn_name)
    _abi.check(fn(hamiltonian.potential.spec().ptr(), C.byref(fr), w.ptr, N, tb.ptr, ntimes, int(bool(save_all)),
                  _abi.Buf(out).ptr, C.byref(opt)))
    return (th, out) if save_all else (th[-1:], out)


def leapfrog_integrate_hamiltonian(hamiltonian, w0, t, save_all=1, out=None):
    """w0 (6,N) -> (t, w[6, ntimes, N]) or (t[-1:], w[6, N]); StaticFrame only (TypeError otherwise),
    like ``leapfrog.pyx:54-121``.  ``out`` (extension): preallocated result array, e.g. ``pinned_empty``."""
    if not isinstance(hamiltonian.frame, StaticFrame):
        _check_c_enabled(hamiltonian)
        raise TypeError("Leapfrog integration is currently only supported for StaticFrame, "
                        f"not {hamiltonian.frame.__class__.__name__}")
    return _fixed_step("gb_leapfrog", hamiltonian, w0, t, save_all, out)


def ruth4_integrate_hamiltonian(hamiltonian, w0, t, save_all=1, allow_rotating_frame=False, out=None):
    """``ruth4.pyx:37-113``.  The Cython function raises TypeError for a non-static frame; pass
    ``allow_rotating_frame=True`` to run the reference's *Python* Ruth4 semantics in a
    ConstantRotatingFrame on the GPU (what ``cython_if_possible=False`` executes in the reference)."""
    if not isinstance(hamiltonian.frame, StaticFrame) and not allow_rotating_frame:
        _check_c_enabled(hamiltonian)
        raise TypeError("Leapfrog integration is currently only supported for StaticFrame, not "
                        f"{hamiltonian.frame.__class__.__name__}.")
    return _fixed_step("gb_ruth4", hamiltonian, w0, t, save_all, out)


def dop853_integrate_hamiltonian(hamiltonian, w0, t, atol=1e-10, rtol=1e-10, nmax=0, dt_max=0.0, nstiff=0,
                                 save_all=1, err_if_fail=1, log_output=0, nbatch=100, return_status=False, out=None):
    """``dop853.pyx:196-250``.  Step-size control is per orbit (the reference's ``nbatch=1``); the
    ``nbatch`` argument is accepted and ignored.  With ``err_if_fail`` a failed orbit raises
    ``RuntimeError("Integration failed with code ...")`` like ``dop853.pyx:184-185``.  ``out`` (extension):
    preallocated result array, e.g. ``pinned_empty`` for the host-buffer path."""
    _check_c_enabled(hamiltonian)
    w = _prep_w0(w0)
    th, tb = _prep_t(t, w)
    N, ntimes = w.arr.shape[1], th.size
    if ntimes < 1:
        raise ValueError("ntimes must be greater than 1")
    shape = (6, ntimes, N) if save_all else (6, N)
    out = _alloc(w, shape) if out is None else _check_out(out, w, shape)
    status = _alloc(w, (N,), "i4")
    stats = None
    stats_struct = None
    if return_status:
        stats = {k: _alloc(w, (N,), "i4") for k in ("nstep", "naccpt", "nrejct", "nfcn")}
        stats_struct = _abi.gb_dop853_stats(*[C.cast(_abi.Buf(stats[k]).ptr, _abi.c_int32_p)
                                              for k in ("nstep", "naccpt", "nrejct", "nfcn")])
    opt = _opts(w, hamiltonian)
    fr = hamiltonian.frame.spec()
    rc = _abi.lib().gb_dop853(hamiltonian.potential.spec().ptr(), C.byref(fr), w.ptr, N, tb.ptr, ntimes,
                              float(atol), float(rtol), int(nmax), float(dt_max), int(nstiff), int(bool(save_all)),
                              _abi.Buf(out).ptr, _abi.Buf(status).ptr,
                              C.byref(stats_struct) if stats_struct is not None else None, C.byref(opt))
    if rc in (-1, -2, -3, -4):
        if err_if_fail:
            _abi.check(rc)
    else:
        _abi.check(rc)
    res = (th, out) if save_all else (th[-1:], out)
    if return_status:
        stats["status"] = status
        return res + (stats,)
    return res


# ---- trajectory reductions on the device (include/gala_b200.h gb_*_extrema) ---------------
def _stats_dict(stats):
    return {name: stats[k] for k, name in enumerate(_abi.EXT_ROWS)}


def integrate_extrema(hamiltonian, w0, t, Integrator="leapfrog", with_energy=False, return_final=False):
    """Integrate ``w0`` (6,N) over the grid ``t`` with Leapfrog or Ruth4 and return, per orbit, the pericentre /
    apocentre statistics of ``Orbit.pericenter`` / ``Orbit.apocenter`` (``dynamics/orbit.py:391-553``: local extrema
    of r(t_j) refined by a parabola; count, mean, min, max, first and last time of each kind), ``max |z|`` and, with
    ``with_energy``, E(t[0]), E(t[-1]) and max |E - E(t[0])| -- computed inside the integration kernel, so no
    (6, ntimes, N) trajectory is stored or copied.  Returns a dict of (N,) arrays (keys ``_abi.EXT_ROWS``), plus the
    final state (6,N) under ``"w_final"`` when asked for."""
    _check_c_enabled(hamiltonian)
    cls = get_integrator(Integrator)
    if cls not in (LeapfrogIntegrator, Ruth4Integrator):
        raise ValueError("integrate_extrema supports the fixed-step integrators (leapfrog, ruth4); for DOPRI853 reduce "
                         "its dense output with orbit_extrema")
    if cls is LeapfrogIntegrator and not isinstance(hamiltonian.frame, StaticFrame):
        raise TypeError("Leapfrog integration is currently only supported for StaticFrame, "
                        f"not {hamiltonian.frame.__class__.__name__}")
    w = _prep_w0(w0)
    th, tb = _prep_t(t, w)
    N, ntimes = w.arr.shape[1], th.size
    stats = _alloc(w, (_abi.EXT_NSTAT, N))
    wfin = _alloc(w, (6, N)) if return_final else None
    opt = _opts(w, hamiltonian)
    fr = hamiltonian.frame.spec()
    _abi.check(_abi.lib().gb_integrate_extrema(hamiltonian.potential.spec().ptr(), C.byref(fr),
                                               0 if cls is LeapfrogIntegrator else 1, w.ptr, N, tb.ptr, ntimes,
                                               int(bool(with_energy)), None if wfin is None else _abi.Buf(wfin).ptr,
                                               _abi.Buf(stats).ptr, C.byref(opt)))
    out = _stats_dict(stats)
    if return_final:
        out["w_final"] = wfin
    return out


def orbit_extrema(hamiltonian, w, t, with_energy=False):
    """The same statistics for a trajectory that already exists: ``w`` (6, ntimes, N) host array or device tensor
    (e.g. the dense output of ``dop853_integrate_hamiltonian`` left on the device)."""
    _check_c_enabled(hamiltonian)
    if _abi._is_torch_cuda(w):
        import torch
        if w.dtype != torch.float64 or w.ndim != 3 or w.shape[0] != 6:
            raise ValueError("w must be a float64 tensor of shape (6, ntimes, N)")
        buf = _abi.Buf(w.contiguous())
    else:
        a = np.ascontiguousarray(w, dtype=np.float64)
        if a.ndim == 2:
            a = a[:, :, None]
        if a.ndim != 3 or a.shape[0] != 6:
            raise ValueError(f"w must have shape (6, ntimes[, N]), got {a.shape}")
        buf = _abi.Buf(np.ascontiguousarray(a))
    th, tb = _prep_t(t, buf)
    ntimes, N = buf.arr.shape[1], buf.arr.shape[2]
    if th.size != ntimes:
        raise ValueError("t must have one entry per saved sample")
    stats = _alloc(buf, (_abi.EXT_NSTAT, N))
    opt = _abi.launch_opts(buf.device, bool(getattr(hamiltonian, "strict_math", False) or hamiltonian.potential.strict_math),
                           *( _stream_dev(buf) ))
    fr = hamiltonian.frame.spec()
    _abi.check(_abi.lib().gb_orbit_extrema(hamiltonian.potential.spec().ptr(), C.byref(fr), buf.ptr, tb.ptr, ntimes, N,
                                           int(bool(with_energy)), _abi.Buf(stats).ptr, C.byref(opt)))
    return _stats_dict(stats)


def orbit_extrema_list(w, t, kind="peri"):
    """Every refined extremum of one kind -- ``"peri"``, ``"apo"`` or ``"zmax"`` -- of every orbit of an existing trajectory
    ``w`` (6, ntimes, N), i.e. ``func=None`` of ``Orbit.pericenter / apocenter / zmax`` (``dynamics/orbit.py:439-656``).
    Returns ``(values, times)``: two lists with one array per orbit, in increasing time."""
    k = {"peri": 0, "apo": 1, "zmax": 2}[kind]
    if _abi._is_torch_cuda(w):
        buf = _abi.Buf(w.contiguous())
    else:
        a = np.ascontiguousarray(w, dtype=np.float64)
        buf = _abi.Buf(a if a.ndim == 3 else np.ascontiguousarray(a[:, :, None]))
    if buf.arr.ndim != 3 or buf.arr.shape[0] != 6:
        raise ValueError("w must have shape (6, ntimes[, N])")
    th, tb = _prep_t(t, buf)
    ntimes, N = buf.arr.shape[1], buf.arr.shape[2]
    if th.size != ntimes:
        raise ValueError("t must have one entry per saved sample")
    kmax = 16
    while True:
        vals, times = _alloc(buf, (kmax, N)), _alloc(buf, (kmax, N))
        counts = _alloc(buf, (N,), "i4")
        opt = _abi.launch_opts(buf.device, False, *_stream_dev(buf))
        _abi.check(_abi.lib().gb_orbit_extrema_list(buf.ptr, tb.ptr, ntimes, N, k, kmax, _abi.Buf(vals).ptr,
                                                    _abi.Buf(times).ptr, _abi.Buf(counts).ptr, C.byref(opt)))
        cmax = int(counts.max()) if N else 0
        if cmax <= kmax:
            break
        kmax = cmax
    if buf.device:
        vals, times, counts = vals.cpu().numpy(), times.cpu().numpy(), counts.cpu().numpy()
    return [vals[:counts[i], i].copy() for i in range(N)], [times[:counts[i], i].copy() for i in range(N)]


def _stream_dev(buf):
    """(stream, block, device) positional tail of launch_opts for a host / device buffer."""
    if buf.device:
        import torch
        return torch.cuda.current_stream(buf.arr.device).cuda_stream, 0, buf.arr.device.index
    return None, 0, -1


# ---- integrator classes (names for Hamiltonian.integrate_orbit dispatch; integrate/__init__.py) ---
class _IntegratorBase:
    name = None

    def __init__(self, func=None, func_args=(), func_units=None, progress=False, save_all=True, **kwargs):
        self.F = func
        self.save_all = save_all
        self.kwargs = kwargs


class LeapfrogIntegrator(_IntegratorBase):
    name = "leapfrog"


class Ruth4Integrator(_IntegratorBase):
    name = "ruth4"


class DOPRI853Integrator(_IntegratorBase):
    name = "dopri853"


_LOOKUP = {"leapfrog": LeapfrogIntegrator, "ruth4": Ruth4Integrator, "dopri853": DOPRI853Integrator,
           "dop853": DOPRI853Integrator}


def get_integrator(Integrator):
    """``integrate/lookup.py``: accept a class or a (case-insensitive) name."""
    if isinstance(Integrator, str):
        try:
            return _LOOKUP[Integrator.lower()]
        except KeyError:
            raise ValueError(f"Unknown integrator name '{Integrator}'. Available: {sorted(_LOOKUP)}")
    if isinstance(Integrator, type) and issubclass(Integrator, _IntegratorBase):
        return Integrator
    name = getattr(Integrator, "__name__", "")
    for cls in (LeapfrogIntegrator, Ruth4Integrator, DOPRI853Integrator):     # real gala classes by name
        if name == cls.__name__:
            return cls
    raise ValueError(f"Integrator {Integrator!r} is not supported by the B200 engine")
