"""Mock stellar streams: ``FardalStreamDF``, ``MockStreamGenerator`` and the two functions at the
reference's Cython boundary, ``mockstream_dop853`` / ``mockstream_leapfrog``
(reference ``dynamics/mockstream/df.pyx``, ``mockstream_generator.py``, ``mockstream.pyx``).

Scope: no massive bodies other than the (massless-for-dynamics)
progenitor -- i.e. ``progenitor_potential=None`` and no extra ``nbody`` -- in a StaticFrame.  Then
every stream particle is an independent test particle and the whole stream is integrated by ONE
batched kernel launch (one thread per particle) instead of the reference's sequential loop over
release groups.  Random deviates are drawn on the host with the caller's numpy RNG in exactly the
reference's order (``df.pyx:393-454``), so sampling is bit-reproducible; the GPU does the geometry.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from .dynamics import MockStream, Orbit, PhaseSpacePosition
from .frame import StaticFrame
from .hamiltonian import Hamiltonian
from .integrate import (DOPRI853Integrator, LeapfrogIntegrator, dop853_integrate_hamiltonian, get_integrator,
                        leapfrog_integrate_hamiltonian, parse_time_specification)
from .units import strip

__all__ = ["BaseStreamDF", "FardalStreamDF", "StreaklineStreamDF", "LagrangeCloudStreamDF", "ChenStreamDF",
           "MockStreamGenerator", "DirectNBody", "mockstream_dop853", "mockstream_leapfrog", "mockstream_dop853_animate"]


_SIDE = {}


def _side_stream(dev, k):
    """Per-device pool of non-blocking torch streams for the stream pipeline (created once)."""
    import torch
    key = (dev.index, k)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def _opts(H, shard=False):
    """``shard``: the entry point deals its particles over the devices of ``_abi.set_devices`` (test-particle mock
    streams); the N-body / snapshot entry points run on one device."""
    return _abi.launch_opts(False, bool(getattr(H, "strict_math", False) or H.potential.strict_math),
                            devices=None if shard else False)


class BaseStreamDF:
    """``BaseStreamDF`` (``df.pyx:28-238``): bookkeeping shared by the stream distribution functions.  A
    subclass provides ``_kind`` (the device DF id), and ``_draws(Np)`` = the random deviates of all
    particles drawn with the caller's numpy RNG in exactly the reference's order (per timestep all
    trailing particles, then all leading ones; per particle the DF's own draw order)."""
    _kind = None
    _flags = 0

    def __init__(self, lead=True, trail=True, random_state=None):
        self._lead, self._trail = int(bool(lead)), int(bool(trail))
        if not self._lead and not self._trail:
            raise ValueError("You must generate either leading or trailing tails (or both!)")
        self.random_state = np.random.RandomState() if random_state is None else random_state

    lead = property(lambda self: self._lead)
    trail = property(lambda self: self._trail)

    def _draws(self, Np, potential):
        return None

    def _plan(self, prog_m, nparticles):
        """Particle bookkeeping in the reference's emission order: per timestep (skipping
        prog_m == 0), all trailing particles, then all leading particles (``df.pyx:393-454``)."""
        prog_m = np.asarray(prog_m)
        n = np.where(prog_m == 0, 0, np.asarray(nparticles, dtype=np.int64))
        steps = np.nonzero(n)[0]
        if steps.size == 0:
            return np.zeros(0, np.int32), np.zeros(0)
        # per timestep: n trailing (+1) then n leading (-1) entries -- vectorised form of the loop
        tails = ([1.0] if self._trail else []) + ([-1.0] if self._lead else [])
        step_rep = np.repeat(steps, len(tails))                 # timestep of each (timestep, tail) block
        sign_rep = np.tile(np.array(tails), steps.size)
        counts = n[step_rep]
        return np.repeat(step_rep, counts).astype(np.int32), np.repeat(sign_rep, counts)

    def _sample(self, potential, prog_x, prog_v, prog_t, prog_m, nparticles):
        """(ntimes,3) progenitor positions / velocities -> particle_x (Np,3), particle_v, particle_t1."""
        prog_idx, sign = self._plan(prog_m, nparticles)
        Np = prog_idx.size
        draws = self._draws(Np, potential)
        ncols = 0 if draws is None else draws.shape[1]
        prog_w = np.ascontiguousarray(np.hstack([prog_x, prog_v]), dtype=np.float64)
        prog_t = np.ascontiguousarray(prog_t, dtype=np.float64)
        prog_m = np.ascontiguousarray(prog_m, dtype=np.float64)
        out = np.empty((Np, 6))
        opt = _abi.launch_opts(False, bool(potential.strict_math))
        _abi.check(_abi.lib().gb_stream_release(
            potential.spec().ptr(), float(potential.G), prog_w.ctypes.data, prog_t.ctypes.data, prog_m.ctypes.data,
            len(prog_t), prog_idx.ctypes.data, sign.ctypes.data, None if draws is None else draws.ctypes.data, ncols,
            Np, int(self._kind), int(self._flags), out.ctypes.data, C.byref(opt)))
        self._last_rows = out                # (Np, 6) rows as the integrate kernels take them: reused by run()
        return out[:, :3], out[:, 3:], prog_t[prog_idx]

    def sample(self, prog_orbit, prog_mass, hamiltonian=None, release_every=1, n_particles=1):
        """``BaseStreamDF.sample`` (``df.pyx:125-238``): returns a ``MockStream`` of initial conditions."""
        H = prog_orbit.hamiltonian if getattr(prog_orbit, "hamiltonian", None) is not None else (
            Hamiltonian(hamiltonian) if hamiltonian is not None else None)
        if H is None:
            raise ValueError("a hamiltonian is required to sample the DF")
        if not isinstance(H.frame, StaticFrame):
            raise NotImplementedError("mock streams are implemented for StaticFrame only")
        prog_x = np.ascontiguousarray(np.asarray(prog_orbit.pos).T)
        prog_v = np.ascontiguousarray(np.asarray(prog_orbit.vel).T)
        prog_t = np.asarray(prog_orbit.t, dtype=np.float64)
        prog_m = np.squeeze(np.asarray(strip(prog_mass), dtype=np.float64))
        if prog_m.shape == ():
            prog_m = np.full_like(prog_t, prog_m)
        if np.iterable(n_particles):
            n_particles = np.array(n_particles).astype("i4")
            if len(n_particles) != len(prog_t):
                raise ValueError("If passing in an array n_particles, its shape must match the number of "
                                 "timesteps in the progenitor orbit.")
        else:
            N = int(n_particles)
            n_particles = np.zeros(len(prog_t), dtype="i4")
            n_particles[::release_every] = N
        x, v, t1 = self._sample(H.potential, prog_x, prog_v, prog_t, prog_m, n_particles)
        _, sign = self._plan(prog_m, n_particles)
        lt = np.where(sign > 0, "t", "l").astype("U1")
        ms = MockStream(pos=x.T, vel=v.T, release_time=t1, lead_trail=lt, frame=H.frame)
        ms._rows = self._last_rows
        return ms


class FardalStreamDF(BaseStreamDF):
    """Fardal, Huang & Weinberg (2015) particle-release distribution function
    (``df.pyx:320-456``).  ``random_state`` may be a ``numpy.random.RandomState`` or ``Generator``;
    only ``.normal(loc, scale)`` is used: four draws per particle in the order kx, z, vt, vz."""
    _kind = 0

    def __init__(self, gala_modified=True, lead=True, trail=True, random_state=None):
        super().__init__(lead=lead, trail=trail, random_state=random_state)
        self._gala_modified = int(bool(gala_modified))
        self._flags = self._gala_modified

    def _draws(self, Np, potential):
        # k_mean / k_disp of df.pyx:378-391; one broadcast call consumes the bit stream in the scalar order
        kvt_fardal = 0.4
        loc = np.array([2.0, 0.0, 0.3, 0.0])
        scale = np.array([0.5 if self._gala_modified else 0.4, 0.5,
                          0.5 if self._gala_modified else kvt_fardal, 0.5])
        # normal(loc, scale) IS loc + scale * standard_normal() in numpy (legacy and Generator), so one
        # standard_normal((Np, 4)) call consumes the bit stream like the reference's scalar draws and the
        # in-place scale/shift reproduces their values bit for bit (tests/test_host_logic_cpu.py)
        x = self.random_state.standard_normal((Np, 4))
        x *= scale
        x += loc
        return x


class StreaklineStreamDF(BaseStreamDF):
    """Kuepper et al. (2012) "streakline" DF (``df.pyx:242-318``): particles leave exactly at the Lagrange
    radius with the progenitor's angular velocity; no random numbers."""
    _kind = 1


class LagrangeCloudStreamDF(BaseStreamDF):
    """Gibbons et al. (2014) Lagrange-cloud stripping (``df.pyx:460-552``): released at the Lagrange radius
    with an isotropic Gaussian velocity offset of dispersion ``v_disp`` (kpc/Myr here; the reference converts
    an astropy speed to the potential's units); three ``normal(0, v_disp)`` draws per particle."""
    _kind = 2

    def __init__(self, v_disp, lead=True, trail=True, random_state=None):
        super().__init__(lead=lead, trail=trail, random_state=random_state)
        self.v_disp = float(strip(v_disp))

    def _draws(self, Np, potential):
        x = self.random_state.standard_normal((Np, 3))       # normal(0, v_disp) = 0 + v_disp * standard_normal()
        x *= self.v_disp
        x += 0.0
        return x


class ChenStreamDF(BaseStreamDF):
    """Chen et al. (2024) DF (``df.pyx:556-702``): per particle one ``multivariate_normal(mean, cov)`` draw of
    (r, phi, theta, v, alpha, beta).  The reference calls it once per particle (one SVD of ``cov`` per call);
    a single ``size=Np`` call consumes the same standard normals in the same order and applies the same
    factorisation, so the rows agree with the per-particle loop to 1 ulp (matrix-matrix instead of
    matrix-vector product inside numpy; ``tests/test_host_logic_cpu.py``)."""
    _kind = 3
    mean = np.array([1.6, -30.0, 0.0, 1.0, 20.0, 0.0])
    cov = np.diag([0.1225, 529.0, 144.0, 0.0, 400.0, 484.0])
    cov[0, 4] = cov[4, 0] = -4.9

    def _draws(self, Np, potential):
        if Np == 0:
            return np.zeros((0, 6))
        return np.ascontiguousarray(self.random_state.multivariate_normal(self.mean, self.cov, size=Np), dtype=np.float64)


def _same_potential(a, b):
    if a is b:
        return True
    ca, cb = a._components(), b._components()
    return len(ca) == len(cb) and all(
        x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(np.zeros(3) if x[2] is None else x[2],
                                                                       np.zeros(3) if y[2] is None else y[2])
        for x, y in zip(ca, cb))


def _is_null(pp):
    return pp is None or pp.__class__.__name__ == "NullPotential"


class _BodySpec:
    """ctypes view of the per-body potentials (``gb_bodies``); bodies at index >= ``n_sources`` are passed
    as massless (``leapfrog_integrate_nbody`` only loops over the first ``nbody`` = number-of-non-Null
    potentials as force sources, leapfrog.pyx:200-202,236-247)."""

    def __init__(self, particle_potentials, n_sources=None):
        nb = len(particle_potentials)
        if nb < 1 or nb > 16:
            raise NotImplementedError("the B200 N-body kernels carry 1..16 bodies (massive or massless) per system")
        self._keep = []
        self.pots = (_abi.gb_potential * nb)()
        for b, pp in enumerate(particle_potentials):
            if _is_null(pp) or (n_sources is not None and b >= n_sources):
                self.pots[b].n_components = 0
                self.pots[b].n_dim = 3
                continue
            sp = pp.spec()
            self._keep.append(sp)
            self.pots[b] = sp.pot
        self.struct = _abi.gb_bodies(nb, 0, C.cast(self.pots, C.POINTER(_abi.gb_potential)))

    def ptr(self):
        return C.byref(self.struct)


def _nbody_leapfrog(H, pps, body_w0, t0, tfinal, nsteps, dt, w0_rows=None, t1=None, group=None, body_writer=0,
                    save_all=False, n_sources=None, scheme=0):
    """One ``gb_nbody_leapfrog`` call.  Returns (particles (Np,6) | None, bodies (nb,6), traj | None)."""
    bs = _BodySpec(pps, n_sources)
    body_w0 = np.ascontiguousarray(body_w0, dtype=np.float64).reshape(-1, len(pps), 6)
    Np = 0 if w0_rows is None else w0_rows.shape[0]
    rows = np.ascontiguousarray(w0_rows, dtype=np.float64) if Np else None
    out_p = np.empty((Np, 6)) if Np else None
    out_b = np.empty((len(pps), 6))
    traj = np.empty((nsteps + 1, len(pps) + Np, 6)) if save_all else None
    t1a = None if t1 is None else np.ascontiguousarray(t1, dtype=np.float64)
    grp = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    opt = _opts(H)
    ptr = lambda a: None if a is None else a.ctypes.data
    _abi.check(_abi.lib().gb_nbody_leapfrog(H.potential.spec().ptr(), bs.ptr(), body_w0.ctypes.data, body_w0.shape[0],
                                            ptr(grp), ptr(rows), ptr(t1a), Np, float(t0), float(tfinal), int(nsteps),
                                            float(dt), int(scheme), ptr(out_p), out_b.ctypes.data, int(body_writer), ptr(traj),
                                            C.byref(opt)))
    return out_p, out_b, traj


def _nbody_dop853(H, pps, body_w0, tgrid, tfinal, dt0, step_mode, w0_rows=None, t1=None, group=None, body_writer=0,
                  save_all=False, atol=1e-10, rtol=1e-10, nmax=0, dt_max=0.0, err_if_fail=1):
    """One ``gb_nbody_dop853`` call.  Returns (particles | None, bodies, traj | None, status)."""
    bs = _BodySpec(pps)
    body_w0 = np.ascontiguousarray(body_w0, dtype=np.float64).reshape(-1, len(pps), 6)
    Np = 0 if w0_rows is None else w0_rows.shape[0]
    rows = np.ascontiguousarray(w0_rows, dtype=np.float64) if Np else None
    out_p = np.empty((Np, 6)) if Np else None
    out_b = np.empty((len(pps), 6))
    tg = None if tgrid is None else np.ascontiguousarray(tgrid, dtype=np.float64)
    traj = np.empty((tg.size, len(pps) + Np, 6)) if save_all else None
    status = np.empty(max(Np, 1), dtype=np.int32)
    t1a = None if t1 is None else np.ascontiguousarray(t1, dtype=np.float64)
    grp = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    opt = _opts(H)
    ptr = lambda a: None if a is None else a.ctypes.data
    rc = _abi.lib().gb_nbody_dop853(H.potential.spec().ptr(), bs.ptr(), body_w0.ctypes.data, body_w0.shape[0], ptr(grp),
                                    ptr(rows), ptr(t1a), Np, ptr(tg), 0 if tg is None else tg.size, float(tfinal),
                                    float(dt0), float(atol), float(rtol), int(nmax), float(dt_max), int(step_mode),
                                    ptr(out_p), out_b.ctypes.data, int(body_writer), ptr(traj), status.ctypes.data,
                                    C.byref(opt))
    if rc in (-1, -2, -3, -4):
        if err_if_fail:
            _abi.check(rc)
    else:
        _abi.check(rc)
    return out_p, out_b, traj, status


class DirectNBody:
    """``gala.dynamics.nbody.DirectNBody`` (``dynamics/nbody/core.py``): initial conditions of the bodies,
    one potential per body (``None`` = test particle) and the external Hamiltonian.  Up to 16 massive
    bodies plus any number of test particles run on the device (one lane = the massive bodies + one test
    particle, ``csrc/nbody.cuh``); without massive bodies every body is an independent orbit."""

    def __init__(self, w0, particle_potentials, external_potential=None, frame=None, units=None, save_all=True):
        w = w0.w() if isinstance(w0, PhaseSpacePosition) else np.asarray(w0, dtype=np.float64)
        w = w.reshape(6, -1)
        self._c_w0 = np.ascontiguousarray(w.T)            # (nbodies, 6) like the reference
        self.particle_potentials = [None if _is_null(p) else p for p in particle_potentials]
        if len(self.particle_potentials) != self._c_w0.shape[0]:
            raise ValueError("The number of initial conditions in `w0` must match the number of particle "
                             "potentials passed in with `particle_potentials`.")
        if external_potential is None:
            from .potential import NullPotential
            external_potential = NullPotential()
        self.H = Hamiltonian(external_potential, frame)
        self.external_potential, self.frame, self.units = self.H.potential, self.H.frame, self.H.units
        self.save_all = save_all

    @property
    def n_massive(self):
        return sum(p is not None for p in self.particle_potentials)

    def integrate_orbit(self, Integrator=None, Integrator_kwargs=None, **time_spec):
        """``nbody/core.py:166-282``: massive bodies are moved to the front, integrated together with the
        test particles, and the original order is restored.  DOPRI853 runs with the stiffness test
        disabled (``nbody.pyx:106``)."""
        Integrator = get_integrator(Integrator or DOPRI853Integrator)
        kw = dict(Integrator_kwargs or {})
        t = parse_time_specification(self.units, **time_spec)
        from .integrate import Ruth4Integrator, ruth4_integrate_hamiltonian
        if Integrator not in (LeapfrogIntegrator, DOPRI853Integrator, Ruth4Integrator):
            raise NotImplementedError(f"N-body integration is currently not supported with the {Integrator} "
                                      "integrator class")
        kw = {k: kw[k] for k in ("atol", "rtol", "nmax", "dt_max", "err_if_fail") if k in kw}
        if self.n_massive == 0:
            w0 = np.ascontiguousarray(self._c_w0.T)
            if Integrator is LeapfrogIntegrator:
                _, w = leapfrog_integrate_hamiltonian(self.H, w0, t, save_all=int(self.save_all))
            elif Integrator is Ruth4Integrator:
                _, w = ruth4_integrate_hamiltonian(self.H, w0, t, save_all=int(self.save_all))
            else:
                _, w = dop853_integrate_hamiltonian(self.H, w0, t, nstiff=-1, save_all=int(self.save_all), **kw)
        else:
            if not isinstance(self.frame, StaticFrame):
                raise TypeError("N-body integration with massive bodies is only supported for StaticFrame")
            front = [i for i, pp in enumerate(self.particle_potentials) if pp is not None]
            end = [i for i, pp in enumerate(self.particle_potentials) if pp is None]
            idx = np.array(front + end)
            pps = [self.particle_potentials[i] for i in front]
            body_w0 = self._c_w0[front]
            rows = self._c_w0[end] if end else None
            dt = t[1] - t[0]
            if Integrator in (LeapfrogIntegrator, Ruth4Integrator):
                out_p, out_b, traj = _nbody_leapfrog(self.H, pps, body_w0, t[0], t[-1], len(t) - 1, dt, w0_rows=rows,
                                                     save_all=self.save_all, scheme=int(Integrator is Ruth4Integrator))
            else:
                out_p, out_b, traj, _ = _nbody_dop853(self.H, pps, body_w0, t, t[-1], dt, 0, w0_rows=rows,
                                                      save_all=self.save_all, **kw)
            undo = np.argsort(idx)
            if self.save_all:
                w = np.ascontiguousarray(traj[:, undo, :].transpose(2, 0, 1))      # (6, ntimes, N)
            else:
                ws = out_b if out_p is None else np.vstack([out_b, out_p])
                w = np.ascontiguousarray(ws[undo].T)
        if self.save_all:
            return Orbit.from_w(w, t=t, hamiltonian=self.H)
        return PhaseSpacePosition.from_w(w, frame=self.frame)


def _validate_time_arrays(ntimes, n_t1, n_nstream):
    if n_t1 != ntimes or n_nstream != ntimes:
        raise ValueError("time, stream_t1 and nstream must have the same length")


def mockstream_dop853(nbody, time, stream_w0, stream_t1, tfinal, nstream, atol=1e-10, rtol=1e-10, nmax=0,
                      dt_max=0.0, nstiff=-1, progress=0, err_if_fail=1, log_output=0):
    """``mockstream.pyx:176-303``.  Returns ``(nbody_w (nbodies,6), stream_w (Np,6))``.
    Deviation from the reference (documented in DESIGN.md): the members of a release group do not
    share one DOP853 step size; each particle is its own n=6 system."""
    time = np.ascontiguousarray(time, dtype=np.float64)
    stream_w0 = np.ascontiguousarray(stream_w0, dtype=np.float64)
    stream_t1 = np.ascontiguousarray(stream_t1, dtype=np.float64)
    nstream = np.asarray(nstream)
    ntimes = time.shape[0]
    _validate_time_arrays(ntimes, stream_t1.shape[0], nstream.shape[0])
    if stream_w0.shape != (int(nstream.sum()), 6):
        raise ValueError("stream_w0 must have shape (sum(nstream), 6)")
    H = nbody.H
    nbodies = nbody._c_w0.shape[0]
    dt0 = time[1] - time[0]
    if nbody.n_massive:
        # massive bodies: every lane integrates [the bodies, one stream particle] (csrc/nbody.cuh)
        pps = nbody.particle_potentials
        _, _, traj, _ = _nbody_dop853(H, pps, nbody._c_w0, time, time[-1], dt0, 0, save_all=True, atol=atol, rtol=rtol,
                                      nmax=nmax, dt_max=dt_max, err_if_fail=err_if_fail)      # mockstream.pyx:247-255
        groups = np.nonzero(nstream)[0]
        if groups.size == 0:
            return traj[-1].copy(), np.empty((0, 6))
        group = np.repeat(np.arange(ntimes, dtype=np.int32), nstream)
        t1 = np.repeat(stream_t1, nstream)
        writer = int(np.nonzero(group == groups[-1])[0][0])
        out_p, out_b, _, _ = _nbody_dop853(H, pps, traj, None, tfinal, dt0, 1, w0_rows=stream_w0, t1=t1, group=group,
                                           body_writer=writer, atol=atol, rtol=rtol, nmax=nmax, err_if_fail=err_if_fail)
        return out_b, out_p
    # 1) the bodies at every release time: dense output of one DOP853 run (mockstream.pyx:247-255)
    _, nbody_w = dop853_integrate_hamiltonian(H, np.ascontiguousarray(nbody._c_w0.T), time, atol=atol, rtol=rtol,
                                              nmax=nmax, dt_max=dt_max, nstiff=nstiff, save_all=1,
                                              err_if_fail=err_if_fail)
    # 2) every stream particle from its release time to tfinal, plus the bodies from the LAST group's
    #    release time (that is what w_tmp[:nbodies] holds when the reference's loop ends, :285-290)
    groups = np.nonzero(nstream)[0]
    last = groups[-1] if groups.size else ntimes - 1
    rows = np.vstack([stream_w0, nbody_w[:, last, :].T])
    t1 = np.concatenate([np.repeat(stream_t1, nstream), np.full(nbodies, stream_t1[last])])
    out = np.empty_like(rows)
    status = np.empty(rows.shape[0], dtype=np.int32)
    opt = _opts(H, shard=True)
    fr = H.frame.spec()
    rc = _abi.lib().gb_mockstream_dop853(H.potential.spec().ptr(), C.byref(fr), rows.ctypes.data, t1.ctypes.data,
                                         rows.shape[0], float(tfinal), float(dt0), float(atol), float(rtol),
                                         int(nmax), out.ctypes.data, status.ctypes.data, C.byref(opt))
    if rc in (-1, -2, -3, -4):
        if err_if_fail:
            _abi.check(rc)
    else:
        _abi.check(rc)
    Np = stream_w0.shape[0]
    return out[Np:], out[:Np]


def _write_snapshots(filename, overwrite, groups, units_name):
    """The reference's on-disk layout (mockstream.pyx:107-173): groups ``stream`` / ``nbody`` with datasets
    ``pos``, ``vel`` of shape (3, n_out, n) and ``time``.  HDF5 (gzip-9, NaN fill value, ``unit`` attributes)
    when h5py is importable; otherwise the same arrays go to ``<filename>.npz`` under the keys
    ``stream/pos`` ... ``nbody/time`` (this image has no h5py)."""
    import os
    try:
        import h5py
    except ImportError:
        h5py = None
    target = str(filename) if h5py is not None else str(filename) + ".npz"
    if os.path.exists(target) and not overwrite:
        raise IOError(f"Mockstream output file {target} already exists! Use overwrite=True to overwrite the file.")
    if h5py is None:
        np.savez_compressed(target, **{f"{g}/{k}": v for g, d in groups.items() for k, v in d.items()})
        return target
    with h5py.File(target, "w") as f:
        for g, d in groups.items():
            grp = f.create_group(g)
            for k, v in d.items():
                if k == "time":
                    ds = grp.create_dataset(k, data=v)
                else:
                    ds = grp.create_dataset(k, data=v, dtype="f8", fillvalue=np.nan, compression="gzip", compression_opts=9)
                ds.attrs["unit"] = units_name.get(k, "")
    return target


def mockstream_dop853_animate(nbody, t, stream_w0, nstream, output_every=1, output_filename="", overwrite=False,
                              check_filesize=True, atol=1e-10, rtol=1e-10, nmax=0, dt_max=0.0, nstiff=-1, progress=0,
                              err_if_fail=1, log_output=0):
    """``mockstream.pyx:306-440``: march the stream over ``t`` interval by interval, storing a snapshot every
    ``output_every`` intervals (and the final state) in ``output_filename``.  Returns
    ``(nbody_w (nbodies,6), stream_w (Np,6))`` like the reference; the snapshot arrays are also kept on the
    function object as ``mockstream_dop853_animate.last`` (dict of the two groups)."""
    import warnings
    t = np.ascontiguousarray(t, dtype=np.float64)
    stream_w0 = np.ascontiguousarray(stream_w0, dtype=np.float64)
    nstream = np.asarray(nstream)
    ntimes = t.shape[0]
    if nstream.shape[0] != ntimes:
        raise ValueError("nstream must have one entry per time")
    if stream_w0.shape != (int(nstream.sum()), 6):
        raise ValueError("stream_w0 must have shape (sum(nstream), 6)")
    H = nbody.H
    nbodies = nbody._c_w0.shape[0]
    Np = stream_w0.shape[0]
    nout = (ntimes - 1) // output_every + 1 + (1 if (ntimes - 1) % output_every else 0)
    if Np * nout * 8 >= 8e9 and check_filesize:
        warnings.warn("Estimated mockstream output file is expected to be >8 GB in size! If you're sure, turn this "
                      "warning off with `check_filesize=False`")
    rows = np.vstack([nbody._c_w0, stream_w0])
    ridx = np.concatenate([np.zeros(nbodies, dtype=np.int32), np.repeat(np.arange(ntimes, dtype=np.int32), nstream)])
    snap = np.empty((nout, rows.shape[0], 6))
    fin = np.empty_like(rows)
    status = np.empty(rows.shape[0], dtype=np.int32)
    opt = _opts(H)
    fr = H.frame.spec()
    if nbody.n_massive:
        # massive bodies: bodies-only march + one lane per particle = [bodies, particle] (csrc/nbody.cuh)
        bs = _BodySpec(nbody.particle_potentials)
        sw0 = np.ascontiguousarray(stream_w0)
        pr = np.ascontiguousarray(ridx[nbodies:])
        out_p = np.empty((Np, 6)); out_b = np.empty((nbodies, 6))
        rc = _abi.lib().gb_nbody_dop853_animate(H.potential.spec().ptr(), bs.ptr(), np.ascontiguousarray(nbody._c_w0).ctypes.data,
                                                sw0.ctypes.data, pr.ctypes.data, Np, t.ctypes.data, ntimes, float(atol),
                                                float(rtol), int(nmax), int(output_every), snap.ctypes.data,
                                                out_p.ctypes.data, out_b.ctypes.data, status.ctypes.data, C.byref(opt))
        fin = np.vstack([out_b, out_p])
    else:
        rc = _abi.lib().gb_mockstream_dop853_animate(H.potential.spec().ptr(), C.byref(fr), rows.ctypes.data, ridx.ctypes.data,
                                                     rows.shape[0], t.ctypes.data, ntimes, float(atol), float(rtol), int(nmax),
                                                     int(output_every), snap.ctypes.data, fin.ctypes.data, status.ctypes.data,
                                                     C.byref(opt))
    if rc in (-1, -2, -3, -4):
        if err_if_fail:
            _abi.check(rc)
    else:
        _abi.check(rc)
    out_i = [i for i in range(ntimes) if i == 0 or i % output_every == 0 or i == ntimes - 1]
    times = t[out_i]
    sw = snap.transpose(2, 0, 1)                                    # (6, nout, n)
    groups = {"stream": {"pos": np.ascontiguousarray(sw[:3, :, nbodies:]), "vel": np.ascontiguousarray(sw[3:, :, nbodies:]),
                         "time": times},
              "nbody": {"pos": np.ascontiguousarray(sw[:3, :, :nbodies]), "vel": np.ascontiguousarray(sw[3:, :, :nbodies]),
                        "time": times}}
    mockstream_dop853_animate.last = groups
    if output_filename:
        mockstream_dop853_animate.last_file = _write_snapshots(output_filename, overwrite, groups,
                                                               {"pos": "kpc", "vel": "kpc / Myr", "time": "Myr"})
    return fin[:nbodies].copy(), fin[nbodies:].copy()


def mockstream_leapfrog(nbody, full_time, spawn_time, stream_w0, stream_t1, tfinal, nstream, progress=0,
                        err_if_fail=1):
    """``mockstream.pyx:442-620``: fixed-step version; particle p takes
    ``int((tfinal - t1)/dt + 0.5)`` steps of ``dt = full_time[1]-full_time[0]``."""
    full_time = np.ascontiguousarray(full_time, dtype=np.float64)
    stream_w0 = np.ascontiguousarray(stream_w0, dtype=np.float64)
    stream_t1 = np.ascontiguousarray(stream_t1, dtype=np.float64)
    nstream = np.asarray(nstream)
    ntimes = len(spawn_time)
    _validate_time_arrays(ntimes, stream_t1.shape[0], nstream.shape[0])
    H = nbody.H
    nbodies = nbody._c_w0.shape[0]
    dt = full_time[1] - full_time[0]
    if nbody.n_massive:
        pps = nbody.particle_potentials
        # leapfrog_integrate_nbody over the full grid (mockstream.pyx:528-531); only the first n_massive
        # potentials act as sources there (leapfrog.pyx:200-202)
        _, _, full = _nbody_leapfrog(H, pps, nbody._c_w0, full_time[0], full_time[-1], len(full_time) - 1, dt,
                                     save_all=True, n_sources=nbody.n_massive)
        idx = ((stream_t1 - full_time[0]) / dt + 0.5).astype(np.int64)              # mockstream.pyx:548
        nbody_w = full[idx]                                                         # (ntimes, nbodies, 6)
        groups = np.nonzero(nstream)[0]
        if groups.size == 0:
            raise ValueError("no stream particles to integrate")
        group = np.repeat(np.arange(ntimes, dtype=np.int32), nstream)
        t1 = np.repeat(stream_t1, nstream)
        writer = int(np.nonzero(group == groups[-1])[0][0])
        out_p, out_b, _ = _nbody_leapfrog(H, pps, nbody_w, 0.0, tfinal, 0, dt, w0_rows=stream_w0, t1=t1, group=group,
                                          body_writer=writer)
        return out_b, out_p
    _, traj = leapfrog_integrate_hamiltonian(H, np.ascontiguousarray(nbody._c_w0.T), full_time, save_all=1)
    groups = np.nonzero(nstream)[0]
    last = groups[-1] if groups.size else ntimes - 1
    idx = int((stream_t1[last] - full_time[0]) / dt + 0.5)          # mockstream.pyx:548
    rows = np.vstack([stream_w0, traj[:, idx, :].T])
    t1 = np.concatenate([np.repeat(stream_t1, nstream), np.full(nbodies, stream_t1[last])])
    out = np.empty_like(rows)
    opt = _opts(H, shard=True)
    _abi.check(_abi.lib().gb_mockstream_leapfrog(H.potential.spec().ptr(), rows.ctypes.data, t1.ctypes.data,
                                                 rows.shape[0], float(tfinal), float(dt), out.ctypes.data,
                                                 C.byref(opt)))
    Np = stream_w0.shape[0]
    return out[Np:], out[:Np]


def _scatter_isclose(orbit_t, unq_t1s, nstream):
    """``all_nstream[np.isclose(orbit_t, t1)] = n`` for every (t1, n) (mockstream_generator.py:297-299)
    without the O(ntimes^2) loop: candidates come from a sorted search, the same ``np.isclose`` predicate
    (rtol 1e-5, atol 1e-8, relative to t1) decides, later t1 overwrite earlier ones as in the loop."""
    all_nstream = np.zeros(len(orbit_t), dtype=int)
    order = np.argsort(orbit_t, kind="stable")
    ts = orbit_t[order]
    tol = 1e-8 + 1e-5 * np.abs(unq_t1s)
    lo = np.searchsorted(ts, unq_t1s - tol, side="left")
    hi = np.searchsorted(ts, unq_t1s + tol, side="right")
    if np.all(hi - lo <= 1):
        hit = hi > lo
        cand = order[lo[hit]]
        ok = np.isclose(orbit_t[cand], unq_t1s[hit])
        all_nstream[cand[ok]] = nstream[hit][ok]
        return all_nstream
    for t1, n in zip(unq_t1s, nstream):          # overlapping tolerance windows: keep the loop's semantics
        all_nstream[np.isclose(orbit_t, t1)] = n
    return all_nstream


class MockStreamGenerator:
    """``dynamics/mockstream/mockstream_generator.py``: orchestration is unchanged; the three heavy
    steps (progenitor orbit, particle release, stream integration) each run as one GPU call."""

    def __init__(self, df, hamiltonian, progenitor_potential=None):
        if not hasattr(df, "sample"):
            raise TypeError("The input distribution function (DF) instance must be a stream DF")
        self.df = df
        self.hamiltonian = Hamiltonian(hamiltonian)
        if progenitor_potential is not None and not hasattr(progenitor_potential, "spec"):
            raise TypeError("If specified, the progenitor_potential must be a gala.potential class instance.")
        self.progenitor_potential = progenitor_potential
        self.self_gravity = progenitor_potential is not None

    def _get_nbody(self, prog_w0, nbody):
        """``mockstream_generator.py:79-117``: the progenitor becomes body 0 of a DirectNBody, followed by
        the bodies of the caller's ``nbody`` (perturbers)."""
        pw = prog_w0.w() if isinstance(prog_w0, PhaseSpacePosition) else np.asarray(prog_w0, dtype=np.float64)
        pw = pw.reshape(6, -1)
        pps = [self.progenitor_potential]
        if nbody is not None:
            if not _same_potential(nbody.external_potential, self.hamiltonian.potential):
                raise ValueError("The external potential of the input nbody instance must match the potential of the "
                                 "mock stream input hamiltonian!")
            pw = np.hstack([pw, nbody._c_w0.T])
            pps = pps + list(nbody.particle_potentials)
        return DirectNBody(PhaseSpacePosition.from_w(pw, frame=self.hamiltonian.frame), pps,
                           external_potential=self.hamiltonian.potential, frame=self.hamiltonian.frame,
                           units=self.hamiltonian.units)

    # ---- the test-particle stream as one device-resident pipeline ------------------------------------------
    def _run_pipelined(self, prog_w0, prog_mass, release_every, n_particles, Integrator, Integrator_kwargs, t):
        """``run`` for the common case (no massive bodies, no snapshots), same numbers, different schedule.
        The step-by-step form below costs host RNG (the reference's stream, 4 normals per Fardal particle) + two
        single-lane progenitor orbits + release + the particle kernel one after the other (17.7 ms for C3,
        profiles/bench_r1_final_c3.json).  Here everything stays on the device and overlaps:
          * a worker thread draws the deviates chunk by chunk (numpy releases the GIL) while the progenitor
            orbit runs; consecutive draws from the caller's RandomState / Generator are the same bit stream as
            one big draw;
          * particle rows are in release order = longest integration first, so chunk 0 (the earliest releases)
            starts integrating while the deviates of the later chunks are still being drawn; every chunk has its
            own stream (each is a fraction of one wave), the kernels overlap;
          * the progenitor's second integration (the reference re-integrates the bodies forward from the
            earliest state, mockstream.pyx:528-548 / :247-255 -- it decides the returned progenitor state) runs on a
            side stream next to the particle kernels;
          * one D2H of the (Np, 6) result at the end.
        Returns (stream rows (Np,6) host, prog row (1,6) host, release_time, lead_trail)."""
        import queue
        import threading
        import time
        import torch
        trace = [] if os.environ.get("GB_STREAM_TRACE") else None
        t_start = time.perf_counter()
        mark = (lambda what: trace.append((what, (time.perf_counter() - t_start) * 1e3))) if trace is not None else (lambda what: None)
        H = self.hamiltonian
        dev = torch.device("cuda", torch.cuda.current_device())
        kw = {k: Integrator_kwargs[k] for k in ("atol", "rtol", "nmax", "dt_max", "err_if_fail") if k in Integrator_kwargs}
        lf = Integrator is LeapfrogIntegrator
        pw = prog_w0.w() if isinstance(prog_w0, PhaseSpacePosition) else np.asarray(prog_w0, dtype=np.float64)
        pw = np.ascontiguousarray(np.asarray(pw, dtype=np.float64).reshape(6, 1))
        backward = t[1] < t[0]
        orbit_t = np.ascontiguousarray(t[::-1] if backward else t, dtype=np.float64)
        ntimes = orbit_t.size
        # ---- host bookkeeping that needs no GPU result (BaseStreamDF.sample / run, same code paths) ----------
        prog_m = np.squeeze(np.asarray(strip(prog_mass), dtype=np.float64))
        if prog_m.shape == ():
            prog_m = np.full(ntimes, float(prog_m))
        if np.iterable(n_particles):
            npart = np.array(n_particles).astype("i4")
            if len(npart) != ntimes:
                raise ValueError("If passing in an array n_particles, its shape must match the number of "
                                 "timesteps in the progenitor orbit.")
        else:
            npart = np.zeros(ntimes, dtype="i4")
            npart[::release_every] = int(n_particles)
        prog_idx, sign = self.df._plan(prog_m, npart)
        Np = prog_idx.size
        if Np == 0:
            raise ValueError("no stream particles to integrate")
        # ---- deviates: drawn on a worker thread, piece by piece, in the reference's order; started first, it is the
        # longest host-side item (4 legacy-gauss normals per Fardal particle: ~17 ns each) ---------------------------
        # the adaptive particle call reads its status array back (it returns the worst code), so its kernel chunks
        # cannot overlap one another: one kernel chunk there, the deviates still arrive in cache-sized pieces
        npiece = max(1, min(int(os.environ.get("GB_STREAM_CHUNKS", "8")), (Np + 4095) // 4096))
        bounds = [(Np * c) // npiece for c in range(npiece + 1)]
        group = 1 if lf else npiece               # pieces per kernel chunk
        has_draws = type(self.df)._draws is not BaseStreamDF._draws
        q = queue.Queue()

        def draw_all():
            try:
                for c in range(npiece):
                    n = bounds[c + 1] - bounds[c]
                    q.put(self.df._draws(n, H.potential) if has_draws else None)
                    mark(f"draws {c} ready")
            except BaseException as e:           # surfaces on the main thread
                q.put(e)
        worker = threading.Thread(target=draw_all, daemon=True)
        worker.start()
        # ---- progenitor orbit: launched before the rest of the host bookkeeping (asynchronous for the fixed-step
        # integrator; the adaptive call returns when its status is back) -------------------------------------------
        w0_dev = torch.as_tensor(pw, device=dev)
        if lf:
            _, traj = leapfrog_integrate_hamiltonian(H, w0_dev, t, save_all=1)
        else:
            _, traj = dop853_integrate_hamiltonian(H, w0_dev, t, nstiff=-1, save_all=1, **kw)
        mark("progenitor issued")
        release_time = orbit_t[prog_idx]
        lead_trail = np.where(sign > 0, "t", "l").astype("U1")
        unq_t1s, nstream = np.unique(release_time, return_counts=True)
        all_nstream = _scatter_isclose(orbit_t, unq_t1s, nstream)
        nstream_idx = np.where(all_nstream != 0)[0]
        if 0 not in nstream_idx:
            nstream_idx = np.insert(nstream_idx, 0, 0)
            unq_t1s = np.insert(unq_t1s, 0, orbit_t[0])
        ns = all_nstream[nstream_idx]
        if int(ns.sum()) != Np:
            raise ValueError("stream_w0 must have shape (sum(nstream), 6)")
        t1_rows = np.repeat(unq_t1s, ns)                  # what mockstream_* pair with the rows (== release_time)
        groups = np.nonzero(ns)[0]
        last = int(groups[-1])
        tfinal = float(orbit_t[-1])
        # fixed step: dt of the full grid (mockstream.pyx:500); adaptive: initial step = spacing of the release-time
        # grid handed to mockstream_dop853 (time[1] - time[0], mockstream.pyx:228)
        tgrid = orbit_t if lf else np.ascontiguousarray(orbit_t[nstream_idx])
        dt0 = float(tgrid[1] - tgrid[0]) if tgrid.size > 1 else float(orbit_t[1] - orbit_t[0])
        mark("host plan done")
        traj = traj[:, :, 0]                                                  # (6, ntimes)
        if backward:
            traj = torch.flip(traj, dims=[1])                                # earliest state first (run :237-250)
        prog_rows = traj.t().contiguous()                                     # (ntimes, 6) = hstack([prog_x, prog_v])
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
        d_t, d_m, d_idx, d_sign, d_t1 = up(orbit_t), up(prog_m), up(prog_idx), up(sign), up(t1_rows)
        rows0 = torch.empty((Np, 6), dtype=torch.float64, device=dev)
        rows1 = torch.empty((Np + 1, 6), dtype=torch.float64, device=dev)    # + the progenitor row (nbodies = 1)
        status = torch.empty((Np + 1,), dtype=torch.int32, device=dev)
        main = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(main)
        # ---- second progenitor integration on a side stream (it only feeds the returned progenitor state) -----
        side = _side_stream(dev, 0)
        side.wait_event(ready)
        lib, pot = _abi.lib(), H.potential.spec().ptr()
        fr = H.frame.spec()
        strict = bool(getattr(H, "strict_math", False) or H.potential.strict_math)
        def second_progenitor():
            # the reference re-integrates the bodies forward from the earliest state (mockstream.pyx:528-548, :247-255);
            # only the returned progenitor state depends on it
            with torch.cuda.stream(side):
                nb0 = prog_rows[0].reshape(6, 1).contiguous()
                if lf:
                    _, full = leapfrog_integrate_hamiltonian(H, nb0, orbit_t, save_all=1)
                    idx = int((unq_t1s[last] - orbit_t[0]) / (orbit_t[1] - orbit_t[0]) + 0.5)          # mockstream.pyx:548
                    body0 = full[:, idx, 0]
                else:
                    akw = {k: v for k, v in Integrator_kwargs.items() if k in ("atol", "rtol", "nmax", "dt_max", "nstiff", "err_if_fail")}
                    akw.setdefault("nstiff", -1)
                    _, full = dop853_integrate_hamiltonian(H, nb0, np.ascontiguousarray(orbit_t[nstream_idx]), save_all=1, **akw)
                    body0 = full[:, last, 0]
                prow = body0.reshape(1, 6).contiguous()
                pt1 = torch.full((1,), float(unq_t1s[last]), dtype=torch.float64, device=dev)
                self._mock_integrate(lib, pot, fr, lf, prow, pt1, 1, tfinal, dt0, Integrator_kwargs, rows1[Np:], status[Np:],
                                     strict, side, dev)
                ev = torch.cuda.Event()
                ev.record(side)
                return ev
        done = []
        side_thread, side_result = None, {}
        if lf:
            done.append(second_progenitor())          # asynchronous launches
        else:
            # the adaptive calls block until their status is back: run this chain on its own host thread (ctypes and
            # torch release the GIL) next to the particle call below
            def side_main():
                try:
                    torch.cuda.set_device(dev)
                    side_result["ev"] = second_progenitor()
                except BaseException as e:
                    side_result["err"] = e
            side_thread = threading.Thread(target=side_main, daemon=True)
            side_thread.start()
        mark("progenitor + side stream issued")
        # ---- kernel chunks: deviates -> release -> integrate, each on its own stream ------------------------------
        pieces = []
        for c in range(npiece):
            d = q.get()
            if isinstance(d, BaseException):
                raise d
            pieces.append(d)
            if (c + 1) % group and c + 1 < npiece:
                continue
            a, b = bounds[c + 1 - len(pieces)], bounds[c + 1]
            d = None if pieces[0] is None else (pieces[0] if len(pieces) == 1 else np.concatenate(pieces))
            pieces = []
            st = _side_stream(dev, 1 + c % 4)
            st.wait_event(ready)
            with torch.cuda.stream(st):
                ncols = 0
                d_dr = None
                if d is not None:
                    ncols = d.shape[1]
                    d_dr = torch.as_tensor(d, device=dev)
                opt = _abi.launch_opts(True, strict, st.cuda_stream, device=dev.index)
                _abi.check(lib.gb_stream_release(pot, float(H.potential.G), prog_rows.data_ptr(), d_t.data_ptr(), d_m.data_ptr(),
                                                 ntimes, d_idx[a:b].data_ptr(), d_sign[a:b].data_ptr(),
                                                 None if d_dr is None else d_dr.data_ptr(), ncols, b - a, int(self.df._kind),
                                                 int(self.df._flags), rows0[a:b].data_ptr(), C.byref(opt)))
                self._mock_integrate(lib, pot, fr, lf, rows0[a:b], d_t1[a:b], b - a, tfinal, dt0, Integrator_kwargs,
                                     rows1[a:b], status[a:b], strict, st, dev)
                if d_dr is not None:
                    d_dr.record_stream(st)
            ev = torch.cuda.Event()
            ev.record(st)
            done.append(ev)
            mark(f"chunk ending at piece {c} issued")
        if side_thread is not None:
            side_thread.join()
            if "err" in side_result:
                raise side_result["err"]
            done.append(side_result["ev"])
        for ev in done:
            main.wait_event(ev)
        out = torch.empty((Np + 1, 6), dtype=torch.float64, pin_memory=True) if Np > 65536 else None
        if out is not None:
            out.copy_(rows1, non_blocking=True)
            main.synchronize()
            res = out.numpy()
        else:
            res = rows1.cpu().numpy()
        worker.join()
        mark("result on the host")
        if trace is not None:
            import sys
            print("[stream pipeline ms] " + ", ".join(f"{w} {ms:.2f}" for w, ms in trace), file=sys.stderr)
        if not lf and kw.get("err_if_fail", 1):
            worst = int(status.min().item())
            if worst < 0:
                raise RuntimeError(f"Integration failed with code {worst}")
        return res[:Np], res[Np:], release_time, lead_trail

    @staticmethod
    def _mock_integrate(lib, pot, fr, lf, rows, t1, n, tfinal, dt0, Integrator_kwargs, out, status, strict, st, dev):
        opt = _abi.launch_opts(True, strict, st.cuda_stream, device=dev.index)
        if lf:
            _abi.check(lib.gb_mockstream_leapfrog(pot, rows.data_ptr(), t1.data_ptr(), n, float(tfinal),
                                                  dt0, out.data_ptr(), C.byref(opt)))
            return
        kw = Integrator_kwargs
        rc = lib.gb_mockstream_dop853(pot, C.byref(fr), rows.data_ptr(), t1.data_ptr(), n, float(tfinal),
                                      dt0, float(kw.get("atol", 1e-10)), float(kw.get("rtol", 1e-10)),
                                      int(kw.get("nmax", 0)), out.data_ptr(), status.data_ptr(), C.byref(opt))
        if rc not in (-1, -2, -3, -4):          # per-particle failures are reported through `status`
            _abi.check(rc)

    def run(self, prog_w0, prog_mass, nbody=None, release_every=1, n_particles=1, output_every=None,
            output_filename=None, check_filesize=True, overwrite=False, progress=False, Integrator=None,
            Integrator_kwargs=None, **time_spec):
        """``mockstream_generator.py:119-372``.  Returns ``(stream: MockStream, prog: PhaseSpacePosition)``."""
        Integrator_kwargs = dict(Integrator_kwargs or {})
        Integrator = get_integrator(Integrator or DOPRI853Integrator)
        units = self.hamiltonian.units
        t = parse_time_specification(units, **time_spec)
        if (nbody is None and not self.self_gravity and output_every is None and output_filename is None
                and Integrator in (LeapfrogIntegrator, DOPRI853Integrator) and isinstance(self.df, BaseStreamDF)
                and isinstance(self.hamiltonian.frame, StaticFrame) and len(t) > 1 and _abi.get_devices() is None
                and not os.environ.get("GB_NO_STREAM_PIPELINE")):
            rows, prow, release_time, lead_trail = self._run_pipelined(prog_w0, prog_mass, release_every, n_particles,
                                                                       Integrator, Integrator_kwargs, t)
            stream = MockStream(pos=rows[:, :3].T, vel=rows[:, 3:].T, release_time=release_time, lead_trail=lead_trail,
                                frame=self.hamiltonian.frame)
            return stream, PhaseSpacePosition(pos=prow[:, :3].T, vel=prow[:, 3:].T, frame=self.hamiltonian.frame)
        prog_nbody = self._get_nbody(prog_w0, nbody)
        nbody_orbits = prog_nbody.integrate_orbit(t=t, Integrator=Integrator, Integrator_kwargs=Integrator_kwargs)
        if t[1] < t[0]:
            # initial conditions are at the END time: flip, restart from the earliest state (:237-250)
            nbody_orbits = Orbit(pos=nbody_orbits.pos[:, ::-1], vel=nbody_orbits.vel[:, ::-1], t=nbody_orbits.t[::-1],
                                 hamiltonian=self.hamiltonian)
            nbody0 = DirectNBody(nbody_orbits[0], prog_nbody.particle_potentials,
                                 external_potential=self.hamiltonian.potential, frame=self.hamiltonian.frame,
                                 units=units)
        else:
            nbody0 = prog_nbody
        prog_orbit = Orbit(pos=nbody_orbits.pos[:, :, 0], vel=nbody_orbits.vel[:, :, 0], t=nbody_orbits.t,
                           hamiltonian=self.hamiltonian)
        orbit_t = np.asarray(prog_orbit.t, dtype=np.float64)
        stream_w0 = self.df.sample(prog_orbit, prog_mass, hamiltonian=self.hamiltonian,
                                   release_every=release_every, n_particles=n_particles)
        w0 = getattr(stream_w0, "_rows", None)          # the sampler's (Np, 6) array: no re-assembly
        if w0 is None:
            w0 = np.ascontiguousarray(np.vstack((stream_w0.pos, stream_w0.vel)).T)
        unq_t1s, nstream = np.unique(stream_w0.release_time, return_counts=True)
        all_nstream = _scatter_isclose(orbit_t, unq_t1s, nstream)
        nstream_idx = np.where(all_nstream != 0)[0]
        if 0 not in nstream_idx:
            nstream_idx = np.insert(nstream_idx, 0, 0)
            unq_t1s = np.insert(unq_t1s, 0, orbit_t[0])
        if output_every is not None and output_filename is None:
            raise ValueError("If output_every is specified, you must also pass in a filename to store the snapshots in")
        if Integrator is DOPRI853Integrator and output_every is not None:
            raw_nbody, raw_stream = mockstream_dop853_animate(nbody0, orbit_t, w0, all_nstream.astype("i4"),
                                                              output_every=output_every, output_filename=output_filename,
                                                              check_filesize=check_filesize, overwrite=overwrite,
                                                              progress=int(progress), **Integrator_kwargs)
        elif Integrator is DOPRI853Integrator:
            raw_nbody, raw_stream = mockstream_dop853(nbody0, orbit_t[nstream_idx], w0, unq_t1s, orbit_t[-1],
                                                      all_nstream[nstream_idx].astype("i4"), progress=int(progress),
                                                      **Integrator_kwargs)
        elif Integrator is LeapfrogIntegrator and output_every is not None:
            raise NotImplementedError("Animation output for LeapfrogIntegrator is not implemented")
        elif Integrator is LeapfrogIntegrator:
            raw_nbody, raw_stream = mockstream_leapfrog(nbody0, orbit_t, orbit_t[nstream_idx], w0, unq_t1s,
                                                        orbit_t[-1], all_nstream[nstream_idx].astype("i4"),
                                                        progress=int(progress))
        else:
            raise ValueError("Currently, only the DOPRI853Integrator and LeapfrogIntegrator are supported for "
                             "mock stream generation.")
        stream = MockStream(pos=raw_stream[:, :3].T, vel=raw_stream[:, 3:].T, release_time=stream_w0.release_time,
                            lead_trail=stream_w0.lead_trail, frame=self.hamiltonian.frame)
        prog = PhaseSpacePosition(pos=raw_nbody[:, :3].T, vel=raw_nbody[:, 3:].T, frame=self.hamiltonian.frame)
        return stream, prog
