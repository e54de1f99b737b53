"""Thin containers standing in for ``gala.dynamics.PhaseSpacePosition`` / ``Orbit``
(reference ``dynamics/core.py:57``, ``dynamics/orbit.py:22``) in the unit system of the
Hamiltonian: plain arrays, no astropy.  Only the touch points of the hot path exist:
``w()`` (``dynamics/core.py:476-525``), ``from_w``, slicing by orbit, and ``energy``."""
from __future__ import annotations

import numpy as np

__all__ = ["PhaseSpacePosition", "Orbit", "MockStream", "peak_to_peak_period", "estimate_dt_n_steps",
           "combine"]


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def peak_to_peak_period(t, f, amplitude_threshold=1e-2):
    """``gala.dynamics.util.peak_to_peak_period`` (``dynamics/util.py:15-69``): the mean of the mean spacing of the
    interior local maxima and of the interior local minima of f(t) (wrap-around comparison, end points dropped);
    NaN when the oscillation is smaller than ``amplitude_threshold``."""
    from scipy.signal import argrelmax, argrelmin
    t = np.asarray(t, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    last = len(f) - 1
    max_ix = argrelmax(f, mode="wrap")[0]
    max_ix = max_ix[(max_ix != 0) & (max_ix != last)]
    min_ix = argrelmin(f, mode="wrap")[0]
    min_ix = min_ix[(min_ix != 0) & (min_ix != last)]
    if len(max_ix) < 2 or len(min_ix) < 2:       # no interior peak pair or trough pair: the reference's means are NaN
        return np.nan
    if abs(np.mean(f[max_ix]) - np.mean(f[min_ix])) < amplitude_threshold:
        return np.nan
    return np.mean([np.mean(np.diff(t[max_ix])), np.mean(np.diff(t[min_ix]))])


class PhaseSpacePosition:
    def __init__(self, pos, vel, frame=None):
        self.pos = pos if _is_torch(pos) else np.asarray(pos, dtype=np.float64)
        self.vel = vel if _is_torch(vel) else np.asarray(vel, dtype=np.float64)
        if tuple(self.pos.shape) != tuple(self.vel.shape):
            raise ValueError("pos and vel must have the same shape")
        self.frame = frame

    @property
    def ndim(self):
        return self.pos.shape[0]

    @property
    def shape(self):
        return tuple(self.pos.shape[1:])

    xyz = property(lambda self: self.pos)
    v_xyz = property(lambda self: self.vel)

    def w(self, units=None):
        """(2*ndim, ...) array [pos; vel]."""
        if _is_torch(self.pos):
            import torch
            return torch.cat([self.pos, self.vel], dim=0)
        return np.concatenate([self.pos, self.vel], axis=0)

    @classmethod
    def from_w(cls, w, units=None, **kw):
        n = w.shape[0] // 2
        return cls(pos=w[:n], vel=w[n:], **kw)

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        return self.__class__(pos=self.pos[(slice(None),) + key], vel=self.vel[(slice(None),) + key],
                              frame=self.frame)

    def to_frame(self, frame, current_frame=None, **kwargs):
        """``PhaseSpacePosition.to_frame`` (``dynamics/core.py:380-429``): static <-> constant-rotating; needs ``t=``."""
        from . import frame as frame_trans
        if self.frame is None and current_frame is None:
            raise ValueError(f"If no frame was specified when this {self} was initialized, you must pass the current "
                             "frame in via the current_frame argument to transform to a new frame.")
        if current_frame is None:
            current_frame = self.frame
        name1 = current_frame.__class__.__name__.rstrip("Frame").lower()
        name2 = frame.__class__.__name__.rstrip("Frame").lower()
        func = getattr(frame_trans, f"{name1}_to_{name2}", None)
        if func is None:
            raise ValueError(f"Unsupported frame transformation: {current_frame} to {frame}")
        pos, vel = func(current_frame, frame, self, **kwargs)
        return PhaseSpacePosition(pos=pos, vel=vel, frame=frame)

    # -- the per-point quantities of ``dynamics/core.py:652-740`` (plain array arithmetic: numpy or torch) --------
    def kinetic_energy(self):
        """0.5 |v|^2 per unit mass (``dynamics/core.py:652-665``)."""
        return 0.5 * (self.vel * self.vel).sum(0)

    def potential_energy(self, potential, t=0.0):
        """Phi(pos) per unit mass, evaluated on the device (``dynamics/core.py:667-686``)."""
        flat = self.pos.reshape(self.pos.shape[0], -1)
        return potential.energy(flat, t).reshape(self.shape)

    def guiding_radius(self, potential, t=0.0, xtol=1e-12, maxiter=60):
        """``PhaseSpacePosition.guiding_radius`` (``dynamics/core.py:742-764``, ``_guiding_radius_helper`` ``:951-967``):
        the cylindrical radius R_g at which a circular orbit in the z = 0 plane has this point's |L_z|, i.e. the root of
        |L_z| - R v_circ(R).  The reference runs one scipy ``root`` (hybr, xtol 1e-5) per point, each iteration a
        single-point ``circular_velocity`` call; here ALL points take secant steps together, one batched device gradient
        per iteration, started like the reference from the point's own R.  Points that do not converge give NaN."""
        tor = _is_torch(self.pos)
        flat_p = self.pos.reshape(3, -1)
        Lz = abs(self.angular_momentum()[2].reshape(-1))
        R0 = (flat_p[0] * flat_p[0] + flat_p[1] * flat_p[1]) ** 0.5
        if tor:
            import torch
            zeros, where, isfinite = torch.zeros_like, torch.where, torch.isfinite
        else:
            zeros, where, isfinite = np.zeros_like, np.where, np.isfinite

        def resid(R):
            q = zeros(flat_p)
            q[0] = R
            return Lz - R * potential.circular_velocity(q, t)

        Ra, Rb = R0, R0 * 1.05
        fa, fb = resid(Ra), resid(Rb)
        done = zeros(R0) != 0
        for _ in range(maxiter):
            denom = fb - fa
            ok = denom != 0
            Rn = Rb - where(ok, fb * (Rb - Ra) / where(ok, denom, denom + 1), zeros(R0))
            Rn = where(Rn > 0, Rn, 0.5 * Rb)                      # the root is a radius: never step across R = 0
            Rn = where(done, Rb, Rn)                              # converged points stay where they are
            done = done | (abs(Rn - Rb) <= xtol * abs(Rn))
            Ra, fa, Rb = Rb, fb, Rn
            if bool(done.all()):
                break
            fb = resid(Rb)
        out = where(done & isfinite(Rb), Rb, Rb * float("nan"))
        return out.reshape(self.shape)

    def angular_momentum(self):
        """q x p per unit mass, shape (3, ...) (``dynamics/core.py:708-740``)."""
        x, y, z = self.pos
        vx, vy, vz = self.vel
        L = [y * vz - z * vy, z * vx - x * vz, x * vy - y * vx]
        if _is_torch(self.pos):
            import torch
            return torch.stack(L, dim=0)
        return np.stack(L, axis=0)


class Orbit(PhaseSpacePosition):
    """pos, vel of shape (3, ntimes[, norbits]) plus the time grid."""

    def __init__(self, pos, vel, t=None, hamiltonian=None, frame=None):
        super().__init__(pos, vel, frame=frame if frame is not None else getattr(hamiltonian, "frame", None))
        self.t = t
        self.hamiltonian = hamiltonian

    @property
    def ntimes(self):
        return self.pos.shape[1]

    @property
    def norbits(self):
        return 1 if self.pos.ndim < 3 else self.pos.shape[2]

    @classmethod
    def from_w(cls, w, units=None, t=None, hamiltonian=None, **kw):
        n = w.shape[0] // 2
        return cls(pos=w[:n], vel=w[n:], t=t, hamiltonian=hamiltonian, **kw)

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        t = self.t
        if t is not None and len(key) >= 1 and not isinstance(key[0], (int, np.integer)):
            t = t[key[0]]
        elif t is not None and len(key) >= 1:
            return PhaseSpacePosition(pos=self.pos[(slice(None),) + key], vel=self.vel[(slice(None),) + key],
                                      frame=self.frame)
        return Orbit(pos=self.pos[(slice(None),) + key], vel=self.vel[(slice(None),) + key], t=t,
                     hamiltonian=self.hamiltonian, frame=self.frame)

    # -- pericentre / apocentre / zmax / eccentricity (dynamics/orbit.py:391-681), reduced on the GPU ----------------
    def _extrema(self, kind, func, return_times, approximate):
        from .integrate import orbit_extrema, orbit_extrema_list
        if approximate:
            raise NotImplementedError("approximate=True (no parabola refinement) is not part of the device reduction")
        if self.t is None:
            raise ValueError("pericenter / apocenter / zmax need the orbit's time grid")
        if return_times and func is not None:
            raise ValueError("Cannot return times if reducing using an input function. Pass `func=None` if you want to "
                             "return all individual values and times.")
        w = self.w()
        single = w.ndim == 2
        w3 = w[:, :, None] if single else w
        if func is None:             # every extremum (and its time): one array per orbit
            vals, times = orbit_extrema_list(w3, self.t, kind)
            if single:
                return (vals[0], times[0]) if return_times else vals[0]
            return (vals, times) if return_times else vals
        row = {np.mean: "mean", np.min: "min", np.max: "max", np.amin: "min", np.amax: "max"}.get(func)
        if row is None:              # any other reduction: applied to the list, like the reference applies func
            vals, _ = orbit_extrema_list(w3, self.t, kind)
            out = np.array([func(v) for v in vals])
            return out[0] if single else out
        if self.hamiltonian is None:
            raise ValueError("the device reduction needs the orbit's hamiltonian")
        st = orbit_extrema(self.hamiltonian, w3, self.t)
        out = st[f"{kind}_{row}"]
        return out[0] if single else out

    def pericenter(self, return_times=False, func=np.mean, approximate=False):
        """``Orbit.pericenter`` (``dynamics/orbit.py:439-493``): mean (default) / min / max of the refined local minima of
        r(t) straight from the device reduction; ``func=None`` all of them (with ``return_times`` also their times);
        any other ``func`` is applied to the list."""
        return self._extrema("peri", func, return_times, approximate)

    def apocenter(self, return_times=False, func=np.mean, approximate=False):
        """``Orbit.apocenter`` (``dynamics/orbit.py:495-553``)."""
        return self._extrema("apo", func, return_times, approximate)

    def zmax(self, return_times=False, func=np.mean, approximate=False):
        """``Orbit.zmax`` (``dynamics/orbit.py:600-656``): refined local maxima of |z|."""
        return self._extrema("zmax", func, return_times, approximate)

    def eccentricity(self, **kw):
        """``Orbit.eccentricity`` (``dynamics/orbit.py:658-681``): (r_apo - r_per) / (r_apo + r_per) of the means."""
        ra, rp = self.apocenter(**kw), self.pericenter(**kw)
        return (ra - rp) / (ra + rp)

    def potential_energy(self, potential=None, t=0.0):
        """``Orbit.potential_energy`` (``dynamics/orbit.py:339-359``): the orbit's own potential unless one is given."""
        if potential is None:
            if self.hamiltonian is None:
                raise ValueError("To compute the potential energy, a potential object must be provided!")
            potential = self.hamiltonian.potential
        return super().potential_energy(potential, t)

    def to_frame(self, frame, current_frame=None, **kwargs):
        """``Orbit.to_frame`` (``dynamics/orbit.py:1256-1296``): the orbit's own time grid is the default ``t``."""
        if current_frame is None:
            current_frame = self.frame
        if current_frame is not None and frame == current_frame and not kwargs:
            return self
        kw = dict(kwargs)
        kw.setdefault("t", self.t)
        psp = PhaseSpacePosition.to_frame(self, frame, current_frame, **kw)
        return Orbit(pos=psp.pos, vel=psp.vel, t=self.t, hamiltonian=self.hamiltonian, frame=frame)

    # -- circulation / period estimates (dynamics/orbit.py:683-871, dynamics/util.py:15-69) --------------------------
    def circulation(self):
        """``Orbit.circulation`` (``dynamics/orbit.py:736-792``): 1 for every axis about which the angular momentum
        never changes sign (nor drops below 1e-13) after the first sample; shape (3,) or (3, norbits).  A reduction
        over the time axis -- a device-resident trajectory is reduced where it is."""
        L = self.angular_momentum()
        single = L.ndim == 2
        if single:
            L = L[..., None]
        if _is_torch(L):
            import torch
            flip = (torch.sign(L[:, :1]) != torch.sign(L[:, 1:])) | (L[:, 1:].abs() < 1e-13)
            circ = (~flip.any(dim=1)).to(torch.int64).cpu().numpy()
        else:
            flip = (np.sign(L[:, :1]) != np.sign(L[:, 1:])) | (np.abs(L[:, 1:]) < 1e-13)
            circ = (~flip.any(axis=1)).astype(int)
        return circ.reshape(3) if single else circ

    def align_circulation_with_z(self, circulation=None):
        """``Orbit.align_circulation_with_z`` (``dynamics/orbit.py:794-871``): tube orbits about x or y get that
        axis exchanged with z (positions and velocities); z-tubes and boxes are returned unchanged."""
        circ = self.circulation() if circulation is None else np.asarray(circulation)
        circ = circ.reshape(3, -1)
        single = self.pos.ndim == 2
        pos = self.pos[..., None] if single else self.pos
        vel = self.vel[..., None] if single else self.vel
        if circ.shape[1] != pos.shape[2]:
            raise ValueError("Shape of 'circulation' array should match the shape of the position/velocity (minus "
                             "the time axis).")
        new_pos, new_vel = (a.clone() if _is_torch(a) else a.copy() for a in (pos, vel))
        for n in range(pos.shape[2]):
            if circ[2, n] == 1 or not circ[:, n].any():
                continue
            if circ[:, n].sum() > 1:
                import warnings
                warnings.warn("Circulation about multiple axes - are you sure the orbit has been integrated for long "
                              "enough?")
            ax = 0 if circ[0, n] == 1 else 1
            for new, old in ((new_pos, pos), (new_vel, vel)):
                new[ax, :, n] = old[2, :, n]
                new[2, :, n] = old[ax, :, n]
        return Orbit(pos=new_pos.reshape(self.pos.shape), vel=new_vel.reshape(self.vel.shape), t=self.t,
                     hamiltonian=self.hamiltonian, frame=self.frame)

    def estimate_period(self, components=("x", "y", "z")):
        """``Orbit.estimate_period`` (``dynamics/orbit.py:683-730``): mean peak-to-peak / trough-to-trough spacing of
        every requested component (``peak_to_peak_period``, ``dynamics/util.py:15-69``) as a dict of (norbits,)
        arrays.  Besides x, y, z the cylindrical ``rho`` and ``phi`` of ``orbit.cylindrical.estimate_period()``
        and the spherical ``r`` of ``orbit.physicsspherical`` can be asked for by name.  Host-side analysis (scipy ``argrelmax``), like the reference."""
        if self.t is None:
            raise ValueError("To compute the period, a time array is needed. Specify a time array when creating this "
                             "object.")
        pos = self.pos.cpu().numpy() if _is_torch(self.pos) else self.pos
        t = np.asarray(self.t.cpu().numpy() if _is_torch(self.t) else self.t, dtype=np.float64)
        pos = pos.reshape(3, pos.shape[1], -1)
        series = {"x": pos[0], "y": pos[1], "z": pos[2], "rho": np.hypot(pos[0], pos[1]),
                  "phi": np.arctan2(pos[1], pos[0]), "r": np.sqrt((pos * pos).sum(0))}
        return {k: np.array([peak_to_peak_period(t, series[k][:, n]) for n in range(pos.shape[2])]) for k in components}

    def energy(self, hamiltonian=None):
        """Hamiltonian value along the orbit, shape (ntimes[, norbits]) -- evaluated on the GPU."""
        H = hamiltonian or self.hamiltonian
        if H is None:
            raise ValueError("an Orbit without a hamiltonian needs one passed in")
        w = self.w()
        flat = w.reshape(w.shape[0], -1)
        return H.energy(flat).reshape(tuple(w.shape[1:]))


class MockStream(PhaseSpacePosition):
    """reference ``dynamics/mockstream/core.py`` MockStream: adds release_time and lead_trail."""

    def __init__(self, pos, vel, release_time=None, lead_trail=None, frame=None):
        super().__init__(pos, vel, frame=frame)
        self.release_time = release_time
        self.lead_trail = lead_trail


def _autodetermine_initial_dt(w0, H, dE_threshold=1e-9, **integrate_kwargs):
    """``dynamics/util.py:72-93``: the largest dt of logspace(1, -3, 8) whose 1000-time-unit test orbit keeps
    |E_last - E_first| / |E_first| below the threshold (the smallest one if none does)."""
    if w0.shape and w0.shape[0] > 1:
        raise ValueError("Only one set of initial conditions may be passed in at a time.")
    if dE_threshold is None:
        return 1.0
    for dt in np.logspace(-3, 1, 8)[::-1]:
        orbit = H.integrate_orbit(w0, dt=dt, n_steps=round(1000 / dt), **integrate_kwargs)
        E = orbit.energy()
        if abs(float((E[-1] - E[0]) / E[0])) < dE_threshold:
            break
    return float(dt)


def estimate_dt_n_steps(w0, hamiltonian, n_periods, n_steps_per_period, dE_threshold=1e-9, func=np.nanmax,
                        **integrate_kwargs):
    """``gala.dynamics.util.estimate_dt_n_steps`` (``dynamics/util.py:96-207``): a test orbit (10000 time units at
    the dt of ``_autodetermine_initial_dt``) gives the periods -- cylindrical ones after aligning a tube orbit with
    z, Cartesian ones for a box -- ``func`` picks one, and (dt, n_steps) sample ``n_periods`` of it with
    ``n_steps_per_period`` steps each.  Every integration runs on the device; the period analysis is host-side."""
    from .hamiltonian import Hamiltonian
    if not isinstance(w0, PhaseSpacePosition):
        w0 = PhaseSpacePosition.from_w(np.asarray(w0, dtype=np.float64))
    H = hamiltonian if isinstance(hamiltonian, Hamiltonian) else Hamiltonian(hamiltonian)
    dt = _autodetermine_initial_dt(w0, H, dE_threshold=dE_threshold, **integrate_kwargs)
    orbit = H.integrate_orbit(w0, dt=dt, n_steps=round(10000 / dt), **integrate_kwargs)
    circ = orbit.circulation()
    if np.any(circ):
        orbit = orbit.align_circulation_with_z(circulation=circ)
        names = ("rho", "phi", "z")
    else:
        names = ("x", "y", "z")
    per = orbit.estimate_period(components=names)
    T = func(np.array([per[k][0] for k in names]))
    if np.isnan(T):
        raise RuntimeError("Failed to find period.")
    dt = float(T) / float(n_steps_per_period)
    n_steps = round(n_periods * float(T) / dt)
    if dt == 0.0 or dt < 1e-13:
        raise ValueError("Timestep is zero or very small!")
    return dt, n_steps


def combine(objs):
    """``gala.dynamics.combine`` (``dynamics/util.py:209-350``): several PhaseSpacePosition objects -> one with the
    points side by side, several Orbit objects on the same time grid -> one with the orbits side by side (the way a
    batch of initial conditions or of results is assembled for one device call).  Same type, ndim, frame and -- for
    orbits -- hamiltonian and time array are required; numpy or torch arrays."""
    if isinstance(objs, PhaseSpacePosition) or not np.iterable(objs) or len(objs) < 1:
        raise ValueError("You must pass a non-empty iterable to combine.")
    if len(objs) == 1:
        return objs[0]
    first = objs[0]
    if first.__class__ not in (PhaseSpacePosition, Orbit):
        raise TypeError("Objects must be either PhaseSpacePosition or Orbit instances.")
    for obj in objs:
        if obj.__class__ != first.__class__:
            raise TypeError("All objects must have the same type.")
        if obj.ndim != first.ndim:
            raise ValueError("All objects must have the same ndim.")
        if obj.frame != first.frame:
            raise ValueError("All objects must have the same frame.")
        if isinstance(obj, Orbit):
            if obj.hamiltonian is not first.hamiltonian:
                raise ValueError("All objects must have the same potential.")
            if obj.t is not None and first.t is not None and not (
                    len(obj.t) == len(first.t) and bool(abs(obj.t - first.t).max() <= 1e-13)):
                raise ValueError("All orbits must have the same time array.")
    axis = 1 if first.__class__ is PhaseSpacePosition else 2

    def cat(arrs):
        arrs = [a[..., None] if a.ndim == axis else a for a in arrs]
        if _is_torch(arrs[0]):
            import torch
            return torch.cat(arrs, dim=axis)
        return np.concatenate(arrs, axis=axis)

    pos, vel = cat([o.pos for o in objs]), cat([o.vel for o in objs])
    if first.__class__ is PhaseSpacePosition:
        return PhaseSpacePosition(pos=pos, vel=vel, frame=first.frame)
    return Orbit(pos=pos, vel=vel, t=first.t, hamiltonian=first.hamiltonian, frame=first.frame)
