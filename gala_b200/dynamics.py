"""Thin containers standing in for ``gala.dynamics.PhaseSpacePosition`` / ``Orbit``
(reference ``dynamics/core.py:57``, ``dynamics/orbit.py:22``) in the unit system of the
Hamiltonian: plain arrays, no astropy.  Only the touch points of the hot path exist:
``w()`` (``dynamics/core.py:476-525``), ``from_w``, slicing by orbit, and ``energy``."""
from __future__ import annotations

import numpy as np

__all__ = ["PhaseSpacePosition", "Orbit", "MockStream"]


def _is_torch(x):
    return type(x).__module__.startswith("torch")


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
