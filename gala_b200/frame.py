"""Reference frames (reference ``potential/frame/builtin/frames.py``): ``StaticFrame`` and the 3-D
``ConstantRotatingFrame``; ``spec()`` gives the ``gb_frame`` of the C ABI (mirror of CFrameType,
``potential/frame/src/cframe.h:7-17``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .units import galactic, strip

__all__ = ["StaticFrame", "ConstantRotatingFrame", "static_to_constantrotating", "constantrotating_to_static",
           "static_to_static"]


class FrameBase:
    c_enabled = True

    def spec(self):
        raise NotImplementedError

    def __eq__(self, other):
        return type(self) is type(other) and np.array_equal(self.c_parameters, other.c_parameters)

    __hash__ = object.__hash__


class StaticFrame(FrameBase):
    def __init__(self, units=galactic):
        self.units = units
        self.c_parameters = np.array([])

    def spec(self):
        f = _abi.gb_frame()
        f.type_id = _abi.FRAME_STATIC
        return f

    def __repr__(self):
        return "<StaticFrame>"


class ConstantRotatingFrame(FrameBase):
    """Omega is the 3-vector pattern angular velocity in rad/Myr (frame.c_parameters,
    frame/cframe.pyx:98-112).  The 2-D variant of the reference is out of scope."""

    def __init__(self, Omega, units=galactic):
        Omega = np.atleast_1d(np.asarray(strip(Omega), dtype=np.float64))
        if Omega.shape != (3,):
            raise ValueError("Omega must be a 3-vector (the 2-D rotating frame is not supported)")
        self.units = units
        self.Omega = Omega
        self.c_parameters = Omega.copy()

    def spec(self):
        f = _abi.gb_frame()
        f.type_id = _abi.FRAME_ROTATING_3D
        for k in range(3):
            f.omega[k] = self.Omega[k]
        return f

    def __repr__(self):
        return f"<ConstantRotatingFrame Omega={self.Omega}>"


# -- frame transformations of positions / canonical momenta (potential/frame/builtin/transformations.py) -----------
def _axis_angle_rotate(x, k, theta):
    """Rodrigues rotation of x (3, m[, n]) about the unit vector k by theta (scalar or (m,)), numpy or torch
    (``rodrigues_axis_angle_rotate``, transformations.py:12-57): x cos + (k x x) sin + k (k.x)(1 - cos)."""
    tor = type(x).__module__.startswith("torch")
    if tor:
        import torch
        theta = torch.as_tensor(theta, dtype=x.dtype, device=x.device)
        c, s_ = torch.cos(theta), torch.sin(theta)
        stack = lambda rows: torch.stack(rows, dim=0)          # noqa: E731
    else:
        theta = np.asarray(theta, dtype=np.float64)
        c, s_ = np.cos(theta), np.sin(theta)
        stack = lambda rows: np.stack(rows, axis=0)             # noqa: E731
    if theta.ndim == 1:
        shape = (theta.shape[0],) + (1,) * (x.ndim - 2)
        c, s_ = c.reshape(shape), s_.reshape(shape)
    kx = (k[1] * x[2] - k[2] * x[1], k[2] * x[0] - k[0] * x[2], k[0] * x[1] - k[1] * x[0])
    kd = k[0] * x[0] + k[1] * x[1] + k[2] * x[2]
    return stack([c * x[i] + s_ * kx[i] + (1.0 - c) * kd * k[i] for i in range(3)])


def _rotating_static(frame_r, w, t, sign):
    """``_constantrotating_static_helper`` (transformations.py:98-150): positions AND velocities turn by the same
    angle -sign |Omega| t about Omega (the velocities are the canonical momenta, i.e. inertial velocities expressed
    in the rotating axes -- what the rotating-frame kernels integrate)."""
    if t is None:
        t = getattr(w, "t", None)
    if t is None:
        raise ValueError("Time must be supplied either through the input Orbit class instance or through the t "
                         "argument.")
    Om = -np.asarray(frame_r.Omega, dtype=np.float64)
    norm = float(np.linalg.norm(Om))
    if norm == 0.0:
        return w.pos, w.vel
    k = Om / norm
    theta = sign * norm * (t if type(t).__module__.startswith("torch") else np.asarray(strip(t), dtype=np.float64))
    return _axis_angle_rotate(w.pos, k, theta), _axis_angle_rotate(w.vel, k, theta)


def static_to_constantrotating(frame_i, frame_r, w, t=None):
    """transformations.py:153-170."""
    return _rotating_static(frame_r, w, t, 1.0)


def constantrotating_to_static(frame_r, frame_i, w, t=None):
    """transformations.py:173-190."""
    return _rotating_static(frame_r, w, t, -1.0)


def static_to_static(frame_r, frame_i, w, t=None):
    """transformations.py:193-220: no-op."""
    return w.pos, w.vel
