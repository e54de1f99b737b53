"""Reference frames (reference ``potential/frame/builtin/frames.py``): ``StaticFrame`` and the 3-D
``ConstantRotatingFrame``; ``spec()`` gives the ``gb_frame`` of the C ABI (mirror of CFrameType,
``potential/frame/src/cframe.h:7-17``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .units import galactic, strip

__all__ = ["StaticFrame", "ConstantRotatingFrame"]


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
