"""Drop-in hook for a real gala installation (``gala`` is NOT importable in this image: astropy is
absent; this module is exercised only through its duck-typed extractors).

``install()`` replaces the five Cython functions at the boundary with GPU-backed ones::

    import gala_b200.gala_plugin as plug; plug.install()
    # from here on Hamiltonian.integrate_orbit / MockStreamGenerator.run use the B200 engine

The extractor reads exactly what the reference's own Cython code reads from the Python objects
(SURVEY.md section 8b): ``type(pot.c_instance).__name__`` (wrapper class), ``pot.G``,
``pot.c_parameters``, ``pot.origin``, ``pot._R``; composites iterate ``pot.values()`` in insertion
order (``ccompositepotential.pyx:78-86``); frames give ``frame.c_parameters`` = Omega.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .frame import ConstantRotatingFrame, StaticFrame
from .hamiltonian import Hamiltonian
from .potential import PotentialBase, CCompositePotential
from . import integrate as _integ
from . import mockstream as _ms

WRAPPER_TO_TYPE = {           # potential/potential/builtin/cybuiltin.pyx:99-122, scf/bfe_class.pyx:38
    "NullWrapper": _abi.POT_NULL, "HernquistWrapper": _abi.POT_HERNQUIST,
    "SphericalNFWWrapper": _abi.POT_NFW_SPHERICAL, "FlattenedNFWWrapper": _abi.POT_NFW_FLATTENED,
    "TriaxialNFWWrapper": _abi.POT_NFW_TRIAXIAL, "MiyamotoNagaiWrapper": _abi.POT_MIYAMOTONAGAI,
    "MN3ExponentialDiskWrapper": _abi.POT_MN3, "LongMuraliBarWrapper": _abi.POT_LONGMURALIBAR,
    "SCFWrapper": _abi.POT_SCF, "KeplerWrapper": _abi.POT_KEPLER, "PlummerWrapper": _abi.POT_PLUMMER,
    "IsochroneWrapper": _abi.POT_ISOCHRONE, "JaffeWrapper": _abi.POT_JAFFE,
    "MultipoleWrapper": _abi.POT_MULTIPOLE, "StoneWrapper": _abi.POT_STONE, "BurkertWrapper": _abi.POT_BURKERT,
    "SatohWrapper": _abi.POT_SATOH, "KuzminWrapper": _abi.POT_KUZMIN, "LogarithmicWrapper": _abi.POT_LOGARITHMIC,
    "LeeSutoTriaxialNFWWrapper": _abi.POT_LEESUTO, "PowerLawCutoffWrapper": _abi.POT_POWERLAWCUTOFF,
}


class _Extracted(PotentialBase):
    """A gala potential seen through the C-ABI: same parameter vector, no re-derivation."""

    def __init__(self, type_id, G, c_parameters, origin, R, units=None):
        self._type_id = type_id
        self.G = float(G)
        self.units = units
        self.c_parameters = np.asarray(c_parameters, dtype=np.float64)
        self.origin = np.asarray(origin, dtype=np.float64)
        self.R = None if R is None else np.asarray(R, dtype=np.float64)
        self.parameters = {}
        self._spec = None
        self.strict_math = False


def extract_potential(pot):
    """gala potential object (duck-typed) -> gala_b200 potential carrying the identical C vector."""
    if isinstance(pot, PotentialBase):
        return pot
    wrapper = type(pot.c_instance).__name__
    if wrapper == "CCompositePotentialWrapper":
        out = CCompositePotential()
        for k, v in pot.items():
            out[k] = extract_potential(v)
        return out
    if wrapper not in WRAPPER_TO_TYPE:
        raise TypeError(f"potential wrapper {wrapper} is not supported by the B200 engine")
    origin = getattr(pot, "origin", np.zeros(3))
    origin = getattr(origin, "value", origin)
    return _Extracted(WRAPPER_TO_TYPE[wrapper], pot.G, pot.c_parameters, origin, getattr(pot, "_R", None),
                      getattr(pot, "units", None))


def extract_frame(frame):
    if isinstance(frame, (StaticFrame, ConstantRotatingFrame)):
        return frame
    name = type(frame.c_instance).__name__ if hasattr(frame, "c_instance") else type(frame).__name__
    if name.startswith("StaticFrame"):
        return StaticFrame()
    if name.startswith("ConstantRotatingFrameWrapper3D") or name == "ConstantRotatingFrame":
        return ConstantRotatingFrame(np.asarray(frame.c_parameters, dtype=np.float64))
    raise TypeError(f"frame {name} is not supported by the B200 engine")


def extract_hamiltonian(H):
    if isinstance(H, Hamiltonian):
        return H
    return Hamiltonian(extract_potential(H.potential), extract_frame(H.frame))


def _wrap(fn):
    def inner(hamiltonian, w0, t, *a, **kw):
        tt, w = fn(extract_hamiltonian(hamiltonian), np.asarray(w0), np.asarray(t), *a, **kw)[:2]
        return np.asarray(tt), np.asarray(w)
    inner.__name__ = fn.__name__
    inner.__doc__ = fn.__doc__
    return inner


def adapt_nbody(nbody):
    """gala ``DirectNBody`` (duck-typed: ``_c_w0`` (nbodies,6), ``particle_potentials``, ``H``) -> ours."""
    if isinstance(nbody, _ms.DirectNBody):
        return nbody
    H = extract_hamiltonian(nbody.H)
    pps = []
    for pp in nbody.particle_potentials:
        if pp is None or type(getattr(pp, "c_instance", None)).__name__ == "NullWrapper":
            pps.append(None)
        else:
            pps.append(extract_potential(pp))
    return _ms.DirectNBody(np.ascontiguousarray(np.asarray(nbody._c_w0, dtype=np.float64).T), pps,
                           external_potential=H.potential, frame=H.frame)


def _wrap_stream(fn):
    def inner(nbody, *a, **kw):
        return fn(adapt_nbody(nbody), *[np.asarray(x) if hasattr(x, "shape") else x for x in a], **kw)
    inner.__name__ = fn.__name__
    inner.__doc__ = fn.__doc__
    return inner


def install():
    """Swap gala's Cython boundary functions for the GPU-backed ones.  Raises ImportError if gala
    itself cannot be imported."""
    import gala.integrate.cyintegrators.leapfrog as lf
    import gala.integrate.cyintegrators.ruth4 as r4
    import gala.integrate.cyintegrators.dop853 as d8
    import gala.potential.hamiltonian.chamiltonian as ch
    lf.leapfrog_integrate_hamiltonian = _wrap(_integ.leapfrog_integrate_hamiltonian)
    r4.ruth4_integrate_hamiltonian = _wrap(_integ.ruth4_integrate_hamiltonian)
    d8.dop853_integrate_hamiltonian = _wrap(_integ.dop853_integrate_hamiltonian)
    for name in ("leapfrog_integrate_hamiltonian", "ruth4_integrate_hamiltonian", "dop853_integrate_hamiltonian"):
        if hasattr(ch, name):       # chamiltonian.pyx imports them lazily inside integrate_orbit
            setattr(ch, name, _wrap(getattr(_integ, name)))
    try:        # mock-stream boundary (dynamics/mockstream/mockstream.pyx:176-620)
        import gala.dynamics.mockstream.mockstream as msx
        import gala.dynamics.mockstream.mockstream_generator as msg
        for name in ("mockstream_dop853", "mockstream_leapfrog", "mockstream_dop853_animate"):
            w = _wrap_stream(getattr(_ms, name))
            setattr(msx, name, w)
            if hasattr(msg, name):
                setattr(msg, name, w)
    except ImportError:
        pass
    return True
