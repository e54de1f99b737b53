"""Drop-in hook for a real gala installation (``gala`` is NOT importable in this image: astropy is
absent; this module is exercised through its duck-typed extractors and, for the patching
itself, against a fake ``gala`` package tree with the reference's import structure,
tests/test_host_logic_cpu.py::test_plugin_install_patches_the_names_gala_resolves).

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
    if wrapper == "TimeInterpolatedWrapper":
        return _extract_time_interpolated(pot)
    if wrapper not in WRAPPER_TO_TYPE:
        raise TypeError(f"potential wrapper {wrapper} is not supported by the B200 engine")
    origin = getattr(pot, "origin", np.zeros(3))
    origin = getattr(origin, "value", origin)
    return _Extracted(WRAPPER_TO_TYPE[wrapper], pot.G, pot.c_parameters, origin, getattr(pot, "_R", None),
                      getattr(pot, "units", None))


def _value(x):
    return np.asarray(getattr(x, "value", x), dtype=np.float64)


def _extract_time_interpolated(pot):
    """gala ``TimeInterpolatedPotential`` -> ``gala_b200.TimeInterpolatedPotential``.  gala keeps the tables inside
    the Cython wrapper (cytimeinterp.pyx:71-130), out of reach from Python; what it builds them FROM is public:
    ``parameters['potential_cls' | 'time_knots' | 'interpolation_method' | <name>]``, ``_potential_param_names``,
    ``_interp_params``, ``_extra_wrapped_kwargs``, ``origin`` and ``R`` (time_interpolated.py:59-260).  The wrapped
    potential is instantiated by gala at every knot exactly as ``_setup_wrapper`` does for classes with derived C
    parameters (:267-285), and its ``c_parameters`` row is what the device interpolates."""
    P = pot.parameters
    wcls = P["potential_cls"]
    tk = _value(P["time_knots"])
    names = list(pot._potential_param_names)
    interp = set(pot._interp_params)
    extra = dict(getattr(pot, "_extra_wrapped_kwargs", {}) or {})
    rows, wtype = [], None
    for i in range(len(tk) if interp else 1):
        knot = wcls(units=pot.units, **{k: (P[k][i] if k in interp else P[k]) for k in names}, **extra)
        wname = type(knot.c_instance).__name__
        if wname not in WRAPPER_TO_TYPE:
            raise TypeError(f"potential wrapper {wname} inside a TimeInterpolatedPotential is not supported by the B200 engine")
        if wtype not in (None, WRAPPER_TO_TYPE[wname]):
            raise TypeError("the wrapped potential resolves to different C types at different time knots")
        wtype = WRAPPER_TO_TYPE[wname]
        rows.append(_value(knot.c_parameters))
    origin = getattr(pot, "origin", None)
    R = getattr(pot, "R", None)
    from .potential import TimeInterpolatedPotential
    return TimeInterpolatedPotential.from_tables(wtype, pot.G, tk, np.array(rows), None if origin is None else _value(origin),
                                                 None if R is None else _value(R), str(P["interpolation_method"]),
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


def _hook(installed, missing, modname, names, wrap):
    """setattr(module, name, wrapped) for every name in ``names`` of an importable module."""
    import importlib
    try:
        mod = importlib.import_module(modname)
    except ImportError as e:
        missing.append(f"{modname} ({e})")
        return
    for name, fn in names.items():
        setattr(mod, name, wrap(fn))
        installed.append(f"{modname}.{name}")


def install(strict=True):
    """Swap gala's Cython boundary functions for the GPU-backed ones.  Returns the list of hooks installed
    (``"module.function"``).  Raises ImportError if gala itself cannot be imported, and -- with ``strict`` --
    RuntimeError if a module the hot path resolves its functions through could not be patched.

    Where the reference LOOKS THE NAMES UP decides what has to be patched:
      * ``Hamiltonian.integrate_orbit`` does ``from ...integrate.cyintegrators import X`` at call time
        (potential/hamiltonian/chamiltonian.pyx:324-336): that resolves ``X`` as an attribute of the PACKAGE
        ``gala.integrate.cyintegrators``, whose ``__init__`` bound the Cython functions at import
        (integrate/cyintegrators/__init__.py:1-3).  So the package attribute is the one that counts; the
        submodules ``.leapfrog`` / ``.ruth4`` / ``.dop853`` are patched too for code that imports from them.
      * ``MockStreamGenerator.run`` uses the names ``mockstream_generator`` imported from ``._mockstream`` at
        import time (dynamics/mockstream/mockstream_generator.py:8-12; the extension module is
        ``gala.dynamics.mockstream._mockstream``, setup.py:238-255), so both that module and the generator's own
        namespace are patched, and ``gala.dynamics.mockstream`` re-exports ``mockstream_dop853`` (``__init__.py:1``)."""
    import gala  # noqa: F401  (ImportError here = no gala to patch)
    installed, missing = [], []
    integ = {"leapfrog_integrate_hamiltonian": _integ.leapfrog_integrate_hamiltonian,
             "ruth4_integrate_hamiltonian": _integ.ruth4_integrate_hamiltonian,
             "dop853_integrate_hamiltonian": _integ.dop853_integrate_hamiltonian}
    _hook(installed, missing, "gala.integrate.cyintegrators", integ, _wrap)
    for sub, name in (("leapfrog", "leapfrog_integrate_hamiltonian"), ("ruth4", "ruth4_integrate_hamiltonian"),
                      ("dop853", "dop853_integrate_hamiltonian")):
        _hook(installed, missing, f"gala.integrate.cyintegrators.{sub}", {name: integ[name]}, _wrap)
    stream = {n: getattr(_ms, n) for n in ("mockstream_dop853", "mockstream_leapfrog", "mockstream_dop853_animate")}
    _hook(installed, missing, "gala.dynamics.mockstream._mockstream", stream, _wrap_stream)
    _hook(installed, missing, "gala.dynamics.mockstream.mockstream_generator", stream, _wrap_stream)
    _hook(installed, missing, "gala.dynamics.mockstream", {"mockstream_dop853": stream["mockstream_dop853"]}, _wrap_stream)
    required = ("gala.integrate.cyintegrators.leapfrog_integrate_hamiltonian",
                "gala.dynamics.mockstream.mockstream_generator.mockstream_dop853")
    lacking = [r for r in required if r not in installed]
    if missing:
        import warnings
        warnings.warn("gala_b200.gala_plugin.install(): not patched: " + "; ".join(missing), RuntimeWarning)
    if strict and lacking:
        raise RuntimeError("gala_b200.gala_plugin.install(): the hot path would still run gala's own code: "
                           + ", ".join(lacking) + " could not be patched")
    return installed
