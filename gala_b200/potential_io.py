"""Potential specifications on disk: the YAML schema of ``gala.potential.potential.io`` (reference
``potential/potential/io.py:78-366``; fixtures ``tests/potential/potential/{Plummer,Composite,ccomposite,lm10}.yml``)
read and written without astropy -- the data format in front of the hot path.

schema::

    class: HernquistPotential            # one potential
    parameters: {m: 1.0e11, c: 2.0, c_unit: kpc}     # <name>_unit optional; a bare number is in the unit system
    units: {length: kpc, mass: solMass, time: Myr, angle: rad, ...}   # absent = dimensionless

    type: composite | custom             # composite: CCompositePotential / CompositePotential, built piecewise
    class: LM10Potential                 # custom: the class is called with one parameter dict per component name
    components: [{class: ..., name: ..., parameters: ..., units: ...}, ...]

Values with a ``_unit`` are converted into the unit system of the potential with a small table of the units gala
writes (lengths, times, masses, angles and their quotients / powers, e.g. ``km / s``, ``kpc / Myr``, ``mas / yr``).
Only the three unit systems this package knows (galactic, solarsystem, dimensionless) can be loaded."""
from __future__ import annotations

import io as _io
import math
import os
import re
from collections import OrderedDict

import numpy as np

from . import potential as _pot
from .units import dimensionless, galactic, solarsystem

__all__ = ["load", "save", "to_dict", "from_dict"]

# unit -> (value in [kpc, Myr, Msun, rad], which of the four it measures)
_KPC_PER_M = 1.0 / 3.0856775814913673e19
_MYR_PER_S = 1.0 / 3.15576e13
_UNITS = {
    "kpc": (1.0, 0), "pc": (1e-3, 0), "Mpc": (1e3, 0), "AU": (1.495978707e11 * _KPC_PER_M, 0), "km": (1e3 * _KPC_PER_M, 0),
    "m": (_KPC_PER_M, 0), "lyr": (9.4607304725808e15 * _KPC_PER_M, 0),
    "Myr": (1.0, 1), "Gyr": (1e3, 1), "kyr": (1e-3, 1), "yr": (1e-6, 1), "s": (_MYR_PER_S, 1), "d": (86400 * _MYR_PER_S, 1),
    "solMass": (1.0, 2), "Msun": (1.0, 2), "M_sun": (1.0, 2), "kg": (1.0 / 1.988409870698051e30, 2),
    "rad": (1.0, 3), "deg": (math.pi / 180, 3), "arcmin": (math.pi / 10800, 3), "arcsec": (math.pi / 648000, 3),
    "mas": (math.pi / 648000e3, 3),
}
_UNITS["au"] = _UNITS["AU"]
_SYSTEMS = (   # (UnitSystem, base units as gala prints them)
    (galactic, OrderedDict([("angle", "rad"), ("angular speed", "mas / yr"), ("length", "kpc"), ("mass", "solMass"),
                            ("speed", "km / s"), ("time", "Myr")])),
    (solarsystem, OrderedDict([("angle", "rad"), ("length", "AU"), ("mass", "solMass"), ("time", "yr")])),
)


_EXTRA_ARGS = {"MN3ExponentialDiskPotential": ("positive_density", "sech2_z"), "MultipolePotential": ("lmax",)}


def _parse_unit(text):
    """'kpc / Myr', 'km / s', 'kpc3 / (Myr2 solMass)', '' -> (scale in galactic base units, exponents [L, T, M, A])."""
    scale, dims = 1.0, [0, 0, 0, 0]
    text = (text or "").strip()
    if text in ("", "dimensionless"):
        return scale, dims
    parts = text.split("/")
    if len(parts) > 2:
        raise ValueError(f"cannot parse the unit '{text}'")
    for sign, part in zip((1, -1), parts):
        for tok in part.replace("(", " ").replace(")", " ").split():
            m = re.fullmatch(r"([A-Za-z_]+)\^?(-?\d+)?", tok)
            if m is None or m.group(1) not in _UNITS:
                raise ValueError(f"unknown unit '{tok}' in '{text}'")
            s, which = _UNITS[m.group(1)]
            p = sign * int(m.group(2) or 1)
            scale *= s ** p
            dims[which] += p
    return scale, dims


def _system_of(units_dict):
    """the UnitSystem named by a ``units:`` mapping (or list of unit strings); None -> dimensionless"""
    if not units_dict:
        return dimensionless, None
    names = list(units_dict.values()) if isinstance(units_dict, dict) else list(units_dict)
    base = [None, None, None, 1.0]
    for name in names:
        s, d = _parse_unit(str(name))
        if sum(abs(x) for x in d) == 1 and max(d) == 1:
            base[d.index(1)] = s
    for system, printed in _SYSTEMS:
        want = [_parse_unit(printed[k])[0] for k in ("length", "time", "mass")]
        if all(b is not None and abs(b / w - 1) < 1e-9 for b, w in zip(base[:3], want)):
            return system, base
    raise NotImplementedError(f"unit system {units_dict} is not one of galactic / solarsystem / dimensionless")


def _unpack_params(params, base):
    """io.py:14-34: numbers -> float / arrays, '<k>_unit' applied (converted into the potential's unit system)."""
    out = OrderedDict()
    for key, item in params.items():
        if key.endswith("_unit") and key[:-5] in params:
            continue
        if isinstance(item, str) or item is None:
            val = item
        elif np.iterable(item):
            val = np.array(item).astype(float)
        else:
            try:
                val = float(item)
            except Exception:
                val = item
        unit = params.get(key + "_unit")
        if unit not in (None, "") and not isinstance(val, str):
            scale, dims = _parse_unit(unit)
            if base is None:
                raise ValueError(f"parameter '{key}' carries the unit '{unit}' but the potential is dimensionless")
            val = val * (scale / math.prod(b ** p for b, p in zip(base, dims)))
        out[key] = val
    return out


def _cls(module, name):
    cls = getattr(module, name, None) if module is not None else None
    return cls if cls is not None else getattr(_pot, name)


def _parse_component(component, module):
    """io.py:37-75"""
    try:
        class_name = component["class"]
    except KeyError as e:
        raise KeyError("Potential dictionary must contain a key 'class' for specifying the name of the Potential "
                       "class.") from e
    system, base = _system_of(component.get("units"))
    params = _unpack_params(component.get("parameters", {}) or {}, base)
    return _cls(module, class_name)(units=system, **params, **(component.get("extra_args", {}) or {}))


def from_dict(d, module=None):
    """``gala.potential.potential.io.from_dict`` (io.py:78-164)."""
    kind = d.get("type")
    if kind == "composite":
        name = "CCompositePotential" if d["class"] == "CompositePotential" else d["class"]    # every component here is C-enabled
        p = _cls(module, name)()
        for i, component in enumerate(d["components"]):
            p[component.get("name", str(i))] = _parse_component(component, module)
        return p
    if kind == "custom":
        groups = OrderedDict()
        for component in d["components"]:
            if "name" not in component:
                raise KeyError("For custom potentials, component specification must include the component name (e.g., "
                               "name: 'blah')")
            _, base = _system_of(component.get("units"))
            groups[component["name"]] = dict(_unpack_params(component.get("parameters", {}) or {}, base))
        return _cls(module, d["class"])(**groups, **(d.get("extra_args", {}) or {}))
    return _parse_component(d, module)


def _plain(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, np.generic):
        return v.item()
    return v


def _to_dict_help(potential):
    """io.py:180-197; parameters are written as bare numbers in the potential's own unit system."""
    d = {"class": potential.__class__.__name__}
    for system, printed in _SYSTEMS:
        if potential.units is system:
            d["units"] = dict(printed)
    params = {k: _plain(v) for k, v in potential.parameters.items()}
    for k in ("Snlm", "Tnlm"):                      # SCF coefficient arrays are parameters in the reference
        if hasattr(potential, k):
            params[k] = _plain(getattr(potential, k))
    if params:
        d["parameters"] = params
    # what the constructor needs besides the parameters (the reference's `_extra_serialize_args`, io.py:192-195) --
    # plus origin / R, which the reference's writer drops but its reader accepts as constructor keywords
    extra = {k: _plain(getattr(potential, k)) for k in _EXTRA_ARGS.get(potential.__class__.__name__, ())}
    if np.any(np.asarray(potential.origin) != 0):
        extra["origin"] = _plain(np.asarray(potential.origin, dtype=float))
    if potential.R is not None:
        extra["R"] = _plain(np.asarray(potential.R, dtype=float))
    if extra:
        d["extra_args"] = extra
    return d


def to_dict(potential):
    """``gala.potential.potential.io.to_dict`` (io.py:200-269)."""
    if isinstance(potential, _pot.TimeInterpolatedPotential):
        raise NotImplementedError("TimeInterpolatedPotential has no YAML form (the reference cannot serialise the class "
                                  "argument either)")
    if isinstance(potential, _pot.CCompositePotential):
        d = {"class": potential.__class__.__name__, "components": []}
        for k, p in potential.items():
            comp = _to_dict_help(p)
            comp["name"] = k
            d["components"].append(comp)
        d["type"] = "composite" if potential.__class__ is _pot.CCompositePotential else "custom"
        if hasattr(potential, "version"):            # MilkyWayPotential._extra_serialize_args (builtin/special.py:186)
            d["extra_args"] = {"version": potential.version}
        return d
    return _to_dict_help(potential)


def _yaml():
    import yaml

    class Loader(yaml.SafeLoader):
        """SafeLoader + the one python tag gala's files contain (an OrderedDict of parameters)."""

    def ordered(loader, node):
        out = OrderedDict()
        for key_node, value_node in node.value:
            if loader.construct_object(key_node) == "dictitems":
                out.update(loader.construct_mapping(value_node, deep=True))
        return out
    Loader.add_constructor("tag:yaml.org,2002:python/object/apply:collections.OrderedDict", ordered)
    return yaml, Loader


def load(f, module=None):
    """``gala.potential.potential.io.load`` (io.py:272-322): a path, a block of YAML text, or a file-like object."""
    yaml, Loader = _yaml()
    if hasattr(f, "read"):
        text = f.read()
    elif isinstance(f, (str, os.PathLike)) and os.path.exists(os.fspath(f)):
        with open(os.path.abspath(os.fspath(f)), encoding="utf-8") as fil:
            text = fil.read()
    else:
        text = str(f)
    return from_dict(yaml.load(_io.StringIO(text), Loader=Loader), module=module)


def save(potential, f):
    """``gala.potential.potential.io.save`` (io.py:325-366): filename or file-like object."""
    yaml, _ = _yaml()
    d = to_dict(potential)
    if hasattr(f, "write"):
        yaml.safe_dump(d, f, default_flow_style=None)
    else:
        with open(f, "w", encoding="utf-8") as f2:
            yaml.safe_dump(d, f2, default_flow_style=None)
