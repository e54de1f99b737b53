"""CPU: potential specifications in the YAML schema of gala.potential.potential.io (reference
``potential/potential/io.py``; its tests ``tests/potential/potential/test_io.py:20-145``) -- loading files written in
that schema (fixtures under tests/golden/potentials/, incl. the python OrderedDict tag and `<name>_unit` values),
and save -> load round trips of every class."""
import io
import os

import numpy as np
import pytest

import gala_b200 as gb
from gala_b200 import io as gio

HERE = os.path.dirname(os.path.abspath(__file__))
KMS = gb.KMS_TO_KPC_MYR


def same(p, q):
    a, b = p._components(), q._components()
    return type(p) is type(q) and len(a) == len(b) and all(
        x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) and np.array_equal(x[3], y[3])
        for x, y in zip(a, b))


def test_load_files_in_the_reference_schema():
    p = gio.load(os.path.join(HERE, "golden", "potentials", "lm10_like.yml"))
    assert isinstance(p, gb.LM10Potential) and list(p.keys()) == ["disk", "bulge", "halo"]
    want = gb.LM10Potential(disk=dict(m=1e11, a=6.5, b=0.26), bulge=dict(m=3.4e10, c=0.7),
                            halo=dict(v_c=172.33345 * KMS, r_h=12.0, q1=1.38, q2=1.0, q3=1.36, phi=np.deg2rad(97.0)))
    for a, b in zip(p._components(), want._components()):
        assert a[0] == b[0] and np.allclose(a[1], b[1], rtol=1e-14)
    with open(os.path.join(HERE, "golden", "potentials", "composite_plain.yml")) as f:          # a file-like object
        c = gio.load(f)
    assert isinstance(c, gb.CCompositePotential) and list(c.keys()) == ["halo", "disk"]
    assert np.array_equal(c["disk"].c_parameters, [1e11, 6.5, 0.26]) and c["halo"].units is gb.galactic
    text = """
class: HernquistPotential
parameters:
  m: 1.0e11
  c: 2.0
"""
    h = gio.load(text)                                                                        # a block of YAML text
    assert isinstance(h, gb.HernquistPotential) and h.units is gb.dimensionless and h.G == 1.0
    k = gio.from_dict({"class": "KeplerPotential", "parameters": {"m": 2.0, "m_unit": "solMass"},
                       "units": {"length": "AU", "time": "yr", "mass": "solMass", "angle": "rad"}})
    assert k.units is gb.solarsystem and k.parameters["m"] == 2.0
    v = gio.from_dict({"class": "LogarithmicPotential", "units": ["kpc", "Myr", "solMass", "rad"],
                       "parameters": {"v_c": 0.2, "v_c_unit": "kpc / Myr", "r_h": 5000.0, "r_h_unit": "pc", "q1": 1.0,
                                      "q2": 0.9, "q3": 0.8, "phi": 0.0}})
    assert np.allclose([v.parameters["v_c"], v.parameters["r_h"]], [0.2, 5.0], rtol=1e-15)


def test_unit_table():
    s, d = gio._parse_unit("km / s")
    assert d == [1, -1, 0, 0] and np.isclose(s, KMS, rtol=1e-12)
    s, d = gio._parse_unit("kpc3 / (Myr2 solMass)")
    assert d == [3, -2, -1, 0] and s == 1.0
    s, d = gio._parse_unit("mas / yr")
    assert d == [0, -1, 0, 1] and np.isclose(s, np.pi / 648000e3 * 1e6)
    assert gio._parse_unit("") == (1.0, [0, 0, 0, 0])
    for bad in ("furlong", "kpc / Myr / s"):
        with pytest.raises(ValueError):
            gio._parse_unit(bad)


def test_errors():
    with pytest.raises(KeyError):
        gio.from_dict({"parameters": {"m": 1.0}})
    with pytest.raises(KeyError):
        gio.from_dict({"type": "custom", "class": "LM10Potential", "components": [{"class": "HernquistPotential"}]})
    with pytest.raises(NotImplementedError):
        gio.from_dict({"class": "HernquistPotential", "parameters": {"m": 1.0, "c": 1.0},
                       "units": {"length": "Mpc", "time": "Gyr", "mass": "solMass"}})
    with pytest.raises(ValueError):
        gio.from_dict({"class": "HernquistPotential", "parameters": {"m": 1.0, "m_unit": "solMass", "c": 1.0}})
    with pytest.raises(AttributeError):
        gio.from_dict({"class": "NoSuchPotential"})
    T = np.linspace(0, 1, 4)
    with pytest.raises(NotImplementedError):
        gio.to_dict(gb.TimeInterpolatedPotential(gb.KeplerPotential, T, m=1e10 * (1 + T)))


def _all_potentials():
    rng = np.random.default_rng(0)
    S, T = rng.normal(size=(3, 3, 3)), rng.normal(size=(3, 3, 3))
    Rm = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    pots = [
        gb.KeplerPotential(m=1e10), gb.HernquistPotential(m=1e11, c=2.0), gb.PlummerPotential(m=1e10, b=1.0),
        gb.IsochronePotential(m=1e10, b=0.5), gb.JaffePotential(m=1e10, c=2.0), gb.NFWPotential(m=6e11, r_s=16.0),
        gb.NFWPotential(m=6e11, r_s=16.0, a=1.0, b=0.9, c=0.8), gb.MiyamotoNagaiPotential(m=6e10, a=3.0, b=0.3),
        gb.MN3ExponentialDiskPotential(m=5e10, h_R=2.6, h_z=0.3),
        gb.MN3ExponentialDiskPotential(m=5e10, h_R=2.6, h_z=0.3, positive_density=False, sech2_z=False),
        gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=0.4), gb.StonePotential(m=1e10, r_c=0.5, r_h=10.0),
        gb.BurkertPotential(rho=1e7, r0=5.0), gb.SatohPotential(m=1e10, a=3.0, b=0.5), gb.KuzminPotential(m=1e10, a=3.0),
        gb.LogarithmicPotential(v_c=0.2, r_h=5.0, q1=1.0, q2=0.9, q3=0.8, phi=0.3),
        gb.LeeSutoTriaxialNFWPotential(v_c=0.2, r_s=15.0, a=1.0, b=0.9, c=0.8),
        gb.PowerLawCutoffPotential(m=1e10, alpha=1.8, r_c=1.9), gb.SCFPotential(m=1e11, r_s=10.0, Snlm=S, Tnlm=T),
        gb.MultipolePotential(lmax=2, m=1e10, r_s=8.0, inner=True, S00=1.0, S10=0.3, S22=0.2, T22=0.1),
        gb.KeplerPotential(m=1.0, units=gb.solarsystem), gb.PlummerPotential(m=1.0, b=0.5, units=gb.dimensionless),
        gb.HernquistPotential(m=1e11, c=2.0, origin=[1, 2, 3], R=Rm),
        gb.MilkyWayPotential(), gb.MilkyWayPotential(version="v2"), gb.MilkyWayPotential(disk=dict(m=7e10)),
        gb.MilkyWayPotential2022(), gb.LM10Potential(), gb.BovyMWPotential2014(),
    ]
    c = gb.CCompositePotential()
    c["a"], c["b"] = pots[1], pots[6]
    return pots + [c]


@pytest.mark.parametrize("k", range(30))
def test_save_load_round_trip(tmp_path, k):
    """tests/potential/potential/test_io.py:52-145: to a file object, to a filename, and back."""
    p = _all_potentials()[k]
    buf = io.StringIO()
    gio.save(p, buf)
    assert same(p, gio.load(buf.getvalue()))
    fn = tmp_path / "potential.yml"
    gio.save(p, str(fn))
    q = gio.load(fn)
    assert same(p, q) and q.units is p.units
    assert gio.to_dict(q) == gio.to_dict(p)
