"""Minimal unit handling.  The engine works in plain float64 numbers in gala's ``galactic`` unit
system (kpc, Myr, Msun, rad; reference ``units.py:379``); astropy is not required.  When a value
carrying astropy-like units arrives (has ``.decompose``/``.to_value``), it is converted by the
caller-facing shims via ``strip``."""
from __future__ import annotations

# G in kpc^3 Msun^-1 Myr^-2: astropy.constants.G.decompose(galactic), the value the reference's
# PotentialBase stores as self.G (potential/potential/core.py:127-131); SURVEY.md section 8a P8.
G_GALACTIC = 4.498502151469553e-12
KMS_TO_KPC_MYR = 1.0227121650537077e-3      # 1 km/s in kpc/Myr


class UnitSystem:
    def __init__(self, name, G):
        self.name = name
        self.G = G

    def __repr__(self):
        return f"<UnitSystem {self.name}>"


galactic = UnitSystem("galactic (kpc, Myr, Msun, rad)", G_GALACTIC)
dimensionless = UnitSystem("dimensionless", 1.0)
# au, yr, Msun (reference units.py ``solarsystem``): G = 6.6743e-11 m^3 kg^-1 s^-2 x 1.988409870698051e30 kg x (365.25 d)^2 / au^3
solarsystem = UnitSystem("solarsystem (au, yr, Msun, rad)", 39.476926408897626)


def strip(x, units=galactic):
    """Return a plain float/ndarray from an optional astropy Quantity (duck-typed)."""
    if hasattr(x, "decompose") and hasattr(x, "unit"):
        try:
            return x.decompose(units).value
        except Exception:
            return x.value
    return x
