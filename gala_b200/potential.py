"""Host-side mirror of the reference's potential classes for the hot path.

Class names, parameter names and the C parameter layout follow the reference
(``potential/potential/builtin/core.py``, ``special.py``, ``cybuiltin.pyx``,
``ccompositepotential.pyx``): every object exposes ``G``, ``c_parameters`` (so that
``[G] + c_parameters`` is exactly ``CPotentialWrapper._params``, ``cpotential.pyx:306-316``),
``origin`` and ``R``, and ``spec()`` turns it into the flat ``gb_potential`` of the C ABI.
All numerics run in the CUDA library; nothing here evaluates a potential on the CPU.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _abi
from .units import galactic, strip, KMS_TO_KPC_MYR

__all__ = [
    "PotentialBase", "CCompositePotential", "NullPotential", "KeplerPotential", "HernquistPotential",
    "PlummerPotential", "IsochronePotential", "JaffePotential", "NFWPotential", "MiyamotoNagaiPotential",
    "MN3ExponentialDiskPotential", "LongMuraliBarPotential", "SCFPotential", "MultipolePotential", "MilkyWayPotential",
    "MilkyWayPotential2022", "StonePotential", "BurkertPotential", "SatohPotential", "KuzminPotential",
    "LogarithmicPotential", "LeeSutoTriaxialNFWPotential", "PowerLawCutoffPotential", "LM10Potential",
    "BovyMWPotential2014", "TimeInterpolatedPotential",
]


class _Spec:
    """Owns the ctypes structs + parameter arrays of one gb_potential (kept alive with it)."""

    def __init__(self, components):
        self.n = len(components)
        self._arrays = []
        self.comps = (_abi.gb_component * self.n)()
        for i, (type_id, params, origin, R) in enumerate(components):
            p = np.ascontiguousarray(params, dtype=np.float64)
            self._arrays.append(p)
            c = self.comps[i]
            c.type_id = type_id
            c.n_params = p.size
            c.params = p.ctypes.data_as(_abi.c_double_p)
            origin = np.zeros(3) if origin is None else np.asarray(origin, dtype=np.float64)
            R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
            # do_shift_rotate = not (q0 == 0 and R == I)   (cpotential.pyx:79-92)
            c.do_shift_rotate = int(not (np.all(origin == 0.0) and np.array_equal(R, np.eye(3))))
            for k in range(3):
                c.q0[k] = origin[k]
            for k in range(9):
                c.R[k] = R.ravel()[k]
        self.pot = _abi.gb_potential(self.n, 3, C.cast(self.comps, C.POINTER(_abi.gb_component)))

    def ptr(self):
        return C.byref(self.pot)


def _prep_q(q):
    """(3,) or (3,N) host array / device tensor -> (Buf of shape (3,N), orig_shape)."""
    if _abi._is_torch_cuda(q):
        import torch
        if q.dtype != torch.float64:
            raise TypeError("device tensors must be float64")
        shape = tuple(q.shape)
        q2 = q.reshape(shape[0], -1).contiguous()
        return _abi.Buf(q2), shape
    a = np.asarray(strip(q), dtype=np.float64)
    shape = a.shape
    a2 = np.ascontiguousarray(a.reshape(shape[0], -1))
    return _abi.Buf(a2), shape


def _alloc_like(buf, shape, dtype="f8"):
    if buf.device:
        import torch
        return torch.empty(shape, dtype=torch.float64 if dtype == "f8" else torch.int32, device=buf.arr.device)
    return np.empty(shape, dtype=np.float64 if dtype == "f8" else np.int32)


def _stream_of(buf):
    if buf.device:
        import torch
        return torch.cuda.current_stream(buf.arr.device).cuda_stream, buf.arr.device.index
    return None, -1


class PotentialBase:
    """Common behaviour (reference ``potential/potential/core.py`` PotentialBase, C-enabled subset)."""

    ndim = 3
    _type_id = None
    _param_names = ()

    def __init__(self, *, units=galactic, origin=None, R=None, **params):
        self.units = units
        self.G = units.G if units is not None else 1.0
        self.origin = np.zeros(3) if origin is None else np.asarray(strip(origin), dtype=np.float64)
        self.R = None if R is None else np.asarray(R, dtype=np.float64)
        if self.R is not None and self.R.shape != (3, 3):
            raise ValueError("Rotation matrix must be 3x3")
        self.parameters = OrderedDict((k, float(strip(params[k]))) for k in self._param_names)
        self.c_parameters = self._c_parameters()
        self._spec = None
        self.strict_math = False

    # -- C parameter vector (without G), in the order the reference's wrapper receives it
    def _c_parameters(self):
        return np.array([self.parameters[k] for k in self._param_names], dtype=np.float64)

    @property
    def _R(self):
        return np.eye(3) if self.R is None else self.R

    def _components(self):
        return [(self._type_id, np.concatenate([[self.G], self.c_parameters]), self.origin, self.R)]

    def spec(self):
        if self._spec is None:
            self._spec = _Spec(self._components())
        return self._spec

    c_enabled = True

    # -- evaluation ---------------------------------------------------------------------------
    def _eval(self, fn_name, q, t, out_rows):
        buf, shape = _prep_q(q)
        if shape[0] != 3:
            raise ValueError(f"position array must have shape (3, ...), got {shape}")
        N = buf.arr.shape[1]
        out = _alloc_like(buf, (out_rows, N) if out_rows > 1 else (N,))
        stream, dev = _stream_of(buf)
        opt = _abi.launch_opts(buf.device, self.strict_math, stream, device=dev,
                               devices=False if fn_name == "gb_hessian" else None)
        fn = getattr(_abi.lib(), fn_name)
        _abi.check(fn(self.spec().ptr(), buf.ptr, float(strip(t)), N, _abi.Buf(out).ptr, C.byref(opt)))
        if out_rows > 1:
            return out.reshape((out_rows,) + tuple(shape[1:]))
        return out.reshape(tuple(shape[1:]))

    def gradient(self, q, t=0.0):
        """dPhi/dq at q (3,N); mirrors ``PotentialBase.gradient`` (core.py:425-479)."""
        return self._eval("gb_gradient", q, t, 3)

    def acceleration(self, q, t=0.0):
        return -self.gradient(q, t)

    def energy(self, q, t=0.0):
        return self._eval("gb_energy", q, t, 1)

    def density(self, q, t=0.0):
        return self._eval("gb_density", q, t, 1)

    def hessian(self, q, t=0.0):
        """d2Phi/dq_i dq_j at q (3,N) -> (3,3,N); mirrors ``PotentialBase.hessian`` (core.py:535-600).
        Rotated potentials raise ``NotImplementedError`` like the reference."""
        h = self._eval("gb_hessian", q, t, 9)
        return h.reshape((3, 3) + tuple(h.shape[1:]))

    def __call__(self, q, t=0.0):
        return self.energy(q, t)

    def save(self, f):
        """``PotentialBase.save`` (core.py:1186-1200): write the YAML specification (``gala_b200.io.save``, potential_io.py)."""
        from .potential_io import save
        save(self, f)

    def mass_enclosed(self, q, t=0.0):
        """``PotentialBase.mass_enclosed`` (core.py:649-723): r^2 |dPhi/dr| / G from a centred difference of the
        potential along the radius with the reference's step h = 1e-3, negative for a negative ``m`` parameter.
        Two batched energy evaluations on the device; numpy in -> numpy out, torch.cuda in -> stays on the device."""
        h = 1e-3
        r = (q * q).sum(0) ** 0.5
        eps = h * q / r
        diff = self.energy(q + eps, t) - self.energy(q - eps, t)
        m = getattr(self, "parameters", {}).get("m", None)
        sgn = -1.0 if (m is not None and np.ndim(m) == 0 and m < 0) else 1.0
        return sgn * abs(r * r * diff / self.G / (2.0 * h))

    def circular_velocity(self, q, t=0.0):
        """``PotentialBase.circular_velocity`` (core.py:725-784): sqrt(r |grad Phi . rhat|), one batched gradient."""
        r = (q * q).sum(0) ** 0.5
        dPhi_dr = (self.gradient(q, t) * q / r).sum(0)
        return (r * abs(dPhi_dr)) ** 0.5

    # -- orbit integration (core.py:1150-1184 -> Hamiltonian.integrate_orbit) --------------------
    def integrate_orbit(self, w0, Integrator=None, Integrator_kwargs=None, cython_if_possible=True,
                        save_all=True, **time_spec):
        from .hamiltonian import Hamiltonian
        return Hamiltonian(self).integrate_orbit(w0, Integrator=Integrator, Integrator_kwargs=Integrator_kwargs,
                                                 cython_if_possible=cython_if_possible, save_all=save_all,
                                                 **time_spec)

    def __add__(self, other):
        new = CCompositePotential()
        for i, p in enumerate((self, other)):
            if isinstance(p, CCompositePotential):
                for k, v in p.items():
                    new[k] = v
            else:
                new[f"c{len(new)}"] = p
        return new

    def __repr__(self):
        pars = ", ".join(f"{k}={v:g}" for k, v in self.parameters.items())
        return f"<{self.__class__.__name__}: {pars}>"


class NullPotential(PotentialBase):
    _type_id = _abi.POT_NULL


class KeplerPotential(PotentialBase):
    _type_id = _abi.POT_KEPLER
    _param_names = ("m",)

    def __init__(self, m, **kw):
        super().__init__(m=m, **kw)


class HernquistPotential(PotentialBase):
    _type_id = _abi.POT_HERNQUIST
    _param_names = ("m", "c")

    def __init__(self, m, c, **kw):
        super().__init__(m=m, c=c, **kw)


class PlummerPotential(PotentialBase):
    _type_id = _abi.POT_PLUMMER
    _param_names = ("m", "b")

    def __init__(self, m, b, **kw):
        super().__init__(m=m, b=b, **kw)


class IsochronePotential(PotentialBase):
    _type_id = _abi.POT_ISOCHRONE
    _param_names = ("m", "b")

    def __init__(self, m, b, **kw):
        super().__init__(m=m, b=b, **kw)


class JaffePotential(PotentialBase):
    _type_id = _abi.POT_JAFFE
    _param_names = ("m", "c")

    def __init__(self, m, c, **kw):
        super().__init__(m=m, c=c, **kw)


class NFWPotential(PotentialBase):
    """NFW; wrapper chosen like ``NFWPotential._setup_potential`` (builtin/core.py:722-741)."""
    _param_names = ("m", "r_s", "a", "b", "c")

    def __init__(self, m, r_s, a=1.0, b=1.0, c=1.0, **kw):
        super().__init__(m=m, r_s=r_s, a=a, b=b, c=c, **kw)
        a, b, c = (self.parameters[k] for k in "abc")
        if np.allclose([a, b, c], 1.0):
            self._type_id = _abi.POT_NFW_SPHERICAL
        elif np.allclose([a, b], 1.0):
            self._type_id = _abi.POT_NFW_FLATTENED
        else:
            self._type_id = _abi.POT_NFW_TRIAXIAL

    @classmethod
    def from_circular_velocity(cls, v_c, r_s, a=1.0, b=1.0, c=1.0, r_ref=None, units=galactic, **kw):
        """builtin/core.py ``NFWPotential.from_circular_velocity``: v_c in kpc/Myr at r_ref (default r_s)."""
        v_c, r_s = float(strip(v_c)), float(strip(r_s))
        r_ref = r_s if r_ref is None else float(strip(r_ref))
        uu = r_ref / r_s
        vs2 = v_c ** 2 / uu / (np.log(1 + uu) / uu ** 2 - 1 / (uu * (1 + uu)))
        m = vs2 * r_s / units.G
        return cls(m=m, r_s=r_s, a=a, b=b, c=c, units=units, **kw)


class MiyamotoNagaiPotential(PotentialBase):
    _type_id = _abi.POT_MIYAMOTONAGAI
    _param_names = ("m", "a", "b")

    def __init__(self, m, a, b, **kw):
        super().__init__(m=m, a=a, b=b, **kw)


class MN3ExponentialDiskPotential(PotentialBase):
    """Three Miyamoto-Nagai disks approximating an exponential disk (Smith et al. 2015, MNRAS 448,
    2934, tables 1 and 2).  Parameter precompute restates builtin/core.py:607-666; the C vector is
    ``[m1,a1,b1, m2,a2,b2, m3,a3,b3, m, h_R, h_z]`` (c_only first, cpotential.pyx:289-303)."""
    _type_id = _abi.POT_MN3
    _param_names = ("m", "h_R", "h_z")

    # Smith+2015 fitting-coefficient tables (positive-density and negative-density variants)
    _K_pos_dens = np.array([
        [0.0036, -0.0330, 0.1117, -0.1335, 0.1749],
        [-0.0131, 0.1090, -0.3035, 0.2921, -5.7976],
        [-0.0048, 0.0454, -0.1425, 0.1012, 6.7120],
        [-0.0158, 0.0993, -0.2070, -0.7089, 0.6445],
        [-0.0319, 0.1514, -0.1279, -0.9325, 2.6836],
        [-0.0326, 0.1816, -0.2943, -0.6329, 2.3193]])
    _K_neg_dens = np.array([
        [-0.0090, 0.0640, -0.1653, 0.1164, 1.9487],
        [0.0173, -0.0903, 0.0877, 0.2029, -1.3077],
        [-0.0051, 0.0287, -0.0361, -0.0544, 0.2242],
        [-0.0358, 0.2610, -0.6987, -0.1193, 2.0074],
        [-0.0830, 0.4992, -0.7967, -1.2966, 4.4441],
        [-0.0247, 0.1718, -0.4124, -0.5944, 0.7333]])

    def __init__(self, m, h_R, h_z, positive_density=True, sech2_z=True, **kw):
        self.positive_density = positive_density
        self.sech2_z = sech2_z
        super().__init__(m=m, h_R=h_R, h_z=h_z, **kw)

    def _c_parameters(self):
        m, h_R, h_z = (self.parameters[k] for k in ("m", "h_R", "h_z"))
        hzR = h_z / h_R
        K = self._K_pos_dens if self.positive_density else self._K_neg_dens
        if self.sech2_z:
            b_hR = -0.033 * hzR ** 3 + 0.262 * hzR ** 2 + 0.659 * hzR
        else:
            b_hR = -0.269 * hzR ** 3 + 1.08 * hzR ** 2 + 1.092 * hzR
        x = np.vander([b_hR], N=5)[0]
        param_vec = K @ x
        self._ms = param_vec[:3] * m
        self._as = param_vec[3:] * h_R
        self._b = b_hR * h_R
        c_only = []
        for i in range(3):
            c_only += [self._ms[i], self._as[i], self._b]
        return np.array(c_only + [m, h_R, h_z], dtype=np.float64)

    def get_three_potentials(self):
        return {f"disk{i + 1}": MiyamotoNagaiPotential(m=self._ms[i], a=self._as[i], b=self._b, units=self.units,
                                                       origin=self.origin, R=self.R) for i in range(3)}


class LongMuraliBarPotential(PotentialBase):
    """Long & Murali (1992) bar; ``alpha`` in radians (builtin/core.py LongMuraliBarPotential)."""
    _type_id = _abi.POT_LONGMURALIBAR
    _param_names = ("m", "a", "b", "c", "alpha")

    def __init__(self, m, a, b, c, alpha=0.0, **kw):
        super().__init__(m=m, a=a, b=b, c=c, alpha=alpha, **kw)


class StonePotential(PotentialBase):
    """Stone & Ostriker (2015) (builtin/core.py:306-328); C vector [m, r_c, r_h]."""
    _type_id = _abi.POT_STONE
    _param_names = ("m", "r_c", "r_h")

    def __init__(self, m, r_c, r_h, **kw):
        super().__init__(m=m, r_c=r_c, r_h=r_h, **kw)


class BurkertPotential(PotentialBase):
    """Burkert (builtin/core.py:415-432); C vector [rho, r0]."""
    _type_id = _abi.POT_BURKERT
    _param_names = ("rho", "r0")

    def __init__(self, rho, r0, **kw):
        super().__init__(rho=rho, r0=r0, **kw)


class SatohPotential(PotentialBase):
    _type_id = _abi.POT_SATOH
    _param_names = ("m", "a", "b")

    def __init__(self, m, a, b, **kw):
        super().__init__(m=m, a=a, b=b, **kw)


class KuzminPotential(PotentialBase):
    _type_id = _abi.POT_KUZMIN
    _param_names = ("m", "a")

    def __init__(self, m, a, **kw):
        super().__init__(m=m, a=a, **kw)


class LogarithmicPotential(PotentialBase):
    """Triaxial logarithmic halo (builtin/core.py:937-968); C vector [v_c, r_h, q1, q2, q3, phi]."""
    _type_id = _abi.POT_LOGARITHMIC
    _param_names = ("v_c", "r_h", "q1", "q2", "q3", "phi")

    def __init__(self, v_c, r_h, q1=1.0, q2=1.0, q3=1.0, phi=0.0, **kw):
        super().__init__(v_c=v_c, r_h=r_h, q1=q1, q2=q2, q3=q3, phi=phi, **kw)


class LeeSutoTriaxialNFWPotential(PotentialBase):
    """Lee & Suto (2003) (builtin/core.py:983-1013); C vector [v_c, r_s, a, b, c]."""
    _type_id = _abi.POT_LEESUTO
    _param_names = ("v_c", "r_s", "a", "b", "c")

    def __init__(self, v_c, r_s, a=1.0, b=1.0, c=1.0, **kw):
        super().__init__(v_c=v_c, r_s=r_s, a=a, b=b, c=c, **kw)


class PowerLawCutoffPotential(PotentialBase):
    """Power law with exponential cutoff (builtin/core.py:347-374); C vector [m, alpha, r_c], alpha < 3."""
    _type_id = _abi.POT_POWERLAWCUTOFF
    _param_names = ("m", "alpha", "r_c")

    def __init__(self, m, alpha, r_c, **kw):
        super().__init__(m=m, alpha=alpha, r_c=r_c, **kw)


class SCFPotential(PotentialBase):
    """Hernquist-Ostriker basis-function expansion (reference potential/scf/core.py SCFPotential):
    ``Snlm``/``Tnlm`` have shape (nmax+1, lmax+1, lmax+1); the C vector is
    ``[nmax, lmax, m, r_s, Snlm.ravel(), Tnlm.ravel()]`` (scf/bfe.cpp:229-258)."""
    _type_id = _abi.POT_SCF
    _param_names = ("m", "r_s")

    def __init__(self, m, r_s, Snlm, Tnlm=None, **kw):
        self.Snlm = np.ascontiguousarray(Snlm, dtype=np.float64)
        self.Tnlm = np.zeros_like(self.Snlm) if Tnlm is None else np.ascontiguousarray(Tnlm, dtype=np.float64)
        if self.Snlm.ndim != 3 or self.Snlm.shape[1] != self.Snlm.shape[2] or self.Tnlm.shape != self.Snlm.shape:
            raise ValueError("Snlm/Tnlm must have shape (nmax+1, lmax+1, lmax+1)")
        self.nmax = self.Snlm.shape[0] - 1
        self.lmax = self.Snlm.shape[1] - 1
        super().__init__(m=m, r_s=r_s, **kw)

    def _c_parameters(self):
        return np.concatenate([[self.nmax, self.lmax, self.parameters["m"], self.parameters["r_s"]],
                               self.Snlm.ravel(), self.Tnlm.ravel()])


class MultipolePotential(PotentialBase):
    """Inner / outer multipole expansion (reference ``builtin/core.py:1101-1239`` MultipolePotential,
    ``builtin/multipole.cpp``).  ``MultipolePotential(lmax=2, m=..., r_s=..., inner=False, S10=5., T21=...)``:
    unspecified ``S{l}{m}`` / ``T{l}{m}`` default to 0, unknown names raise.  The C vector is
    ``[lmax, n_coeffs, inner, m, r_s, S00, T00, S10, T10, S11, T11, ...]`` (c-only parameters first,
    ``core.py:1220-1222``; then declaration order ``:1122-1146``)."""
    _type_id = _abi.POT_MULTIPOLE
    _param_names = ("m", "r_s")

    def __init__(self, *args, lmax=None, m=1.0, r_s=1.0, inner=False, **kw):
        if lmax is None:
            raise TypeError("Can't initialize a MultipolePotential without specifying the `lmax` keyword argument.")
        if args:
            m = args[0]
            if len(args) > 1:
                r_s = args[1]
        self.lmax = int(lmax)
        self.inner = bool(inner)
        self.coeffs = OrderedDict()
        for l in range(self.lmax + 1):
            for mm in range(l + 1):
                self.coeffs[f"S{l}{mm}"] = float(kw.pop(f"S{l}{mm}", 0.0))
                self.coeffs[f"T{l}{mm}"] = float(kw.pop(f"T{l}{mm}", 0.0))
        bad = [k for k in kw if k[:1] in "ST" and k[1:].isdigit()]
        if bad:
            raise ValueError(f"coefficient(s) {bad} not allowed for lmax={self.lmax}")
        super().__init__(m=m, r_s=r_s, **kw)
        self.parameters["inner"] = self.inner
        self.parameters.update(self.coeffs)

    def _c_parameters(self):
        n_coeffs = (self.lmax + 1) * (self.lmax + 2) // 2
        return np.concatenate([[self.lmax, n_coeffs, float(self.inner), self.parameters["m"], self.parameters["r_s"]],
                               np.array(list(self.coeffs.values()), dtype=np.float64)])


class TimeInterpolatedPotential(PotentialBase):
    """``TimeInterpolatedPotential`` (reference ``builtin/time_interpolated.py:29``): any analytic builtin potential
    with parameters, origin and / or rotation given at ``time_knots`` and interpolated in between with one of GSL's
    spline types -- ``'linear'``, ``'cspline'`` (default), ``'akima'``, ``'steffen'``.  A parameter is constant (scalar)
    or time-varying (array of length ``n_knots``); ``origin`` is ``(3,)`` or ``(n_knots, 3)``; ``R`` is ``(3, 3)`` or
    ``(n_knots, 3, 3)`` (interpolated through its axis-angle form, ``time_interp.cpp:325-405``).  Evaluation outside
    ``[time_knots[0], time_knots[-1]]`` gives NaN, and ``integrate_orbit`` refuses such a time grid
    (``time_interpolated.py:466-474``).

    The C parameter rows are what the reference interpolates for classes with a parameter transform
    (``_requires_c_param_precompute``, ``time_interpolated.py:271-330``), applied to every class: the wrapped
    potential is instantiated at every knot and its ``c_parameters`` are the knot values.  The splines themselves are
    built inside the C ABI (``GB_POT_TIMEINTERP``)."""
    _type_id = _abi.POT_TIMEINTERP
    _METHODS = {"linear": (0, 2), "cspline": (1, 3), "akima": (2, 5), "steffen": (3, 3)}
    _UNSUPPORTED = ("NullPotential", "MultipolePotential", "SCFPotential", "CCompositePotential", "TimeInterpolatedPotential")

    def __init__(self, potential_cls, time_knots, interpolation_method="cspline", units=galactic, origin=None, R=None,
                 **kwargs):
        if not (isinstance(potential_cls, type) and issubclass(potential_cls, PotentialBase)):
            raise TypeError("potential_cls must be a gala_b200 potential class")
        if potential_cls.__name__ in self._UNSUPPORTED or issubclass(potential_cls, CCompositePotential):
            raise NotImplementedError(f"TimeInterpolatedPotential does not currently support {potential_cls.__name__}.")
        if interpolation_method not in self._METHODS:
            raise ValueError(f"Interpolation method '{interpolation_method}' is not recognized. Supported methods are: "
                             f"{list(self._METHODS)}")
        tk = np.ascontiguousarray(strip(time_knots), dtype=np.float64)
        if tk.ndim != 1:
            raise ValueError("time_knots must be one-dimensional")
        n = tk.size
        code, need = self._METHODS[interpolation_method]
        if n < need:
            raise ValueError(f"Interpolation method '{interpolation_method}' requires at least {need} time knots, but only "
                             f"{n} were provided. Either provide more time knots or use 'linear' interpolation.")
        if not np.all(np.diff(tk) > 0):
            raise ValueError("time_knots must be monotonically increasing (and no duplicate times)")
        self.potential_cls, self.time_knots, self.interpolation_method = potential_cls, tk, interpolation_method
        self.units = units
        self.G = units.G if units is not None else 1.0
        self._interp_params = []
        knot_kwargs = [dict() for _ in range(n)]
        for k, v in kwargs.items():
            a = np.asarray(strip(v), dtype=np.float64) if not isinstance(v, (bool, str)) else v
            if isinstance(a, np.ndarray) and a.ndim >= 1 and k in getattr(potential_cls, "_param_names", ()):
                if a.shape[0] != n:
                    raise ValueError(f"Parameter '{k}' has shape {a.shape} but there are {n} time knots. For "
                                     "time-interpolated parameters, the first dimension must match the number of time knots.")
                self._interp_params.append(k)
                for i in range(n):
                    knot_kwargs[i][k] = a[i]
            else:
                for i in range(n):
                    knot_kwargs[i][k] = v
        knots = [potential_cls(units=units, **kw) for kw in knot_kwargs]
        if len({p._type_id for p in knots}) != 1:
            raise ValueError("the wrapped potential must resolve to the same C type at every time knot")
        self._wrapped_type = knots[0]._type_id
        self._rows = np.ascontiguousarray([p.c_parameters for p in knots], dtype=np.float64)      # (n_knots, n_wpar)
        o = np.zeros((1, 3)) if origin is None else np.atleast_2d(np.asarray(strip(origin), dtype=np.float64))
        if o.shape not in ((1, 3), (n, 3)):
            raise ValueError(f"Origin array has wrong shape: expected (3,) or ({n}, 3)")
        Rm = np.eye(3)[None] if R is None else np.asarray(R, dtype=np.float64)
        Rm = Rm[None] if Rm.ndim == 2 else Rm
        if Rm.shape not in ((1, 3, 3), (n, 3, 3)):
            raise ValueError(f"Rotation matrices array has wrong shape {Rm.shape}")
        self._origins, self._Rs = np.ascontiguousarray(o), np.ascontiguousarray(Rm)
        self.origin, self.R = np.zeros(3), None          # they travel inside the parameter vector
        self.parameters = OrderedDict(potential_cls=potential_cls, time_knots=tk, interpolation_method=interpolation_method,
                                      **{k: kwargs[k] for k in kwargs})
        self.c_parameters = np.concatenate([[self._wrapped_type, code, n, self._rows.shape[1], o.shape[0], Rm.shape[0]],
                                            tk, self._rows.ravel(), o.ravel(), Rm.ravel()])
        self._spec = None
        self.strict_math = False

    @classmethod
    def from_tables(cls, wrapped_type, G, time_knots, rows, origins=None, Rs=None, interpolation_method="cspline",
                    units=galactic):
        """Build the device parameter vector from tables that already exist: ``rows`` (n_knots, n_wpar) = the wrapped
        potential's C parameter vector at every knot (one row if nothing varies), ``origins`` (1 | n_knots, 3), ``Rs``
        (1 | n_knots, 3, 3).  Used by ``gala_plugin`` for a gala ``TimeInterpolatedPotential``, whose knot potentials
        gala itself instantiates (time_interpolated.py:161-260)."""
        self = cls.__new__(cls)
        if interpolation_method not in cls._METHODS:
            raise ValueError(f"Interpolation method '{interpolation_method}' is not recognized. Supported methods are: "
                             f"{list(cls._METHODS)}")
        tk = np.ascontiguousarray(time_knots, dtype=np.float64)
        n = tk.size
        code, need = cls._METHODS[interpolation_method]
        if tk.ndim != 1 or n < need or not np.all(np.diff(tk) > 0):
            raise ValueError("time_knots must be one-dimensional, strictly increasing and long enough for the method")
        rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
        if rows.shape[0] == 1:
            rows = np.repeat(rows, n, axis=0)
        o = np.zeros((1, 3)) if origins is None else np.atleast_2d(np.asarray(origins, dtype=np.float64))
        Rm = np.eye(3)[None] if Rs is None else np.asarray(Rs, dtype=np.float64)
        Rm = Rm[None] if Rm.ndim == 2 else Rm
        if rows.shape[0] != n or o.shape not in ((1, 3), (n, 3)) or Rm.shape not in ((1, 3, 3), (n, 3, 3)):
            raise ValueError("tables must have one row per time knot (or a single constant row)")
        self.potential_cls, self.time_knots, self.interpolation_method = None, tk, interpolation_method
        self.units, self.G = units, float(G)
        self._interp_params, self._wrapped_type = [], int(wrapped_type)
        self._rows, self._origins, self._Rs = np.ascontiguousarray(rows), np.ascontiguousarray(o), np.ascontiguousarray(Rm)
        self.origin, self.R = np.zeros(3), None
        self.parameters = OrderedDict(time_knots=tk, interpolation_method=interpolation_method)
        self.c_parameters = np.concatenate([[self._wrapped_type, code, n, rows.shape[1], o.shape[0], Rm.shape[0]],
                                            tk, rows.ravel(), o.ravel(), Rm.ravel()])
        self._spec = None
        self.strict_math = False
        return self

    def _c_parameters(self):
        return self.c_parameters

    def hessian(self, q, t=0.0):
        """``time_interp_hessian`` (time_interp_wrapper.cpp:254-318) at time t.  With rotation matrices the reference's
        ``PotentialBase.hessian`` refuses (core.py:581-589), and so does this class; the C ABI itself returns
        R^T H R like the reference's C function (``gb_hessian``)."""
        if self._Rs.shape[0] > 1 or not np.array_equal(self._Rs[0], np.eye(3)):
            raise NotImplementedError("Computing Hessian matrices for rotated potentials is currently not supported.")
        return super().hessian(q, t)

    @property
    def time_bounds(self):
        return float(self.time_knots[0]), float(self.time_knots[-1])

    def integrate_orbit(self, w0, Integrator=None, Integrator_kwargs=None, cython_if_possible=True, save_all=True,
                        **time_spec):
        from .integrate import parse_time_specification
        t = parse_time_specification(self.units, **time_spec)
        t_min, t_max = self.time_bounds
        if np.any(t < t_min) or np.any(t > t_max):
            raise ValueError("Integration times must be within the range of the Potential's interpolation range "
                             f"that you defined: [{t_min}, {t_max}], your orbit integration range is [{min(t)}, {max(t)}]")
        return super().integrate_orbit(w0, Integrator=Integrator, Integrator_kwargs=Integrator_kwargs,
                                       cython_if_possible=cython_if_possible, save_all=save_all, t=t)

    def __repr__(self):
        return (f"<TimeInterpolatedPotential: {getattr(self.potential_cls, '__name__', f'C type {self._wrapped_type}')} "
                f"interpolation_method='{self.interpolation_method}')>")


class CCompositePotential(PotentialBase, OrderedDict):
    """Ordered collection of C-enabled potentials (reference ``ccompositepotential.pyx:27-86``);
    components are evaluated and summed in insertion order."""

    def __init__(self, **potentials):
        OrderedDict.__init__(self)
        self.units = galactic
        self.G = galactic.G
        self.origin = np.zeros(3)
        self.R = None
        self.parameters = OrderedDict()
        self.c_parameters = np.array([])
        self._spec = None
        self.strict_math = False
        self.lock = False
        for k, v in potentials.items():
            self[k] = v

    def __setitem__(self, key, value):
        if getattr(self, "lock", False):
            raise ValueError("Potential object is locked - new components can only be added to unlocked potentials.")
        if not isinstance(value, PotentialBase) or isinstance(value, CCompositePotential):
            raise TypeError("components must be (non-composite) C-enabled potentials")
        OrderedDict.__setitem__(self, key, value)
        self.units = value.units
        self.G = value.G
        self._spec = None

    def _components(self):
        comps = []
        for p in self.values():
            comps += p._components()
        return comps

    def __repr__(self):
        return "<CCompositePotential " + ",".join(self.keys()) + ">"

    # OrderedDict defines __eq__/__hash__ semantics we do not want to inherit for hashing
    __hash__ = object.__hash__


class MilkyWayPotential(CCompositePotential):
    """v1 Milky Way model (builtin/special.py:88-124 defaults): MN disk, two Hernquist spheres, NFW."""

    def __init__(self, units=galactic, disk=None, halo=None, bulge=None, nucleus=None, version="v1"):
        super().__init__()
        self.version = version
        if version in ("v2", "latest"):
            _setup_mwp_2022(self, units, disk, halo, bulge, nucleus)
        else:
            d = dict(m=6.8e10, a=3.0, b=0.28); d.update(disk or {})
            b = dict(m=5e9, c=1.0); b.update(bulge or {})
            n = dict(m=1.71e9, c=0.07); n.update(nucleus or {})
            h = dict(m=5.4e11, r_s=15.62); h.update(halo or {})
            self["disk"] = MiyamotoNagaiPotential(units=units, **d)
            self["bulge"] = HernquistPotential(units=units, **b)
            self["nucleus"] = HernquistPotential(units=units, **n)
            self["halo"] = NFWPotential(units=units, **h)
        self.lock = True


def _setup_mwp_2022(obj, units, disk=None, halo=None, bulge=None, nucleus=None):
    # defaults and component order: builtin/special.py:127-153
    d = dict(m=4.7717e10, h_R=2.6, h_z=0.3); d.update(disk or {})
    b = dict(m=5e9, c=1.0); b.update(bulge or {})
    n = dict(m=1.8142e9, c=0.0688867); n.update(nucleus or {})
    h = dict(m=5.5427e11, r_s=15.626); h.update(halo or {})
    obj["disk"] = MN3ExponentialDiskPotential(units=units, **d)
    obj["bulge"] = HernquistPotential(units=units, **b)
    obj["nucleus"] = HernquistPotential(units=units, **n)
    obj["halo"] = NFWPotential(units=units, **h)


class MilkyWayPotential2022(CCompositePotential):
    """builtin/special.py:221-271: MN3 disk + Hernquist bulge + Hernquist nucleus + NFW halo."""

    def __init__(self, units=galactic, disk=None, halo=None, bulge=None, nucleus=None):
        super().__init__()
        _setup_mwp_2022(self, units, disk, halo, bulge, nucleus)
        self.lock = True


class LM10Potential(CCompositePotential):
    """Law & Majewski (2010) (builtin/special.py:26-87): MiyamotoNagai disk + Hernquist bulge + triaxial
    Logarithmic halo, in that order; v_c = sqrt(2) * 121.858 km/s in kpc/Myr."""

    def __init__(self, units=galactic, disk=None, bulge=None, halo=None):
        super().__init__()
        d = dict(m=1e11, a=6.5, b=0.26); d.update(disk or {})
        b = dict(m=3.4e10, c=0.7); b.update(bulge or {})
        h = dict(q1=1.38, q2=1.0, q3=1.36, r_h=12.0, phi=np.deg2rad(97.0),
                 v_c=np.sqrt(2) * 121.858 * KMS_TO_KPC_MYR); h.update(halo or {})
        self["disk"] = MiyamotoNagaiPotential(units=units, **d)
        self["bulge"] = HernquistPotential(units=units, **b)
        self["halo"] = LogarithmicPotential(units=units, **h)
        self.lock = True


class BovyMWPotential2014(CCompositePotential):
    """galpy's MWPotential2014 (builtin/special.py:274-347): MiyamotoNagai disk + PowerLawCutoff bulge +
    NFW halo, in that order."""

    def __init__(self, units=galactic, disk=None, halo=None, bulge=None):
        super().__init__()
        d = dict(m=68193902782.346756, a=3.0, b=0.28); d.update(disk or {})
        b = dict(m=4501365375.06545, alpha=1.8, r_c=1.9); b.update(bulge or {})
        h = dict(m=4.3683325e11, r_s=16.0); h.update(halo or {})
        self["disk"] = MiyamotoNagaiPotential(units=units, **d)
        self["bulge"] = PowerLawCutoffPotential(units=units, **b)
        self["halo"] = NFWPotential(units=units, **h)
        self.lock = True
