"""ctypes binding of the C ABI in ``include/gala_b200.h``.

The shared library ``gala_b200/libgala_b200.so`` is built in-tree by
``__graft_entry__.build()`` (``make -C gala_b200/csrc``).  There is no CPU
fallback: if the library is missing, loading raises; if no CUDA device is
present, every compute entry point returns -10 and the wrappers raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GALA_B200_LIB selects an alternative build of the same ABI (A/B experiments on the GPU box)
LIB_PATH = os.environ.get("GALA_B200_LIB") or os.path.join(_HERE, "libgala_b200.so")

# enum gb_pot_type
POT_NULL, POT_HERNQUIST, POT_NFW_SPHERICAL, POT_NFW_FLATTENED, POT_NFW_TRIAXIAL = 0, 1, 2, 3, 4
POT_MIYAMOTONAGAI, POT_MN3, POT_LONGMURALIBAR, POT_SCF = 5, 6, 7, 8
POT_KEPLER, POT_PLUMMER, POT_ISOCHRONE, POT_JAFFE, POT_MULTIPOLE = 9, 10, 11, 12, 13
POT_STONE, POT_BURKERT, POT_SATOH, POT_KUZMIN, POT_LOGARITHMIC, POT_LEESUTO, POT_POWERLAWCUTOFF = 14, 15, 16, 17, 18, 19, 20
POT_TIMEINTERP = 21
FRAME_STATIC, FRAME_ROTATING_3D = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1

# enum gb_extrema_row: rows of the (EXT_NSTAT, N) statistics array of gb_orbit_extrema / gb_integrate_extrema
EXT_ROWS = ("n_peri", "peri_mean", "peri_min", "peri_max", "peri_t_first", "peri_t_last",
            "n_apo", "apo_mean", "apo_min", "apo_max", "apo_t_first", "apo_t_last",
            "E_first", "E_last", "dE_max", "abs_z_max",
            "n_zmax", "zmax_mean", "zmax_min", "zmax_max", "zmax_t_first", "zmax_t_last")
EXT_NSTAT = 22

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class gb_component(C.Structure):
    _fields_ = [("type_id", C.c_int32), ("n_params", C.c_int32), ("do_shift_rotate", C.c_int32),
                ("_pad", C.c_int32), ("params", c_double_p), ("q0", C.c_double * 3), ("R", C.c_double * 9)]


class gb_potential(C.Structure):
    _fields_ = [("n_components", C.c_int32), ("n_dim", C.c_int32), ("comp", C.POINTER(gb_component))]


class gb_frame(C.Structure):
    _fields_ = [("type_id", C.c_int32), ("_pad", C.c_int32), ("omega", C.c_double * 3)]


class gb_launch(C.Structure):
    _fields_ = [("mem", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p),
                ("strict_math", C.c_int32), ("block_threads", C.c_int32),
                ("n_devices", C.c_int32), ("_pad", C.c_int32), ("devices", c_int32_p)]


class gb_bodies(C.Structure):
    _fields_ = [("n_bodies", C.c_int32), ("_pad", C.c_int32), ("body_pot", C.POINTER(gb_potential))]


class gb_dop853_stats(C.Structure):
    _fields_ = [("nstep", c_int32_p), ("naccpt", c_int32_p), ("nrejct", c_int32_p), ("nfcn", c_int32_p)]


P = C.POINTER
# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "gb_gradient": (C.c_int, [P(gb_potential), C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, P(gb_launch)]),
    "gb_energy": (C.c_int, [P(gb_potential), C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, P(gb_launch)]),
    "gb_density": (C.c_int, [P(gb_potential), C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, P(gb_launch)]),
    "gb_hessian": (C.c_int, [P(gb_potential), C.c_void_p, C.c_double, C.c_size_t, C.c_void_p, P(gb_launch)]),
    "gb_hamiltonian_energy": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_double, C.c_size_t,
                                        C.c_void_p, P(gb_launch)]),
    "gb_hamiltonian_gradient": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_double, C.c_size_t,
                                          C.c_void_p, P(gb_launch)]),
    "gb_leapfrog": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                              C.c_int, C.c_void_p, P(gb_launch)]),
    "gb_ruth4": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                           C.c_int, C.c_void_p, P(gb_launch)]),
    "gb_dop853": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                            C.c_double, C.c_double, C.c_long, C.c_double, C.c_long, C.c_int, C.c_void_p,
                            C.c_void_p, P(gb_dop853_stats), P(gb_launch)]),
    "gb_fardal_release": (C.c_int, [P(gb_potential), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                    P(gb_launch)]),
    "gb_stream_release": (C.c_int, [P(gb_potential), C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                    C.c_void_p, P(gb_launch)]),
    "gb_mockstream_dop853": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_double, C.c_double, C.c_double, C.c_double, C.c_long, C.c_void_p,
                                       C.c_void_p, P(gb_launch)]),
    "gb_mockstream_dop853_animate": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_void_p, C.c_size_t,
                                               C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_long, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, P(gb_launch)]),
    "gb_mockstream_leapfrog": (C.c_int, [P(gb_potential), C.c_void_p, C.c_void_p, C.c_size_t, C.c_double,
                                         C.c_double, C.c_void_p, P(gb_launch)]),
    "gb_nbody_leapfrog": (C.c_int, [P(gb_potential), P(gb_bodies), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, P(gb_launch)]),
    "gb_nbody_dop853": (C.c_int, [P(gb_potential), P(gb_bodies), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_long, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p, P(gb_launch)]),
    "gb_nbody_dop853_animate": (C.c_int, [P(gb_potential), P(gb_bodies), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_long, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, P(gb_launch)]),
    "gb_lyapunov_max": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                  C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_long, C.c_void_p,
                                  C.c_void_p, C.c_void_p, P(gb_launch)]),
    "gb_orbit_extrema": (C.c_int, [P(gb_potential), P(gb_frame), C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int,
                                   C.c_void_p, P(gb_launch)]),
    "gb_orbit_extrema_list": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, P(gb_launch)]),
    "gb_integrate_extrema": (C.c_int, [P(gb_potential), P(gb_frame), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, P(gb_launch)]),
    "gb_last_error": (C.c_char_p, []),
    "gb_device_count": (C.c_int, []),
    "gb_launch_count": (C.c_long, []),
    "gb_version": (C.c_char_p, []),
    "gb_fp64_peak_tflops": (C.c_double, [C.c_int]),
    "gb_release_scratch": (C.c_int, []),
    "gb_shard_bounds": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gb_deal_count": (C.c_long, [C.c_size_t, C.c_int, C.c_int]),
    "gb_math_probe": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, P(gb_launch)]),
}

_lib = None


def lib():
    """Load (once) and return the C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C gala_b200/csrc`.  gala_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class GalaB200Error(RuntimeError):
    pass


def check(rc: int):
    """Map a C-ABI return code to the exception type the reference raises on that condition."""
    if rc == 0:
        return
    msg = lib().gb_last_error().decode()
    if rc == -12:
        raise ValueError(msg)                      # dop853.pyx:143-152
    if rc == -14:
        raise NotImplementedError(msg)             # potential/potential/core.py:572-575
    if rc == -13:
        raise TypeError(msg)                       # leapfrog.pyx:64-68, ruth4.pyx:49-52
    if rc in (-1, -2, -3, -4):
        raise RuntimeError(msg)                    # dop853.pyx:184-185
    raise GalaB200Error(f"[{rc}] {msg}")


def device_count() -> int:
    return lib().gb_device_count()


def launch_count() -> int:
    return lib().gb_launch_count()


def release_scratch():
    """Free the device buffers the library caches between calls (see ``gb_release_scratch``)."""
    check(lib().gb_release_scratch())


def math_probe(which: int, x, strict: bool = False):
    """y = f(x) with the math primitive the selected build uses inside its kernels
    (0: 1/x, 1: x^-1/2, 2: x^-3/2, 3: ln x); host arrays.  Diagnostic for the test suite."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    o = launch_opts(False, strict=strict)
    check(lib().gb_math_probe(which, x.ctypes.data, x.size, y.ctypes.data, C.byref(o)))
    return y


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class Buf:
    """A host (numpy) or device (torch.cuda tensor) float64/int32 buffer seen as a raw pointer."""

    def __init__(self, arr):
        self.arr = arr
        self.device = _is_torch_cuda(arr)
        if self.device:
            if not arr.is_contiguous():
                raise ValueError("device tensors must be contiguous")
            self.ptr = arr.data_ptr()
        else:
            self.ptr = arr.ctypes.data


_devices = None      # process-wide default device list for host-buffer calls (set_devices)


def set_devices(devices=None):
    """Shard every host-buffer (numpy) call of the sharding entry points -- ``gb_leapfrog``, ``gb_ruth4``,
    ``gb_dop853``, ``gb_mockstream_dop853 / _leapfrog``, potential / Hamiltonian evaluation -- over these CUDA
    devices inside the C ABI (``gb_launch.n_devices``): contiguous orbit slices, one host thread per device, the
    results land in the one output array.  ``"all"`` = every visible device; ``None`` / ``[]`` = the current device
    only (default).  The environment variable ``GALA_B200_DEVICES`` ("all" or "0,1,2,3") sets the initial value."""
    global _devices
    if devices is None or (not isinstance(devices, str) and len(devices) == 0):
        _devices = None
    elif isinstance(devices, str) and devices.strip().lower() == "all":
        n = device_count()
        _devices = list(range(n)) if n > 1 else None
    else:
        if isinstance(devices, str):
            devices = [int(x) for x in devices.split(",") if x.strip() != ""]
        devs = [int(d) for d in devices]
        if len(set(devs)) != len(devs) or any(d < 0 for d in devs):
            raise ValueError("devices must be distinct non-negative CUDA ordinals")
        _devices = devs
    return _devices


def get_devices():
    return None if _devices is None else list(_devices)


def launch_opts(device_mem: bool, strict: bool = False, stream=None, block: int = 0, device: int = -1,
                devices=False):
    """``devices``: list of CUDA ordinals to shard a host-buffer call over; ``None`` = the process default of
    ``set_devices`` (what the wrappers of the sharding entry points pass); ``False`` = a single device (entry
    points that run on one device reject a device list)."""
    o = gb_launch()
    o.mem = MEM_DEVICE if device_mem else MEM_HOST
    o.device = device
    o.stream = stream
    o.strict_math = 1 if strict else 0
    o.block_threads = block
    if devices is None:
        devices = _devices
    if devices and not device_mem and len(devices) > 0:
        arr = (C.c_int32 * len(devices))(*devices)
        o._keep = arr                      # the struct only holds a pointer
        o.n_devices = len(devices)
        o.devices = C.cast(arr, c_int32_p)
    return o


def as_f64(a, shape_ndim=None):
    """C-contiguous float64 view/copy of a numpy-like array (host path)."""
    return np.ascontiguousarray(a, dtype=np.float64)


if os.environ.get("GALA_B200_DEVICES"):
    try:
        set_devices(os.environ["GALA_B200_DEVICES"])
    except Exception as _e:            # a bad value must not make the package unimportable
        import warnings
        warnings.warn(f"GALA_B200_DEVICES ignored: {_e}")
