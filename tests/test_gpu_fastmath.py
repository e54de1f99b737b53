"""Error in ulp of the fast build's FP64 primitives (gala_b200/csrc/fastmath.cuh), measured on the
device against IEEE / mpmath-grade references.  The fast kernels replace the CUDA library's div/sqrt/
rsqrt/log expansions by MUFU seed + fixed refinement; this pins how much accuracy that costs."""
import numpy as np
import pytest

from gala_b200 import _abi

pytestmark = pytest.mark.gpu


def _ulp_err(y, ref):
    return np.abs(y - ref) / np.spacing(np.abs(ref))


def _inputs(seed, n=400_000):
    rng = np.random.default_rng(seed)
    return np.concatenate([
        np.exp(rng.uniform(np.log(1e-6), np.log(1e9), n)),     # the dynamic range of kpc/Myr/Msun orbits
        1.0 + rng.uniform(0, 1, n // 4),                        # one binade, dense
        2.0 ** rng.integers(-30, 30, 1000).astype(float),       # exact powers of two
    ])


def _longdouble_ref(which, x):
    xl = x.astype(np.longdouble)
    if which == 0:
        return (1 / xl)
    if which == 1:
        return 1 / np.sqrt(xl)
    if which == 2:
        return 1 / (xl * np.sqrt(xl))
    return np.log(xl)


@pytest.mark.parametrize("which,name,bound", [(0, "rcp", 2.0), (1, "rsqrt", 2.0), (2, "pow_m1p5", 2.5), (3, "log", 3.0)])
def test_fast_primitive_ulp(which, name, bound):
    x = _inputs(which)
    if which == 3:
        x = np.concatenate([x, 1.0 + 10.0 ** np.random.default_rng(9).uniform(-12, 0, 100_000)])
    y = _abi.math_probe(which, x, strict=False)
    ref = _longdouble_ref(which, x)
    err = np.abs(y.astype(np.longdouble) - ref) / np.spacing(np.abs(ref.astype(np.float64)))
    err = err.astype(np.float64)
    print(f"{name}: max {err.max():.3f} ulp, mean {err.mean():.3f}, p99.9 {np.quantile(err, 0.999):.3f}")
    assert np.all(np.isfinite(y))
    assert err.max() <= bound, f"{name}: {err.max()} ulp"


@pytest.mark.parametrize("which", [0, 1, 3])
def test_strict_primitive_is_ieee(which):
    x = _inputs(10 + which, 100_000)
    y = _abi.math_probe(which, x, strict=True)
    # IEEE division and sqrt are correctly rounded, so the strict build must equal numpy bit for bit;
    # CUDA's log is documented at <= 1 ulp
    if which == 0:
        assert np.array_equal(y, 1.0 / x)
    elif which == 1:
        assert np.array_equal(y, 1.0 / np.sqrt(x))
    else:
        ref = _longdouble_ref(which, x)
        err = (np.abs(y.astype(np.longdouble) - ref) / np.spacing(np.abs(ref.astype(np.float64)))).astype(np.float64)
        assert err.max() <= 1.0
