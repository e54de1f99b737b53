"""`-m "not gpu"`: the C-ABI library loads, exports every symbol include/gala_b200.h declares, and
its compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gala_b200 as gb
from gala_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "gala_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", txt)))


def test_library_built_and_exports_header_symbols():
    L = _abi.lib()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/gala_b200.h but not exported"
        assert n in _abi.SIGNATURES, f"{n} has no ctypes signature in gala_b200/_abi.py"
    for n in _abi.SIGNATURES:
        assert n in names, f"{n} bound in _abi.py but not declared in the header"


def test_struct_layouts_match_header():
    # sizes implied by the header on LP64: gb_component 4*4 + 8 + 3*8 + 9*8 = 120
    assert ctypes.sizeof(_abi.gb_component) == 120
    assert ctypes.sizeof(_abi.gb_potential) == 16
    assert ctypes.sizeof(_abi.gb_frame) == 32
    assert ctypes.sizeof(_abi.gb_launch) == 40
    assert ctypes.sizeof(_abi.gb_dop853_stats) == 32


def test_version_and_counters():
    L = _abi.lib()
    assert b"sm_100a" in L.gb_version()
    assert _abi.launch_count() >= 0


@pytest.mark.skipif(_abi.device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback():
    """Without a GPU every compute call must raise, never silently compute on the CPU."""
    pot = gb.MilkyWayPotential2022()
    H = gb.Hamiltonian(pot)
    q = np.ones((3, 4))
    w0 = np.ones((6, 4))
    t = np.arange(4.0)
    for call in (lambda: pot.gradient(q), lambda: pot.energy(q), lambda: pot.density(q), lambda: H.energy(w0),
                 lambda: gb.leapfrog_integrate_hamiltonian(H, w0, t),
                 lambda: gb.ruth4_integrate_hamiltonian(H, w0, t),
                 lambda: gb.dop853_integrate_hamiltonian(H, w0, t),
                 lambda: gb.DirectNBody(w0[:, :2], [gb.PlummerPotential(m=1e9, b=0.1), None],
                                        external_potential=pot).integrate_orbit(t=t, Integrator="leapfrog"),
                 lambda: gb.DirectNBody(w0[:, :2], [gb.PlummerPotential(m=1e9, b=0.1), None],
                                        external_potential=pot).integrate_orbit(t=t, Integrator="dopri853"),
                 lambda: gb.StreaklineStreamDF()._sample(pot, np.ones((4, 3)), np.ones((4, 3)), t, np.ones(4),
                                                         np.ones(4, dtype="i4"))):
        with pytest.raises(_abi.GalaB200Error, match="no CUDA device"):
            call()


def test_nbody_host_validation():
    """Argument checks that happen before any device work (DirectNBody / _BodySpec)."""
    from gala_b200.mockstream import _BodySpec
    assert ctypes.sizeof(_abi.gb_bodies) == 16
    pp = gb.PlummerPotential(m=1e9, b=0.1)
    bs = _BodySpec([pp, None, pp], n_sources=2)          # third body is beyond the source count: massless
    assert bs.struct.n_bodies == 3 and bs.pots[0].n_components == 1 and bs.pots[1].n_components == 0
    assert bs.pots[2].n_components == 0
    with pytest.raises(NotImplementedError):
        _BodySpec([pp] * 17)
    with pytest.raises(ValueError):
        gb.DirectNBody(np.ones((6, 2)), [pp])
    nb = gb.DirectNBody(np.ones((6, 3)), [None, pp, None], external_potential=gb.MilkyWayPotential2022())
    assert nb.n_massive == 1
    with pytest.raises(TypeError):
        gb.MockStreamGenerator(gb.FardalStreamDF(), gb.Hamiltonian(gb.MilkyWayPotential2022()), progenitor_potential=3.0)


def test_multi_device_partition_rules():
    """Host-only arithmetic of a multi-device call (gb_launch.n_devices): contiguous orbit slices whose sizes differ by
    at most one and tile [0, N) exactly (== gala_b200.dist.shard_bounds), and the block-cyclic deal of mock-stream rows
    (groups of 128, group g to device g mod nd)."""
    from gala_b200.dist import shard_bounds
    L = _abi.lib()
    lo, n = ctypes.c_size_t(), ctypes.c_size_t()
    for N in (0, 1, 7, 8, 9, 1003, 10_000_000):
        for nd in (1, 2, 3, 8):
            got = []
            for k in range(nd):
                assert L.gb_shard_bounds(N, k, nd, ctypes.byref(lo), ctypes.byref(n)) == 0
                got.append((lo.value, lo.value + n.value))
            assert got == shard_bounds(N, nd)
            assert got[0][0] == 0 and got[-1][1] == N and all(a[1] == b[0] for a, b in zip(got, got[1:]))
    assert L.gb_shard_bounds(10, 3, 3, ctypes.byref(lo), ctypes.byref(n)) == -12
    for Np in (0, 1, 127, 128, 129, 1806, 100_020):
        for nd in (1, 2, 3, 8):
            counts = [L.gb_deal_count(Np, k, nd) for k in range(nd)]
            groups = [list(range(g * 128, min((g + 1) * 128, Np))) for g in range((Np + 127) // 128)]
            want = [sum(len(gr) for g, gr in enumerate(groups) if g % nd == k) for k in range(nd)]
            assert counts == want and sum(counts) == Np
    assert L.gb_deal_count(10, -1, 2) == -12
