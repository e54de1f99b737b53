"""`-m gpu`: one HOST-mode C-ABI call sharded over several devices (gb_launch.n_devices, SURVEY 8e / north_star
"each GPU takes a contiguous slice of orbits and results are gathered to the host") must return the bits of the
single-device call: orbits never interact, so slicing may not change a single result.

Two layers: (1) on ANY box, the slicing / pitched-copy / block-cyclic-deal logic is exercised by listing device 0
several times (test hook GB_ALLOW_DUP_DEVICES: the slices then run one after the other on that device);
(2) with >= 2 GPUs the same comparisons run over really distinct devices.
Reference entry being replaced: potential/hamiltonian/chamiltonian.pyx:315-356 (one call, whole w0)."""
import os

import numpy as np
import pytest

import gala_b200 as gb
from gala_b200 import _abi
from conftest import make_ic

pytestmark = pytest.mark.gpu


def device_lists():
    out = [pytest.param([0, 0, 0], id="dev0x3-sliced")]
    n = _abi.device_count()
    if n >= 2:
        out.append(pytest.param(list(range(min(n, 8))), id=f"{min(n, 8)}gpus"))
    else:
        out.append(pytest.param(None, id="multi-gpu", marks=pytest.mark.skip(reason="needs >= 2 GPUs")))
    return out


class use_devices:
    def __init__(self, devs):
        self.devs = devs

    def __enter__(self):
        os.environ["GB_ALLOW_DUP_DEVICES"] = "1"
        _abi._devices = list(self.devs)

    def __exit__(self, *a):
        _abi._devices = None
        os.environ.pop("GB_ALLOW_DUP_DEVICES", None)


@pytest.fixture(scope="module")
def mw():
    return gb.Hamiltonian(gb.MilkyWayPotential2022())


@pytest.mark.parametrize("devs", device_lists())
@pytest.mark.parametrize("N,save_all", [(1003, 1), (1003, 0), (200_001, 0), (70_000, 1)])
def test_leapfrog_sharded_bit_identical(mw, devs, N, save_all):
    w0 = make_ic(lambda q: mw.potential.gradient(q), N, 11)
    t = np.arange(41 if N > 5000 else 201, dtype=float)
    _, ref = gb.leapfrog_integrate_hamiltonian(mw, w0, t, save_all=save_all)
    with use_devices(devs):
        _, got = gb.leapfrog_integrate_hamiltonian(mw, w0, t, save_all=save_all)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("devs", device_lists())
def test_ruth4_rotating_sharded_bit_identical(devs):
    pot = gb.CCompositePotential()
    pot["bar"] = gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=0.4)
    for k, v in gb.MilkyWayPotential2022().items():
        pot[k] = v
    H = gb.Hamiltonian(pot, gb.ConstantRotatingFrame([0.0, 0.0, 0.03]))
    w0 = make_ic(lambda q: pot.gradient(q), 777, 12)
    t = np.arange(101) * 0.5
    _, ref = gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=1, allow_rotating_frame=True)
    with use_devices(devs):
        _, got = gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=1, allow_rotating_frame=True)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("devs", device_lists())
@pytest.mark.parametrize("save_all", [1, 0])
def test_dop853_sharded_bit_identical(mw, devs, save_all):
    N = 4099          # > 2048 per call: the single-device call sorts its queue; slices below 2048 do not -- same bits
    w0 = make_ic(lambda q: mw.potential.gradient(q), N, 13)
    t = np.linspace(0.0, 300.0, 61)
    _, ref, sref = gb.dop853_integrate_hamiltonian(mw, w0, t, save_all=save_all, return_status=True)
    with use_devices(devs):
        _, got, sgot = gb.dop853_integrate_hamiltonian(mw, w0, t, save_all=save_all, return_status=True)
    assert np.array_equal(got, ref)
    for k in ("status", "nstep", "naccpt", "nrejct", "nfcn"):
        assert np.array_equal(sgot[k], sref[k]), k


@pytest.mark.parametrize("devs", device_lists())
def test_evaluation_sharded_bit_identical(mw, devs):
    rng = np.random.default_rng(3)
    q = rng.normal(0, 15, (3, 2501))
    w = np.vstack([q, rng.normal(0, 0.1, (3, 2501))])
    pot = mw.potential
    ref = [pot.gradient(q), pot.energy(q), pot.density(q), mw.energy(w), mw.gradient(w)]
    with use_devices(devs):
        got = [pot.gradient(q), pot.energy(q), pot.density(q), mw.energy(w), mw.gradient(w)]
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("devs", device_lists())
@pytest.mark.parametrize("integ", ["leapfrog", "dopri853"])
def test_mockstream_dealt_bit_identical(mw, devs, integ):
    """The particle rows are dealt in groups of 128 (capi.cu:Deal): 2 x 3 x 301 = 1806 particles = 14 full groups
    + a partial one, so every branch of the pitched copies runs."""
    prog = gb.PhaseSpacePosition(pos=[13.0, 0.0, 20.0], vel=np.array([0.0, 130.0, 50.0]) * gb.KMS_TO_KPC_MYR)

    def run():
        gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(7)), mw)
        stream, p = gen.run(prog, 2.5e4, dt=-1.0, n_steps=300, n_particles=3, release_every=1, Integrator=integ)
        return stream.w(), p.w()

    ref = run()
    with use_devices(devs):
        got = run()
    assert got[0].shape == (6, 1806)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


def test_device_list_validation(mw):
    w0 = make_ic(lambda q: mw.potential.gradient(q), 64, 1)
    t = np.arange(5.0)
    _abi._devices = [0, 0]
    try:
        with pytest.raises(ValueError, match="twice"):
            gb.leapfrog_integrate_hamiltonian(mw, w0, t)
        _abi._devices = [4096]
        with pytest.raises(ValueError, match="not a CUDA device"):
            gb.leapfrog_integrate_hamiltonian(mw, w0, t)
    finally:
        _abi._devices = None
    with pytest.raises(ValueError):
        gb.set_devices([0, 0])
    assert gb.set_devices(None) is None
    # a device list with device-memory buffers is refused (a device pointer belongs to one device), and entry points
    # that run on one device say so instead of ignoring the list
    import ctypes as C
    import torch
    wd = torch.as_tensor(w0, device="cuda")
    out = torch.empty_like(wd)
    td = torch.as_tensor(t, device="cuda")
    opt = _abi.launch_opts(False, devices=[0])
    opt.mem = _abi.MEM_DEVICE
    fr = mw.frame.spec()
    rc = _abi.lib().gb_leapfrog(mw.potential.spec().ptr(), C.byref(fr), wd.data_ptr(), 64, td.data_ptr(), 5, 0, out.data_ptr(), C.byref(opt))
    assert rc == -12 and b"GB_MEM_HOST" in _abi.lib().gb_last_error()
    opt = _abi.launch_opts(False, devices=[0])
    h = np.empty((9, 64))
    rc = _abi.lib().gb_hessian(mw.potential.spec().ptr(), np.ascontiguousarray(w0[:3]).ctypes.data, 0.0, 64, h.ctypes.data, C.byref(opt))
    assert rc == -12 and b"one device" in _abi.lib().gb_last_error()


@pytest.mark.parametrize("N", [0, 1, 2, 5])
def test_fewer_orbits_than_devices(mw, N):
    """Slices may be empty: N = 0, 1, 2 orbits over three (virtual) devices, and every output shape / value as on one."""
    w0 = make_ic(lambda q: mw.potential.gradient(q), max(N, 1), 3)[:, :N]
    w0 = np.ascontiguousarray(w0)
    t = np.arange(21.0)
    ref_lf = gb.leapfrog_integrate_hamiltonian(mw, w0, t, save_all=1)[1]
    ref_d8 = gb.dop853_integrate_hamiltonian(mw, w0, t, save_all=0, return_status=True)
    with use_devices([0, 0, 0]):
        got_lf = gb.leapfrog_integrate_hamiltonian(mw, w0, t, save_all=1)[1]
        got_d8 = gb.dop853_integrate_hamiltonian(mw, w0, t, save_all=0, return_status=True)
        st = gb.integrate_extrema(mw, w0, t, Integrator="ruth4", return_final=True)
    assert got_lf.shape == (6, 21, N) and np.array_equal(got_lf, ref_lf)
    assert np.array_equal(got_d8[1], ref_d8[1]) and np.array_equal(got_d8[2]["nstep"], ref_d8[2]["nstep"])
    assert st["w_final"].shape == (6, N) and st["n_peri"].shape == (N,)


def test_concurrent_callers_share_the_scratch_safely(mw):
    """host-buffer and device-buffer calls from several threads at once (staging slots, side streams, the
    cached coefficient tables and the DOP853 scratch are per device and locked / pinned per call): every thread gets
    the bits a lone caller gets."""
    import threading
    import torch
    S = np.zeros((3, 3, 3)); S[0, 0, 0] = 1.0; S[1, 0, 0] = 0.05; S[2, 2, 1] = 0.02
    scf = gb.SCFPotential(m=1e11, r_s=10.0, Snlm=S, Tnlm=np.zeros((3, 3, 3)))
    Hs = gb.Hamiltonian(scf)
    w0 = make_ic(lambda q: mw.potential.gradient(q), 70_000, 5)
    ws = np.ascontiguousarray(w0[:, :3000])
    t = np.arange(31.0)
    jobs = [
        lambda: gb.leapfrog_integrate_hamiltonian(mw, w0, t, save_all=0)[1],                 # chunk pipeline, side streams
        lambda: gb.dop853_integrate_hamiltonian(mw, ws, t, save_all=1)[1],                   # dense scratch, status slot
        lambda: gb.leapfrog_integrate_hamiltonian(Hs, ws, t, save_all=0)[1],                 # cached SCF table
        lambda: gb.dop853_integrate_hamiltonian(mw, torch.as_tensor(ws, device="cuda"), t, save_all=0)[1].cpu().numpy(),
    ]
    want = [j() for j in jobs]
    got = [[None] * 3 for _ in jobs]
    errs = []

    def worker(k):
        try:
            for rep in range(3):
                got[k][rep] = jobs[k]()
        except BaseException as e:      # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=worker, args=(k,)) for k in range(len(jobs))]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs
    for k in range(len(jobs)):
        for rep in range(3):
            assert np.array_equal(got[k][rep], want[k], equal_nan=True), (k, rep)


@pytest.mark.parametrize("devs", device_lists())
def test_time_interpolated_composite_sharded_bit_identical(devs):
    """A time-dependent composite: every device builds its own per-step state table (capi.cu: make_ti_table) and its
    own spline tables; the pipelined HOST path (N >= 65536) and the plain one must both return the single-device bits,
    for leapfrog, Ruth4 and DOPRI853."""
    T = np.linspace(0.0, 400.0, 9)
    grow = 1.0 + 0.3 * T / 400.0
    orb = np.stack([8 * np.cos(T / 150), 8 * np.sin(T / 150), 0.5 * np.sin(T / 90)], axis=1)
    pot = gb.CCompositePotential()
    pot["halo"] = gb.NFWPotential(m=6e11, r_s=16.0)
    pot["sat"] = gb.TimeInterpolatedPotential(gb.PlummerPotential, T, m=2e10 * grow, b=1.0, origin=orb)
    H = gb.Hamiltonian(pot)
    t = np.linspace(0.0, 400.0, 81)
    for N in (1003, 70_001):
        w0 = make_ic(lambda q: gb.MilkyWayPotential2022().gradient(q), N, 13)
        one = [gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0)[1], gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=0)[1],
               gb.dop853_integrate_hamiltonian(H, w0[:, :2000], t[::8].copy(), save_all=0)[1]]
        with use_devices(devs):
            many = [gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0)[1], gb.ruth4_integrate_hamiltonian(H, w0, t, save_all=0)[1],
                    gb.dop853_integrate_hamiltonian(H, w0[:, :2000], t[::8].copy(), save_all=0)[1]]
        for a, b in zip(one, many):
            assert np.isfinite(a).all() and np.array_equal(a, b)
