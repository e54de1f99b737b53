"""`-m "not gpu"`: host-side mirror of the reference interface (no compute)."""
import numpy as np
import pytest

import gala_b200 as gb
from gala_b200 import _abi
from gala_b200.dist import shard_bounds


def test_parse_time_specification_matches_reference_semantics():
    # integrate/timespec.py:102,137-139: dt, n_steps -> t1 + cumsum, n_steps + 1 entries
    t = gb.parse_time_specification(None, dt=0.5, n_steps=4)
    assert np.array_equal(t, [0.0, 0.5, 1.0, 1.5, 2.0])
    t = gb.parse_time_specification(None, dt=-1.0, n_steps=3, t1=10.0)
    assert np.array_equal(t, [10.0, 9.0, 8.0, 7.0])
    t = gb.parse_time_specification(None, dt=0.25, t1=0.0, t2=1.0)
    assert np.array_equal(t, [0.0, 0.25, 0.5, 0.75])          # forward: t2 not appended (timespec.py:121-129)
    t = gb.parse_time_specification(None, dt=-0.25, t1=1.0, t2=0.0)
    assert np.array_equal(t, [1.0, 0.75, 0.5, 0.25, 0.0])     # backward: t2 appended (:109-119)
    t = gb.parse_time_specification(None, n_steps=5, t1=0.0, t2=1.0)
    assert np.array_equal(t, np.linspace(0, 1, 5))
    t = gb.parse_time_specification(None, t=np.array([0, 1, 3]))
    assert t.dtype == np.float64 and np.array_equal(t, [0.0, 1.0, 3.0])
    t = gb.parse_time_specification(None, dt=np.array([1.0, 2.0, 3.0]), t1=5.0)
    assert np.array_equal(t, [5.0, 6.0, 8.0])
    for bad in (dict(), dict(dt=1.0), dict(dt=1.0, t1=0.0, t2=-1.0), dict(dt=0.0, t1=0.0, t2=1.0)):
        with pytest.raises(ValueError):
            gb.parse_time_specification(None, **bad)


def test_mw2022_parameter_layout():
    """MN3 precompute (builtin/core.py:639-666) and component order
    (builtin/special.py:127-153); [G] + c_parameters is CPotentialWrapper._params."""
    pot = gb.MilkyWayPotential2022()
    assert list(pot.keys()) == ["disk", "bulge", "nucleus", "halo"]
    d = pot["disk"].c_parameters
    assert np.allclose(d[[0, 3, 6]], [7872306998.700792, -275625221944.33154, 320618418897.9487], rtol=1e-12)
    assert np.allclose(d[[1, 4, 7]], [1.5259431976529216, 6.782764436261113, 5.894799616164217], rtol=1e-12)
    assert np.allclose(d[[2, 5, 8]], 0.20663742603550295, rtol=1e-12)
    assert np.array_equal(d[9:], [4.7717e10, 2.6, 0.3])
    s = pot.spec()
    assert s.n == 4
    assert [s.comps[i].type_id for i in range(4)] == [_abi.POT_MN3, _abi.POT_HERNQUIST, _abi.POT_HERNQUIST,
                                                      _abi.POT_NFW_SPHERICAL]
    assert [s.comps[i].n_params for i in range(4)] == [13, 3, 3, 6]
    assert s.comps[0].params[0] == gb.G_GALACTIC
    assert all(s.comps[i].do_shift_rotate == 0 for i in range(4))
    assert np.array_equal([s.comps[3].params[k] for k in range(6)], [gb.G_GALACTIC, 5.5427e11, 15.626, 1, 1, 1])


def test_nfw_wrapper_choice_and_shift_rotate_flag():
    assert gb.NFWPotential(1e11, 12.0)._type_id == _abi.POT_NFW_SPHERICAL          # builtin/core.py:722-741
    assert gb.NFWPotential(1e11, 12.0, c=0.8)._type_id == _abi.POT_NFW_FLATTENED
    assert gb.NFWPotential(1e11, 12.0, b=0.9)._type_id == _abi.POT_NFW_TRIAXIAL
    p = gb.HernquistPotential(1e10, 1.0, origin=[1.0, 0, 0])
    assert p.spec().comps[0].do_shift_rotate == 1                                   # cpotential.pyx:79-92
    R = np.array([[0., 1, 0], [-1, 0, 0], [0, 0, 1]])
    p = gb.HernquistPotential(1e10, 1.0, R=R)
    c = p.spec().comps[0]
    assert c.do_shift_rotate == 1 and [c.R[k] for k in range(9)] == list(R.ravel())
    assert gb.HernquistPotential(1e10, 1.0).spec().comps[0].do_shift_rotate == 0
    m = gb.NFWPotential.from_circular_velocity(0.2, 15.0).parameters["m"]
    uu = 1.0
    assert np.isclose(m, 0.2 ** 2 / uu / (np.log(2) - 0.5) * 15.0 / gb.G_GALACTIC)


def test_composite_rules():
    c = gb.CCompositePotential()
    c["a"] = gb.HernquistPotential(1e10, 1.0)
    c["b"] = gb.NFWPotential(1e11, 12.0)
    assert c.spec().n == 2
    with pytest.raises(TypeError):
        c["c"] = c
    mw = gb.MilkyWayPotential2022()
    with pytest.raises(ValueError):
        mw["x"] = gb.HernquistPotential(1e10, 1.0)        # locked (special.py:271)
    s = gb.HernquistPotential(1e10, 1.0) + gb.NFWPotential(1e11, 12.0)
    assert isinstance(s, gb.CCompositePotential) and len(s) == 2


def test_scf_parameter_vector():
    S = np.zeros((3, 2, 2)); S[0, 0, 0] = 1.0
    p = gb.SCFPotential(m=1e12, r_s=20.0, Snlm=S)
    v = np.concatenate([[p.G], p.c_parameters])
    assert v[1] == 2 and v[2] == 1 and v[3] == 1e12 and v[4] == 20.0       # scf/bfe.cpp:229-258
    assert v.size == 5 + 2 * 12 and v[5] == 1.0


def test_hamiltonian_dispatch_and_errors():
    pot = gb.NFWPotential(1e11, 12.0)
    H = gb.Hamiltonian(pot)
    assert isinstance(H.frame, gb.StaticFrame) and H.c_enabled
    Hr = gb.Hamiltonian(pot, gb.ConstantRotatingFrame([0, 0, 0.03]))
    w0 = np.ones((6, 2))
    with pytest.raises(TypeError):                      # leapfrog.pyx:64-68
        gb.leapfrog_integrate_hamiltonian(Hr, w0, np.arange(3.0))
    with pytest.raises(TypeError):                      # ruth4.pyx:49-52
        gb.ruth4_integrate_hamiltonian(Hr, w0, np.arange(3.0))
    with pytest.raises(ValueError):
        gb.leapfrog_integrate_hamiltonian(H, np.ones((5, 2)), np.arange(3.0))
    with pytest.raises(ValueError):
        gb.ConstantRotatingFrame([0.03])
    with pytest.raises(ValueError):
        H.integrate_orbit(w0, Integrator="rk5", dt=1.0, n_steps=2)


def test_fardal_plan_order():
    """df.pyx:393-454: per timestep all trailing particles, then all leading; prog_m == 0 skipped."""
    df = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(0))
    idx, sign = df._plan(np.array([1.0, 0.0, 1.0]), np.array([2, 5, 1]))
    assert list(idx) == [0, 0, 0, 0, 2, 2] and list(sign) == [1, 1, -1, -1, 1, -1]
    df = gb.FardalStreamDF(gala_modified=True, lead=False)
    idx, sign = df._plan(np.ones(2), np.array([1, 2]))
    assert list(idx) == [0, 1, 1] and list(sign) == [1, 1, 1]
    with pytest.raises(ValueError):
        gb.FardalStreamDF(lead=False, trail=False)


def test_rng_broadcast_draw_equals_scalar_draws():
    """One broadcast normal(loc[Np,4], scale[Np,4]) consumes the RNG stream exactly like the
    reference's 4 scalar draws per particle (df.pyx:419-425), for both RNG flavours."""
    loc = np.broadcast_to([2.0, 0.0, 0.3, 0.0], (50, 4)); scale = np.broadcast_to([0.5, 0.5, 0.5, 0.5], (50, 4))
    for mk in (lambda: np.random.RandomState(42), lambda: np.random.default_rng(42)):
        a = mk().normal(loc, scale)
        r = mk()
        b = np.array([[r.normal(loc[i, k], scale[i, k]) for k in range(4)] for i in range(50)])
        assert np.array_equal(a, b)
        # what FardalStreamDF._draws does: standard_normal, then scale and shift in place
        df = gb.FardalStreamDF(gala_modified=True, random_state=mk())
        assert np.array_equal(df._draws(50, None), b)


def test_chen_batched_draw_matches_per_particle_calls():
    """ChenStreamDF: one multivariate_normal(mean, cov, size=Np) consumes the RNG like the reference's
    per-particle calls (df.pyx:632, 663) and agrees with them to 1 ulp; LagrangeCloud: one broadcast
    normal(0, v_disp) call == three scalar draws per particle (df.pyx:520-522)."""
    for mk in (lambda: np.random.RandomState(7), lambda: np.random.default_rng(7)):
        df = gb.ChenStreamDF(random_state=mk())
        a = df._draws(500, None)
        r = mk()
        b = np.array([r.multivariate_normal(gb.ChenStreamDF.mean, gb.ChenStreamDF.cov) for _ in range(500)])
        assert np.max(np.abs(a - b)) < 1e-14
        assert np.all(a[:, 3] == 1.0)                      # cov[3,3] = 0: v is exactly the mean
        df = gb.LagrangeCloudStreamDF(v_disp=0.002, random_state=mk())
        a = df._draws(100, None)
        r = mk()
        b = np.array([[r.normal(0, 0.002) for _ in range(3)] for _ in range(100)])
        assert np.array_equal(a, b)
    assert gb.StreaklineStreamDF()._draws(10, None) is None


def test_shard_bounds_and_work_dealing():
    b = shard_bounds(10, 4)
    assert b == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]


def test_gala_plugin_extracts_duck_typed_gala_objects():
    """gala_plugin maps real gala objects to the C-ABI spec through exactly what the Cython code reads
    (type(pot.c_instance).__name__, pot.G, pot.c_parameters, pot.origin, pot._R; frame.c_parameters;
    cpotential.pyx:281-316, ccompositepotential.pyx:27-86, cframe.pyx:98-112).  gala itself is not importable
    here, so stand-ins with the same attribute names are used."""
    from collections import OrderedDict
    from gala_b200 import gala_plugin as gp, _abi

    def fake(wrapper_name, G, c_parameters, origin=(0, 0, 0), R=None):
        W = type(wrapper_name, (), {})
        return type("FakePot", (), {"c_instance": W(), "G": G, "c_parameters": np.asarray(c_parameters, float),
                                    "origin": np.asarray(origin, float), "_R": R, "units": None})()

    G = gb.G_GALACTIC
    h = gp.extract_potential(fake("HernquistWrapper", G, [1e11, 0.5], origin=(1, 2, 3)))
    mine = gb.HernquistPotential(m=1e11, c=0.5, origin=[1, 2, 3])
    c0, c1 = h._components()[0], mine._components()[0]
    assert c0[0] == c1[0] == _abi.POT_HERNQUIST and np.array_equal(c0[1], c1[1]) and np.array_equal(c0[2], c1[2])

    class FakeComposite(OrderedDict):
        c_instance = type("CCompositePotentialWrapper", (), {})()
    comp = FakeComposite()
    comp["disk"] = fake("MiyamotoNagaiWrapper", G, [6.8e10, 3.0, 0.28])
    comp["halo"] = fake("LogarithmicWrapper", G, [0.2, 12.0, 1.38, 1.0, 1.36, 1.7])
    out = gp.extract_potential(comp)
    assert list(out.keys()) == ["disk", "halo"]
    assert [c[0] for c in out._components()] == [_abi.POT_MIYAMOTONAGAI, _abi.POT_LOGARITHMIC]
    with pytest.raises(TypeError):
        gp.extract_potential(fake("CylSplineWrapper", G, [1.0]))
    # every wrapper id of the header has an entry
    assert set(gp.WRAPPER_TO_TYPE.values()) == set(range(21))

    StaticW = type("StaticFrameWrapper", (), {})
    RotW = type("ConstantRotatingFrameWrapper3D", (), {})
    fs = type("F", (), {"c_instance": StaticW(), "c_parameters": np.zeros(0)})()
    frw = type("F", (), {"c_instance": RotW(), "c_parameters": np.array([0.0, 0.0, 0.03])})()
    assert isinstance(gp.extract_frame(fs), gb.StaticFrame)
    rot = gp.extract_frame(frw)
    assert isinstance(rot, gb.ConstantRotatingFrame) and np.allclose(rot.spec().omega[:], [0, 0, 0.03])
    Hf = type("H", (), {"potential": comp, "frame": frw})()
    H = gp.extract_hamiltonian(Hf)
    assert isinstance(H, gb.Hamiltonian) and H.c_enabled


def test_gala_plugin_adapts_nbody():
    from gala_b200 import gala_plugin as gp
    G = gb.G_GALACTIC
    W = type("PlummerWrapper", (), {})
    N = type("NullWrapper", (), {})
    pp = type("P", (), {"c_instance": W(), "G": G, "c_parameters": np.array([1e9, 0.1]), "origin": np.zeros(3), "_R": None})()
    nullp = type("P", (), {"c_instance": N(), "G": G, "c_parameters": np.zeros(0), "origin": np.zeros(3), "_R": None})()
    Hf = type("H", (), {"potential": gb.MilkyWayPotential2022(), "frame": gb.StaticFrame()})()
    nb = type("NB", (), {"_c_w0": np.arange(12.0).reshape(2, 6), "particle_potentials": [pp, nullp], "H": Hf})()
    ours = gp.adapt_nbody(nb)
    assert ours.n_massive == 1 and ours._c_w0.shape == (2, 6) and np.array_equal(ours._c_w0, nb._c_w0)
    assert ours.particle_potentials[1] is None


def test_plugin_install_patches_the_names_gala_resolves(tmp_path, monkeypatch):
    """A fake ``gala`` tree with the reference's import structure: the package ``gala.integrate.cyintegrators``
    binds the three functions in its ``__init__`` (integrate/cyintegrators/__init__.py:1-3) and
    ``integrate_orbit`` resolves them with ``from ...integrate.cyintegrators import X`` at call time
    (chamiltonian.pyx:324-336); ``mockstream_generator`` imports its three functions from ``._mockstream`` at import
    time (mockstream_generator.py:8-12).  After ``install()`` both lookups must find the GPU-backed wrappers."""
    import importlib
    import sys
    import textwrap
    root = tmp_path / "site"
    files = {
        "gala/__init__.py": "",
        "gala/integrate/__init__.py": "",
        "gala/integrate/cyintegrators/__init__.py": """
            from .dop853 import dop853_integrate_hamiltonian
            from .leapfrog import leapfrog_integrate_hamiltonian
            from .ruth4 import ruth4_integrate_hamiltonian
        """,
        "gala/integrate/cyintegrators/leapfrog.py": "def leapfrog_integrate_hamiltonian(*a, **k):\n    return 'cython leapfrog'\n",
        "gala/integrate/cyintegrators/ruth4.py": "def ruth4_integrate_hamiltonian(*a, **k):\n    return 'cython ruth4'\n",
        "gala/integrate/cyintegrators/dop853.py": "def dop853_integrate_hamiltonian(*a, **k):\n    return 'cython dop853'\n",
        "gala/potential/__init__.py": "",
        "gala/potential/hamiltonian/__init__.py": "",
        "gala/potential/hamiltonian/chamiltonian.py": """
            def integrate_orbit_lookup(which):
                # the reference's late import, verbatim in structure (chamiltonian.pyx:324-336)
                if which == 'leapfrog':
                    from ...integrate.cyintegrators import leapfrog_integrate_hamiltonian as f
                elif which == 'ruth4':
                    from ...integrate.cyintegrators import ruth4_integrate_hamiltonian as f
                else:
                    from ...integrate.cyintegrators import dop853_integrate_hamiltonian as f
                return f
        """,
        "gala/dynamics/__init__.py": "",
        "gala/dynamics/mockstream/__init__.py": """
            from ._mockstream import mockstream_dop853
            from .mockstream_generator import *
        """,
        "gala/dynamics/mockstream/_mockstream.py": """
            def mockstream_dop853(*a, **k): return 'cython'
            def mockstream_leapfrog(*a, **k): return 'cython'
            def mockstream_dop853_animate(*a, **k): return 'cython'
        """,
        "gala/dynamics/mockstream/mockstream_generator.py": """
            from ._mockstream import (mockstream_dop853, mockstream_dop853_animate, mockstream_leapfrog)
            __all__ = ['run_lookup']
            def run_lookup(which):
                return {'dop853': mockstream_dop853, 'leapfrog': mockstream_leapfrog,
                        'animate': mockstream_dop853_animate}[which]
        """,
    }
    for rel, src in files.items():
        f = root / rel
        f.parent.mkdir(parents=True, exist_ok=True)
        f.write_text(textwrap.dedent(src))
    monkeypatch.syspath_prepend(str(root))
    for m in [k for k in sys.modules if k == "gala" or k.startswith("gala.")]:
        monkeypatch.delitem(sys.modules, m)
    from gala_b200 import gala_plugin as gp
    try:
        ch = importlib.import_module("gala.potential.hamiltonian.chamiltonian")
        gen = importlib.import_module("gala.dynamics.mockstream.mockstream_generator")
        assert ch.integrate_orbit_lookup("leapfrog")() == "cython leapfrog"
        hooks = gp.install()
        for which in ("leapfrog", "ruth4", "dop853"):
            f = ch.integrate_orbit_lookup(which)
            assert f.__name__ == f"{which}_integrate_hamiltonian" and f.__module__ == gp.__name__, which
        for which in ("dop853", "leapfrog", "animate"):
            assert gen.run_lookup(which).__module__ == gp.__name__, which
        import gala.dynamics.mockstream as msp
        assert msp.mockstream_dop853.__module__ == gp.__name__
        assert "gala.integrate.cyintegrators.leapfrog_integrate_hamiltonian" in hooks
        assert "gala.dynamics.mockstream._mockstream.mockstream_leapfrog" in hooks and len(hooks) == 13
        # a gala without the mock-stream extension: loud, not silent
        (root / "gala/dynamics/mockstream/mockstream_generator.py").unlink()
        (root / "gala/dynamics/mockstream/__init__.py").write_text("")
        for m in [k for k in sys.modules if k.startswith("gala.dynamics")]:
            del sys.modules[m]
        importlib.invalidate_caches()
        with pytest.warns(RuntimeWarning, match="not patched"):
            with pytest.raises(RuntimeError, match="could not be patched"):
                gp.install()
        with pytest.warns(RuntimeWarning):
            assert len(gp.install(strict=False)) == 10
    finally:
        for m in [k for k in sys.modules if k == "gala" or k.startswith("gala.")]:
            del sys.modules[m]


# -- Orbit analysis helpers that are plain array reductions (no kernel involved): the reference's own tests -----------
def test_estimate_period_like_the_reference():
    """tests/dynamics/test_orbit.py:449-469: R(t) = 1 + 0.25 sin(2 pi t / T_R), phi = 2 pi t."""
    import gala_b200 as gb
    ntimes = 16384
    for true_T_R in (1.0, 2.0, 4.123):
        t = np.linspace(0, 10.0, ntimes)
        R = 0.25 * np.sin(2 * np.pi / true_T_R * t) + 1.0
        phi = (2 * np.pi * t) % (2 * np.pi)
        pos = np.zeros((3, ntimes))
        pos[0] = R * np.cos(phi)
        pos[1] = R * np.sin(phi)
        orb = gb.Orbit(pos, np.zeros_like(pos), t=t)
        T = orb.estimate_period()
        assert set(T) == {"x", "y", "z"} and T["x"].shape == (1,)
        T = orb.estimate_period(components=("rho", "phi"))
        assert np.allclose(T["rho"], true_T_R, rtol=1e-3)
        assert np.allclose(T["phi"], 1.0, rtol=1e-3)
    with pytest.raises(ValueError):
        gb.Orbit(pos, np.zeros_like(pos)).estimate_period()


def test_align_circulation_like_the_reference():
    """tests/dynamics/test_orbit.py:525-574: loops about x, y, z and a box."""
    import gala_b200 as gb
    t = np.linspace(0, 100, 1024)
    w = np.zeros((6, 1024, 4))
    w[1, :, 0] = np.cos(t); w[2, :, 0] = np.sin(t); w[4, :, 0] = -np.sin(t); w[5, :, 0] = np.cos(t)
    w[0, :, 1] = -np.cos(t); w[2, :, 1] = np.sin(t); w[3, :, 1] = np.sin(t); w[5, :, 1] = np.cos(t)
    w[0, :, 2] = np.cos(t); w[1, :, 2] = np.sin(t); w[3, :, 2] = -np.sin(t); w[4, :, 2] = np.cos(t)
    w[0, :, 3] = np.cos(t); w[1, :, 3] = -np.cos(0.5 * t); w[2, :, 3] = np.cos(0.25 * t)
    w[3, :, 3] = -np.sin(t); w[4, :, 3] = 0.5 * np.sin(0.5 * t); w[5, :, 3] = -0.25 * np.sin(0.25 * t)
    for i in range(4):
        orb = gb.Orbit.from_w(w[..., i], t=t)
        assert orb.circulation().shape == (3,)
        circ = orb.align_circulation_with_z().circulation()
        if i == 3:
            assert circ.sum() == 0
        else:
            assert circ[2] == 1
    orb = gb.Orbit.from_w(w, t=t)
    circ = orb.circulation()
    assert circ.shape == (3, 4)
    assert np.array_equal(circ, np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]))
    new_orb = orb.align_circulation_with_z()
    new_circ = new_orb.circulation()
    assert np.all(new_circ[2, :3] == 1) and np.all(new_circ[:, 3] == 0)
    # the exchanged axes carry the other one's samples, the box is untouched
    assert np.array_equal(new_orb.pos[2, :, 0], w[0, :, 0]) and np.array_equal(new_orb.pos[0, :, 0], w[2, :, 0])
    assert np.array_equal(new_orb.pos[:, :, 3], w[:3, :, 3])
    with pytest.raises(ValueError):
        orb.align_circulation_with_z(circulation=np.ones((3, 2), dtype=int))
    # PhaseSpacePosition point quantities on plain arrays
    L = gb.PhaseSpacePosition.from_w(w[:, 5]).angular_momentum()
    assert L.shape == (3, 4) and np.allclose(L[0, 0], 1.0) and np.allclose(L[2, 2], 1.0)
    assert np.allclose(gb.PhaseSpacePosition.from_w(w[:, 5]).kinetic_energy()[:3], 0.5)


def test_guiding_radius_batched_secant_against_brentq():
    """tests/dynamics/test_dynamics_core.py:378-394 (Hernquist m = 1e11, c = 10; R in [4, 10] kpc; v_y ~ N(v_c, 15 km/s)):
    the batched secant iteration of PhaseSpacePosition.guiding_radius against one scipy brentq per point, with the
    closed-form circular velocity standing in for the device call (the GPU version of this test is in
    test_gpu_point_quantities.py)."""
    from scipy.optimize import brentq
    import gala_b200 as gb
    m, c, G = 1e11, 10.0, gb.G_GALACTIC

    class Hern:
        calls = 0

        def circular_velocity(self, q, t=0.0):
            Hern.calls += 1
            r = np.sqrt((q * q).sum(0))
            return np.sqrt(G * m * r) / (r + c)

    rng = np.random.default_rng(42)
    R = rng.uniform(4, 10, 128)
    xyz = R[None] * np.array([1.0, 0, 0])[:, None]
    vc = Hern().circular_velocity(xyz)
    vxyz = np.zeros((3, R.size))
    vxyz[1] = rng.normal(vc / gb.KMS_TO_KPC_MYR, 15.0) * gb.KMS_TO_KPC_MYR
    w0 = gb.PhaseSpacePosition(xyz, vxyz)
    Hern.calls = 0
    Rg = w0.guiding_radius(Hern())
    assert Rg.shape == (128,) and np.all(Rg > 0) and np.all(Rg < 25)
    assert Hern.calls < 20                                        # a handful of BATCHED evaluations, not 128 root solves
    Lz = np.abs(R * vxyz[1])
    want = np.array([brentq(lambda x: L - x * np.sqrt(G * m * x) / (x + c), 1e-3, 1e3, xtol=1e-14, rtol=1e-14) for L in Lz])
    assert np.allclose(Rg, want, rtol=1e-10)
    # a point with L_z = 0 has no guiding radius other than 0: NaN or ~0, never an exception
    z = gb.PhaseSpacePosition([[5.0], [0.0], [0.0]], [[0.1], [0.0], [0.0]]).guiding_radius(Hern())
    assert z.shape == (1,) and (np.isnan(z[0]) or z[0] < 1e-6)


def test_frame_transform_like_the_reference():
    """tests/dynamics/test_orbit.py:576-603 + the arithmetic of potential/frame/builtin/transformations.py:98-150
    against scipy's Rotation: static -> rotating turns positions and velocities by -Omega t."""
    from scipy.spatial.transform import Rotation
    import gala_b200 as gb
    static = gb.StaticFrame()
    Om = np.array([0.53, 1.241, 0.9394])
    rotating = gb.ConstantRotatingFrame(Omega=Om)
    rng = np.random.default_rng(5)
    x = rng.random((3, 10)); v = rng.random((3, 10))
    t = np.linspace(0, 1, 10)
    o = gb.Orbit(pos=x, vel=v, t=t)
    with pytest.raises(ValueError):
        o.to_frame(rotating)                                      # no frame specified at init
    a = o.to_frame(rotating, current_frame=static, t=o.t)
    b = o.to_frame(rotating, current_frame=static)
    assert np.array_equal(a.pos, b.pos) and a.frame is rotating and isinstance(b, gb.Orbit)
    o = gb.Orbit(pos=x, vel=v, t=t, frame=static)
    assert np.array_equal(o.to_frame(rotating).pos, a.pos) and np.array_equal(o.to_frame(rotating, t=o.t).vel, a.vel)
    assert o.to_frame(static) is o
    for j in range(10):
        Rm = Rotation.from_rotvec(-Om * t[j]).as_matrix()
        assert np.allclose(a.pos[:, j], Rm @ x[:, j], atol=1e-15) and np.allclose(a.vel[:, j], Rm @ v[:, j], atol=1e-15)
    back = a.to_frame(static)
    assert np.allclose(back.pos, x, atol=1e-15) and np.allclose(back.vel, v, atol=1e-15)
    # several orbits: the angle follows the time axis; a PhaseSpacePosition needs t
    x3 = rng.random((3, 10, 4)); v3 = rng.random((3, 10, 4))
    o3 = gb.Orbit(pos=x3, vel=v3, t=t, frame=static).to_frame(rotating)
    for n in range(4):
        assert np.allclose(o3.pos[:, :, n], gb.Orbit(pos=x3[:, :, n], vel=v3[:, :, n], t=t, frame=static).to_frame(rotating).pos)
    p = gb.PhaseSpacePosition(x, v, frame=static)
    with pytest.raises(ValueError):
        p.to_frame(rotating)
    assert np.allclose(p.to_frame(rotating, t=t).pos, a.pos) and np.allclose(p.to_frame(rotating, t=0.5).pos[:, 0],
                                                                            Rotation.from_rotvec(-Om * 0.5).as_matrix() @ x[:, 0])


def test_combine_like_the_reference():
    """tests/dynamics/test_dynamics_util.py:51-122 (TestCombine), without the unit bookkeeping."""
    import gala_b200 as gb
    rng = np.random.default_rng(8)
    static = gb.StaticFrame()
    x, v = rng.random(3), rng.random(3)
    p1 = gb.PhaseSpacePosition(pos=x, vel=v)
    p2 = gb.PhaseSpacePosition(pos=x, vel=v, frame=static)
    x, v = rng.random((3, 5)), rng.random((3, 5))
    p3 = gb.PhaseSpacePosition(pos=x, vel=v)
    x, v = rng.random((2, 5)), rng.random((2, 5))
    p5 = gb.PhaseSpacePosition(pos=x, vel=v)
    psps = [p1, p2, p3, p5]
    x, v = rng.random((3, 8)), rng.random((3, 8))
    o1 = gb.Orbit(pos=x, vel=v)
    o2 = gb.Orbit(pos=x, vel=v, t=np.arange(8.0))
    o3 = gb.Orbit(pos=x, vel=v, t=np.arange(8.0), frame=static)
    x3, v3 = rng.random((3, 8, 2)), rng.random((3, 8, 2))
    o5 = gb.Orbit(pos=x3, vel=v3, t=np.arange(8.0))
    orbs = [o1, o2, o3, o5]
    for bad, exc in (([], ValueError), (p1, ValueError), ([p1, o1], TypeError), ([5, 5, 5], TypeError), (psps, ValueError),
                     (orbs, ValueError), ([o2, gb.Orbit(pos=x, vel=v, t=np.arange(8.0) + 1e-3)], ValueError)):
        with pytest.raises(exc):
            gb.combine(bad)
    assert gb.combine([p3]) is p3
    for psp in psps:
        new = gb.combine([psp] * 3)
        n = psp.pos.shape[1] if psp.pos.ndim > 1 else 1
        assert new.ndim == psp.ndim and new.pos.shape == (psp.ndim, 3 * n) and new.frame == psp.frame
    for orb in orbs:
        new = gb.combine([orb] * 4)
        assert new.pos.shape == (3, 8, 4 * orb.norbits) and new.frame == orb.frame and new.hamiltonian is orb.hamiltonian
        assert np.array_equal(new.pos[:, :, -1], orb.pos.reshape(3, 8, -1)[:, :, -1])
    w0 = gb.combine((p3, gb.PhaseSpacePosition(pos=p3.pos + 1.0, vel=p3.vel)))     # tests/dynamics/nbody/test_nbody.py:38
    assert w0.w().shape == (6, 10) and np.array_equal(w0.pos[:, 5:], p3.pos + 1.0)


def test_gala_plugin_extracts_time_interpolated_potentials():
    """A gala TimeInterpolatedPotential seen through its public attributes (time_interpolated.py:59-260) becomes the
    same device parameter vector as the native class built from the same inputs -- incl. a class with derived C
    parameters (MN3), a moving origin and rotation matrices.  Stand-ins with gala's attribute names are used."""
    from gala_b200 import gala_plugin as gp
    T = np.linspace(0.0, 400.0, 9)
    grow = 1.0 + 0.3 * T / 400.0
    orb = np.stack([10 * np.cos(T / 100), 10 * np.sin(T / 100), 0 * T], axis=1)
    ang = 0.04 * T
    R = np.array([[[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]] for a in ang])

    def gala_like_class(native_cls, wrapper_name):
        W = type(wrapper_name, (), {})

        class K:                                     # what gala's potential_cls(units=..., **params) returns
            def __init__(self, units=None, **kw):
                self.c_parameters = native_cls(**kw).c_parameters
                self.c_instance = W()
        return K

    def gala_like_ti(native_cls, wrapper_name, method, origin=None, R=None, **params):
        names = [k for k in params]
        interp = [k for k in names if np.ndim(params[k]) >= 1]
        P = dict(potential_cls=gala_like_class(native_cls, wrapper_name), time_knots=T, interpolation_method=method, **params)
        return type("FakeTI", (), {"c_instance": type("TimeInterpolatedWrapper", (), {})(), "parameters": P,
                                   "_potential_param_names": names, "_interp_params": interp, "_extra_wrapped_kwargs": {},
                                   "origin": origin, "R": R, "G": gb.G_GALACTIC, "units": None})()

    cases = [
        (gb.HernquistPotential, "HernquistWrapper", "cspline", dict(m=5e10 * grow, c=1.5), dict(origin=orb)),
        (gb.MN3ExponentialDiskPotential, "MN3ExponentialDiskWrapper", "akima", dict(m=5e10 * grow, h_R=2.6, h_z=0.3), {}),
        (gb.LongMuraliBarPotential, "LongMuraliBarWrapper", "linear", dict(m=1e10, a=4.0, b=0.8, c=0.25, alpha=0.0), dict(R=R)),
    ]
    for native_cls, wname, method, params, where in cases:
        got = gp.extract_potential(gala_like_ti(native_cls, wname, method, **where, **params))
        want = gb.TimeInterpolatedPotential(native_cls, T, interpolation_method=method, **params, **where)
        assert isinstance(got, gb.TimeInterpolatedPotential)
        assert np.array_equal(got.c_parameters, want.c_parameters), wname
        assert got._components()[0][0] == want._components()[0][0] == gb._abi.POT_TIMEINTERP
        assert got.time_bounds == (0.0, 400.0)
    # inside a composite, next to a static component
    from collections import OrderedDict

    class FakeComposite(OrderedDict):
        c_instance = type("CCompositePotentialWrapper", (), {})()
    comp = FakeComposite()
    comp["halo"] = type("FakePot", (), {"c_instance": type("SphericalNFWWrapper", (), {})(), "G": gb.G_GALACTIC,
                                        "c_parameters": np.array([6e11, 16.0, 1.0, 1.0, 1.0]), "origin": np.zeros(3),
                                        "_R": None, "units": None})()
    comp["bar"] = gala_like_ti(*cases[2][:3], **cases[2][4], **cases[2][3])
    out = gp.extract_potential(comp)
    assert [c[0] for c in out._components()] == [gb._abi.POT_NFW_SPHERICAL, gb._abi.POT_TIMEINTERP]
    with pytest.raises(TypeError):
        gp.extract_potential(gala_like_ti(gb.HernquistPotential, "CylSplineWrapper", "linear", m=1e10, c=1.0))
