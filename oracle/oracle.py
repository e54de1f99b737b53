"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes front-end of the CPU checkers built by ``oracle/Makefile``:

* ``Ref("strict")``  -> ``oracle/_ref/libgala_ref.so``: the reference's own C++ compiled unmodified
  with strict-IEEE flags + the restated Cython loops of ``oracle/ref_driver.cpp``.  THE parity oracle.
* ``Ref("fast")``    -> ``oracle/_ref/libgala_ref_fast.so``: same sources with the reference's shipped
  flags (``-Ofast -march=x86-64-v3``); used for CPU timing and for reference-vs-reference floors.
* ``Port()`` / ``Port(long_double=True)`` -> ``oracle/_ref/libgala_port*.so``: our plain-C restatement
  (``oracle/port.c``), validated against ``Ref`` in tests/test_oracle_cpu.py.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may import this module.  Potentials / frames are passed as gala_b200 host objects; only their
``spec()`` (the C-ABI structs) is used.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")


def build(quiet=True):
    """Run oracle/Makefile (builds the port always, the reference libs when /root/reference exists)."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return r.stdout


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class _Lib:
    prefix = "ref_"

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built; run `make -C oracle`")
        self.path = path
        self.L = C.CDLL(path)

    def _fn(self, name, restype=C.c_int):
        f = getattr(self.L, self.prefix + name)
        f.restype = restype
        return f

    # ---- potential evaluation ----------------------------------------------------------------
    def gradient(self, pot, q, t=0.0):
        q = _f64(q); N = q.shape[1]
        out = np.empty((3, N))
        rc = self._fn("gradient")(pot.spec().ptr(), q.ctypes.data_as(C.c_void_p), C.c_double(t), C.c_size_t(N),
                                  out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def _scalar(self, name, pot, q, t):
        q = _f64(q); N = q.shape[1]
        out = np.empty(N)
        rc = self._fn(name)(pot.spec().ptr(), q.ctypes.data_as(C.c_void_p), C.c_double(t), C.c_size_t(N),
                            out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def energy(self, pot, q, t=0.0):
        return self._scalar("energy", pot, q, t)

    def density(self, pot, q, t=0.0):
        return self._scalar("density", pot, q, t)

    def hessian(self, pot, q, t=0.0):
        q = _f64(q); N = q.shape[1]
        out = np.empty((3, 3, N))
        rc = self._fn("hessian")(pot.spec().ptr(), q.ctypes.data_as(C.c_void_p), C.c_double(t), C.c_size_t(N),
                                 out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def hamiltonian_energy(self, H, w, t=0.0):
        w = _f64(w); N = w.shape[1]
        out = np.empty(N)
        fr = H.frame.spec()
        rc = self._fn("hamiltonian_energy")(H.potential.spec().ptr(), C.byref(fr), w.ctypes.data_as(C.c_void_p),
                                            C.c_double(t), C.c_size_t(N), out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def hamiltonian_gradient(self, H, w, t=0.0):
        w = _f64(w); N = w.shape[1]
        out = np.empty((6, N))
        fr = H.frame.spec()
        rc = self._fn("hamiltonian_gradient")(H.potential.spec().ptr(), C.byref(fr), w.ctypes.data_as(C.c_void_p),
                                              C.c_double(t), C.c_size_t(N), out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    # ---- integrators ---------------------------------------------------------------------------
    def leapfrog(self, pot, w0, t, save_all=True):
        w0 = _f64(w0); t = _f64(t); N = w0.shape[1]
        out = np.empty((6, t.size, N) if save_all else (6, N))
        rc = self._fn("leapfrog")(pot.spec().ptr(), w0.ctypes.data_as(C.c_void_p), C.c_size_t(N),
                                  t.ctypes.data_as(C.c_void_p), C.c_int(t.size), C.c_int(int(save_all)),
                                  out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def ruth4(self, H, w0, t, save_all=True):
        w0 = _f64(w0); t = _f64(t); N = w0.shape[1]
        out = np.empty((6, t.size, N) if save_all else (6, N))
        fr = H.frame.spec()
        rc = self._fn("ruth4")(H.potential.spec().ptr(), C.byref(fr), w0.ctypes.data_as(C.c_void_p), C.c_size_t(N),
                               t.ctypes.data_as(C.c_void_p), C.c_int(t.size), C.c_int(int(save_all)),
                               out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def dop853(self, H, w0, t, atol=1e-10, rtol=1e-10, nmax=0, dt_max=0.0, nstiff=0, save_all=True, nbatch=1):
        w0 = _f64(w0); t = _f64(t); N = w0.shape[1]
        out = np.empty((6, t.size, N) if save_all else (6, N))
        status = np.empty(N, dtype=np.int32)
        fr = H.frame.spec()
        rc = self._fn("dop853")(H.potential.spec().ptr(), C.byref(fr), w0.ctypes.data_as(C.c_void_p), C.c_size_t(N),
                                t.ctypes.data_as(C.c_void_p), C.c_int(t.size), C.c_double(atol), C.c_double(rtol),
                                C.c_long(nmax), C.c_double(dt_max), C.c_long(nstiff), C.c_int(int(save_all)),
                                C.c_int(nbatch), out.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p))
        return out, status, rc

    def dop853_nfcn(self, H, w0, t, atol=1e-10, rtol=1e-10, nmax=0, dt_max=0.0, nstiff=0, save_all=True, nbatch=1):
        """As ``dop853`` plus, per orbit, the number of right-hand-side calls the reference's own ``dop853()`` made
        (observed through a counting wrapper around ``Fwrapper_T``; = dopcor's ``nfcn`` for ``nbatch=1``).
        Compiled-reference checker only."""
        w0 = _f64(w0); t = _f64(t); N = w0.shape[1]
        out = np.empty((6, t.size, N) if save_all else (6, N))
        status = np.empty(N, dtype=np.int32)
        nfcn = np.empty(N, dtype=np.int32)
        fr = H.frame.spec()
        rc = self._fn("dop853_nfcn")(H.potential.spec().ptr(), C.byref(fr), w0.ctypes.data_as(C.c_void_p), C.c_size_t(N),
                                     t.ctypes.data_as(C.c_void_p), C.c_int(t.size), C.c_double(atol), C.c_double(rtol),
                                     C.c_long(nmax), C.c_double(dt_max), C.c_long(nstiff), C.c_int(int(save_all)),
                                     C.c_int(nbatch), out.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p),
                                     nfcn.ctypes.data_as(C.c_void_p))
        return out, status, rc, nfcn

    def dop853_step_rows(self, H, rows, t1, t2, dt0, atol=1e-10, rtol=1e-10, nmax=0, group=True):
        """rows (Np,6) integrated t1->t2 with dop853_step's settings; group=True: each row alone."""
        rows = _f64(rows).copy(); Np = rows.shape[0]
        status = np.empty(Np, dtype=np.int32)
        fr = H.frame.spec()
        rc = self._fn("dop853_step_rows")(H.potential.spec().ptr(), C.byref(fr), rows.ctypes.data_as(C.c_void_p),
                                          C.c_size_t(Np), C.c_double(t1), C.c_double(t2), C.c_double(dt0),
                                          C.c_double(atol), C.c_double(rtol), C.c_long(nmax), C.c_int(int(group)),
                                          status.ctypes.data_as(C.c_void_p))
        return rows, status, rc

    # ---- massive bodies (reference c_nbody_* + Fwrapper_direct_nbody; ref_driver.cpp) -------------------
    @staticmethod
    def _body_specs(pps):
        import gala_b200 as gb
        null = gb.NullPotential()
        keep = [null.spec()] + [(null if p is None else p).spec() for p in pps]
        arr = (type(keep[0].pot) * len(pps))()
        for b in range(len(pps)):
            arr[b] = keep[b + 1].pot
        return keep, arr

    def nbody_ruth4(self, H, pps, w_rows, t0, nsteps, dt, nsrc=None, save_all=False):
        return self.nbody_leapfrog(H, pps, w_rows, t0, nsteps, dt, nsrc, save_all, fn="nbody_ruth4")

    def nbody_leapfrog(self, H, pps, w_rows, t0, nsteps, dt, nsrc=None, save_all=False, fn="nbody_leapfrog"):
        """rows (n,6): bodies (one potential each in ``pps``, None = massless) then test particles.
        Returns (final rows, traj (nsteps+1, n, 6) | None)."""
        keep, arr = self._body_specs(pps)
        rows = _f64(w_rows).copy(); n = rows.shape[0]
        traj = np.empty((nsteps + 1, n, 6)) if save_all else None
        rc = self._fn(fn)(H.potential.spec().ptr(), arr, C.c_int(len(pps)), keep[0].ptr(),
                                        rows.ctypes.data_as(C.c_void_p), C.c_size_t(n),
                                        C.c_int(len(pps) if nsrc is None else nsrc), C.c_double(t0), C.c_int(nsteps),
                                        C.c_double(dt), None if traj is None else traj.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return rows, traj

    def nbody_dop853(self, H, pps, w_rows, tgrid=None, t1=0.0, t2=0.0, dt0=0.0, atol=1e-10, rtol=1e-10, nmax=0,
                     dt_max=0.0, mode=0, save_all=False):
        """mode 0: direct_nbody_dop853 over ``tgrid`` (dense output when save_all); mode 1: dop853_step t1->t2."""
        keep, arr = self._body_specs(pps)
        rows = _f64(w_rows).copy(); n = rows.shape[0]
        tg = None if tgrid is None else _f64(tgrid)
        traj = np.empty((tg.size, n, 6)) if (save_all and tg is not None) else None
        rc = self._fn("nbody_dop853")(H.potential.spec().ptr(), arr, C.c_int(len(pps)), keep[0].ptr(),
                                      rows.ctypes.data_as(C.c_void_p), C.c_size_t(n),
                                      None if tg is None else tg.ctypes.data_as(C.c_void_p),
                                      C.c_int(0 if tg is None else tg.size), C.c_double(t1), C.c_double(t2),
                                      C.c_double(dt0), C.c_double(atol), C.c_double(rtol), C.c_long(nmax),
                                      C.c_double(dt_max), C.c_int(mode),
                                      None if traj is None else traj.ctypes.data_as(C.c_void_p))
        return rows, traj, rc

    def lyapunov(self, H, w0, d0_vec, t, d0, pullback, atol=1e-10, rtol=1e-10, nmax=0, save=False):
        """dop853_lyapunov_max for ONE parent orbit: returns (LEs_raw (niter, noff), all_w | None, code)."""
        w0 = _f64(w0); d0_vec = _f64(d0_vec); t = _f64(t)
        noff = d0_vec.shape[0]; n_steps = t.size
        LEs = np.zeros((n_steps // pullback, noff))
        allw = np.zeros((n_steps, 1 + noff, 6)) if save else None
        fr = H.frame.spec()
        rc = self._fn("lyapunov")(H.potential.spec().ptr(), C.byref(fr), w0.ctypes.data_as(C.c_void_p),
                                  d0_vec.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), C.c_int(n_steps),
                                  C.c_double(d0), C.c_int(pullback), C.c_int(noff), C.c_double(atol), C.c_double(rtol),
                                  C.c_long(nmax), LEs.ctypes.data_as(C.c_void_p),
                                  None if allw is None else allw.ctypes.data_as(C.c_void_p))
        return LEs, allw, rc

    def d2_dr2(self, pot, q3, t=0.0):
        q3 = _f64(q3)
        return self._fn("d2_dr2", C.c_double)(pot.spec().ptr(), C.c_double(t), q3.ctypes.data_as(C.c_void_p))


class Ref(_Lib):
    def __init__(self, flavor="strict"):
        name = {"strict": "libgala_ref.so", "fast": "libgala_ref_fast.so"}[flavor]
        super().__init__(os.path.join(_REF_DIR, name))
        self.flavor = flavor

    def build_flags(self):
        return self._fn("build_flags", C.c_char_p)().decode()


def have_ref(flavor="strict"):
    name = {"strict": "libgala_ref.so", "fast": "libgala_ref_fast.so"}[flavor]
    return os.path.exists(os.path.join(_REF_DIR, name))


# ---- pure-numpy restatements of the Python-level pieces of the path (no C needed) -------------------
def _sat_frame(ref, pot, px, pv, m, t):
    """get_rj_vj_R (df.pyx:61-92) in numpy scalars; c_d2_dr2 from the compiled reference."""
    G = pot.G
    dist = np.sqrt(px[0] ** 2 + px[1] ** 2 + px[2] ** 2)
    L = np.array([px[1] * pv[2] - px[2] * pv[1], -px[0] * pv[2] + px[2] * pv[0], px[0] * pv[1] - px[1] * pv[0]])
    Lnorm = np.sqrt(L[0] ** 2 + L[1] ** 2 + L[2] ** 2)
    R = np.zeros((3, 3))
    R[0] = px / dist
    R[2] = L / Lnorm
    Om = Lnorm / dist ** 2
    d2r = ref.d2_dr2(pot, px, t)
    rj = (G * m / (Om * Om - d2r)) ** (1 / 3.)
    vj = Om * rj
    a, b = R[0], R[2]
    R[1] = -np.array([a[1] * b[2] - a[2] * b[1], -a[0] * b[2] + a[2] * b[0], a[0] * b[1] - a[1] * b[0]])
    return rj, vj, R


def stream_release_numpy(ref, pot, kind, prog_x, prog_v, prog_t, prog_m, nparticles, random_state, gala_modified=True,
                         lead=True, trail=True, v_disp=0.0):
    """``*StreamDF._sample`` + get_rj_vj_R + transform_from_sat restated in numpy scalars
    (dynamics/mockstream/df.pyx:61-106 and :242-318 streakline, :363-456 fardal, :460-552 lagrange,
    :556-702 chen), drawing the RNG exactly like the reference (scalar ``normal`` calls, or one
    ``multivariate_normal(mean, cov)`` call per particle).  ``ref`` supplies c_d2_dr2
    (cpotential.cpp:346-371)."""
    G = pot.G
    k_mean = np.zeros(6); k_disp = np.zeros(6)
    k_mean[0] = 2.; k_disp[0] = 0.5 if gala_modified else 0.4
    k_mean[2] = 0.; k_disp[2] = 0.5
    k_mean[4] = 0.3; k_disp[4] = 0.5 if gala_modified else 0.4
    k_mean[5] = 0.; k_disp[5] = 0.5
    mean = np.array([1.6, -30., 0., 1., 20., 0.])
    cov = np.zeros((6, 6))
    cov[0, 0] = 0.1225; cov[1, 1] = 529.; cov[2, 2] = 144.; cov[3, 3] = 0.; cov[4, 4] = 400.; cov[5, 5] = 484.
    cov[0, 4] = -4.9; cov[4, 0] = -4.9
    X, V, T1 = [], [], []
    for i in range(len(prog_t)):
        if prog_m[i] == 0:
            continue
        px, pv = prog_x[i], prog_v[i]
        rj, vj, R = _sat_frame(ref, pot, px, pv, prog_m[i], prog_t[i])
        for sgn, on in ((1.0, trail), (-1.0, lead)):
            if not on:
                continue
            for _ in range(int(nparticles[i])):
                tmp_x = np.zeros(3); tmp_v = np.zeros(3)
                if kind == "fardal":
                    kx = random_state.normal(k_mean[0], k_disp[0])
                    tmp_x[0] = kx * (sgn * rj)
                    tmp_x[2] = random_state.normal(k_mean[2], k_disp[2]) * (sgn * rj)
                    tmp_v[1] = random_state.normal(k_mean[4], k_disp[4]) * (sgn * vj)
                    if gala_modified:
                        tmp_v[1] *= kx
                    tmp_v[2] = random_state.normal(k_mean[5], k_disp[5]) * (sgn * vj)
                elif kind == "streakline":
                    tmp_x[0] = sgn * rj
                    tmp_v[1] = sgn * vj
                elif kind == "lagrange":
                    tmp_x[0] = sgn * rj
                    tmp_v[0] = random_state.normal(0, v_disp)
                    tmp_v[1] = random_state.normal(0, v_disp)
                    tmp_v[2] = random_state.normal(0, v_disp)
                elif kind == "chen":
                    pvl = np.array(random_state.multivariate_normal(mean, cov))
                    Dr = pvl[0] * rj
                    Dv = pvl[3] * np.sqrt(2 * G * prog_m[i] / Dr)
                    off = np.pi if sgn < 0 else 0.0
                    pvl[1] = pvl[1] * (np.pi / 180) + off if sgn < 0 else pvl[1] * (np.pi / 180)
                    pvl[2] = pvl[2] * (np.pi / 180)
                    pvl[4] = pvl[4] * (np.pi / 180) + off if sgn < 0 else pvl[4] * (np.pi / 180)
                    pvl[5] = pvl[5] * (np.pi / 180)
                    tmp_x[0] = Dr * np.cos(pvl[2]) * np.cos(pvl[1])
                    tmp_x[1] = Dr * np.cos(pvl[2]) * np.sin(pvl[1])
                    tmp_x[2] = Dr * np.sin(pvl[2])
                    tmp_v[0] = Dv * np.cos(pvl[5]) * np.cos(pvl[4])
                    tmp_v[1] = Dv * np.cos(pvl[5]) * np.sin(pvl[4])
                    tmp_v[2] = Dv * np.sin(pvl[5])
                else:
                    raise ValueError(kind)
                ox = np.array([R[0, k] * tmp_x[0] + R[1, k] * tmp_x[1] + R[2, k] * tmp_x[2] for k in range(3)]) + px
                ov = np.array([R[0, k] * tmp_v[0] + R[1, k] * tmp_v[1] + R[2, k] * tmp_v[2] for k in range(3)]) + pv
                X.append(ox); V.append(ov); T1.append(prog_t[i])
    return np.array(X).reshape(-1, 3), np.array(V).reshape(-1, 3), np.array(T1)


def fardal_release_numpy(ref, pot, prog_x, prog_v, prog_t, prog_m, nparticles, random_state, gala_modified=True,
                         lead=True, trail=True):
    return stream_release_numpy(ref, pot, "fardal", prog_x, prog_v, prog_t, prog_m, nparticles, random_state,
                                gala_modified=gala_modified, lead=lead, trail=trail)


class Port(_Lib):
    """oracle/port.c restatement (double, or long double with ``long_double=True``)."""
    prefix = "port_"

    def __init__(self, long_double=False):
        super().__init__(os.path.join(_REF_DIR, "libgala_port_ld.so" if long_double else "libgala_port.so"))
        self.long_double = long_double

    def build_flags(self):
        return "gcc -std=c11 -O2 -ffp-contract=off (oracle/port.c)"
