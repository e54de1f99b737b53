#!/usr/bin/env python
"""bench.py -- orbit-steps/s of the hot path on N B200s (one process per GPU, no collectives on the
data path: orbits are sharded by index, each rank integrates its own contiguous slice).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic orbits (one kernel launch that
advances every orbit of the batch through the whole time grid).  Default workload = the headline
configuration BASELINE.json quotes its target on: MilkyWayPotential2022, LeapfrogIntegrator, dt = 1 Myr,
1000 steps, final-state-only output.  Other workloads (c1, c2, c4) are
the BASELINE.json configs; they print the same JSON line when selected explicitly.

value : whole-job orbit-steps/s with inputs resident in HBM (CUDA events, max over ranks).
e2e   : the same metric through the public host-buffer API (numpy in pinned memory -> C ABI with
        GB_MEM_HOST: H2D of the initial conditions, kernel, D2H of the result, every step).
roofline : FP64 CUDA-core roofline.  achieved = algorithmic flops per launch / mean launch time;
           peak = DFMA throughput measured live on the same GPU by gala_b200/csrc/peak.cu
           (MEASURED_PEAKS.json has no FP64 entry; its HBM figure is quoted for the save_all case).
cpu_baseline : the reference's own C++ (oracle/_ref/libgala_ref_fast.so, built with the reference's
           shipped flags) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SM_FILL = 148 * 2048          # resident threads of one B200 at full occupancy

# algorithmic flops per orbit-step (source-level op counts)
FLOPS = {"headline": 146, "c1": 46, "c1x": 46, "c4": 666, "c2": 3970, "c5": 3400, "c3": 146, "c3d": 0,
         "c3sg": 2 * 146 + 20, "c3sgd": 0}      # self-gravity lane: progenitor + particle gradients + one Plummer term
# workloads whose kernel is bound by the trajectory write, not by the FP64 pipe: the roofline is HBM GB/s against
# MEASURED_PEAKS.json.  Algorithmic bytes per orbit-step = 48 (one (x, v) row written), plus
# the 48-byte initial condition read once per orbit.
HBM_BOUND = {"c1", "c1x"}
# total orbits of the strong-scaling single-call measurement (one host array, one C-ABI call over all devices):
# 8 x the per-GPU default, C5 at exactly the 10^7 orbits of BASELINE.json configs[4]
TOTAL_ORBITS = {"headline": 8 * 10 * SM_FILL, "c4": 8 * 4 * SM_FILL, "c5": 10_000_000,
                # trajectory-output workloads: the call is bound by the PCIe read-back, which every device does over
                # its own link -- C1 at exactly BASELINE configs[0]'s 10^4 orbits, C1x and the C2 chunk at the 1-GPU sizes
                "c1": 10_000, "c1x": 1 << 20, "c2": 148 * 256 * 8}


def make_ic(N, seed, pot_gradient, rmin=4.0, rmax=50.0):
    """Seeded bound orbits: r = exp(U[ln rmin, ln rmax]) kpc with isotropic direction; speed =
    f * v_circ(r), f ~ U[0.5, 1.0]; velocity direction mostly tangential (radial direction cosine
    mu ~ U[-0.5, 0.5]).  SURVEY.md 8d proposed a fully isotropic velocity direction; that produces
    plunging orbits (pericentre < 0.3 kpc through the 0.07-kpc nucleus) on which the reference does
    not reproduce ITSELF between its -O2 and -Ofast builds after 1000 steps (differences of order
    unity), so parity there measures chaos, not the implementation.  ``pot_gradient(q)`` -> (3,N)."""
    rng = np.random.default_rng(seed)
    r = np.exp(rng.uniform(np.log(rmin), np.log(rmax), N))
    mu = rng.uniform(-1, 1, N); ph = rng.uniform(0, 2 * np.pi, N)
    s = np.sqrt(1 - mu * mu)
    rhat = np.vstack([s * np.cos(ph), s * np.sin(ph), mu])
    q = r * rhat
    # two unit vectors orthogonal to rhat
    e1 = np.vstack([-np.sin(ph), np.cos(ph), np.zeros(N)])
    e2 = np.cross(rhat.T, e1.T).T
    psi = rng.uniform(0, 2 * np.pi, N)
    cr = rng.uniform(-0.5, 0.5, N)
    vhat = cr * rhat + np.sqrt(1 - cr * cr) * (np.cos(psi) * e1 + np.sin(psi) * e2)
    g = pot_gradient(np.ascontiguousarray(q))
    vc = np.sqrt(r * np.sqrt((g * g).sum(0)))
    v = rng.uniform(0.5, 1.0, N) * vc * vhat
    return np.ascontiguousarray(np.vstack([q, v]))


def workload(name, n_orbits):
    import gala_b200 as gb
    if name == "headline":
        H = gb.Hamiltonian(gb.MilkyWayPotential2022())
        t = np.arange(1001, dtype=float)
        N = n_orbits or 10 * SM_FILL
        desc = f"MilkyWayPotential2022 leapfrog dt=1Myr 1000 steps final-state-only, {N} orbits/GPU"
        run = lambda w0, tt, out=None: gb.leapfrog_integrate_hamiltonian(H, w0, tt, save_all=0, out=out)[1]
        units = lambda N_, out: N_ * 1000
    elif name == "c1":
        H = gb.Hamiltonian(gb.NFWPotential(m=1e11, r_s=12.0))
        t = np.arange(1001, dtype=float)
        N = n_orbits or 10_000
        desc = f"C1: NFWPotential(m=1e11,r_s=12) leapfrog dt=1Myr 1000 steps save_all, {N} orbits/GPU"
        run = lambda w0, tt, out=None: gb.leapfrog_integrate_hamiltonian(H, w0, tt, save_all=1, out=out)[1]
        units = lambda N_, out: N_ * 1000
    elif name == "c1x":
        # the save_all path at a size that fills the GPU: the kernel writes 48 B per orbit-step and does 46 flops
        # for it, so it is bound by HBM write bandwidth (north_star: "achieved HBM GB/s when trajectory output dominates")
        H = gb.Hamiltonian(gb.NFWPotential(m=1e11, r_s=12.0))
        t = np.arange(257, dtype=float)
        N = n_orbits or (1 << 20)
        desc = f"C1x: NFWPotential(m=1e11,r_s=12) leapfrog dt=1Myr 256 steps save_all, {N} orbits/GPU (C1 at GPU-filling size)"
        run = lambda w0, tt, out=None: gb.leapfrog_integrate_hamiltonian(H, w0, tt, save_all=1, out=out)[1]
        units = lambda N_, out: N_ * 256
    elif name == "c4":
        pot = gb.CCompositePotential()
        pot["bar"] = gb.LongMuraliBarPotential(m=1e10, a=4.0, b=0.8, c=0.25, alpha=np.deg2rad(25.0))
        for k, v in gb.MilkyWayPotential2022().items():
            pot[k] = v
        H = gb.Hamiltonian(pot, gb.ConstantRotatingFrame([0.0, 0.0, 0.030681]))
        t = np.arange(1001) * 0.5
        N = n_orbits or 4 * SM_FILL
        desc = f"C4: LongMuraliBar+MW2022 Ruth4 dt=0.5Myr 1000 steps ConstantRotatingFrame final-state, {N} orbits/GPU"
        run = lambda w0, tt, out=None: gb.ruth4_integrate_hamiltonian(H, w0, tt, save_all=0, allow_rotating_frame=True, out=out)[1]
        units = lambda N_, out: N_ * 1000
    elif name == "c2":
        H = gb.Hamiltonian(gb.MilkyWayPotential2022())
        t = np.linspace(0, 1000, 1000)
        N = n_orbits or 148 * 256 * 8
        desc = f"C2: MW2022 DOP853 atol=rtol=1e-10, 1000 dense-output times, {N} orbits/GPU (chunk of the 1e6)"
        stats = {}

        def run(w0, tt, out=None):
            res = gb.dop853_integrate_hamiltonian(H, w0, tt, save_all=1, return_status=True, out=out)
            stats["nstep"] = res[2]["nstep"]
            stats["naccpt"] = res[2]["naccpt"]
            return res[1]

        def units(N_, out):
            ns = stats["nstep"]
            return int(ns.sum().item() if hasattr(ns, "cpu") else ns.sum())

        def flops_of(N_, out):
            # algorithmic flops of the steps actually taken: an accepted dense step costs 3970 (12 + 1 + 3 right-hand
            # sides, stage sums, error norm, dense-output coefficients), a rejected one 2460 (12 right-hand sides, stage
            # sums, error norm); SURVEY.md appendix C.  Round 1 charged 3970 to every attempted step.
            ns, na = stats["nstep"], stats["naccpt"]
            tot = lambda a: int(a.sum().item() if hasattr(a, "cpu") else a.sum())
            return 3970.0 * tot(na) + 2460.0 * (tot(ns) - tot(na))
        units.flops_of = flops_of
    elif name in ("c3", "c3d", "c3sg", "c3sgd"):
        # C3: 10^5-particle Fardal stream in MW2022 (tests/dynamics/mockstream/test_mockstream.py:676-678
        # progenitor); c3 = LeapfrogIntegrator (exact orbit-step count), c3d = DOPRI853 (reference default)
        H = gb.Hamiltonian(gb.MilkyWayPotential2022())
        prog = gb.PhaseSpacePosition(pos=[13.0, 0.0, 20.0], vel=np.array([0.0, 130.0, 50.0]) * gb.KMS_TO_KPC_MYR)
        n_steps, n_part = 5000, 10
        t = np.arange(n_steps + 1) * -1.0
        N = 2 * n_part * (n_steps + 1)
        integ = gb.LeapfrogIntegrator if name in ("c3", "c3sg") else gb.DOPRI853Integrator
        # c3sg / c3sgd: the same stream with the progenitor's own gravity (a 2.5e4 Msun Plummer sphere, b = 50 pc):
        # every device lane integrates [progenitor, one particle] (csrc/nbody.cuh)
        selfgrav = name in ("c3sg", "c3sgd")
        desc = (f"C3: FardalStreamDF(gala_modified, RandomState(42)) in MW2022, dt=-1Myr x {n_steps}, {n_part} particles "
                f"per tail per step = {N} particles, {integ.__name__}, "
                f"{'progenitor self-gravity (Plummer 2.5e4 Msun, b=50pc), ' if selfgrav else ''}"
                f"whole MockStreamGenerator.run per step")

        def run(w0, tt, out=None):
            gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)), H,
                                         progenitor_potential=gb.PlummerPotential(m=2.5e4, b=0.05) if selfgrav else None)
            stream, _ = gen.run(prog, 2.5e4, dt=-1.0, n_steps=n_steps, n_particles=n_part, release_every=1,
                                Integrator=integ, Integrator_kwargs={"err_if_fail": 0} if integ is gb.DOPRI853Integrator else None)
            return stream.pos          # the MockStream object IS the result (pos / vel views of the (Np, 6) rows, like the reference's)

        # fixed-step count: particle released at step k takes k steps (+ the progenitor orbit itself)
        units = lambda N_, out: 2 * n_part * (n_steps * (n_steps + 1) // 2) + n_steps
    elif name == "c5":
        rng = np.random.default_rng(5)
        nmax, lmax = 10, 6
        S = np.zeros((nmax + 1, lmax + 1, lmax + 1)); T = np.zeros_like(S)
        for n in range(nmax + 1):
            for l in range(lmax + 1):
                for m in range(l + 1):
                    sig = 0.05 / (1 + n + l) ** 2
                    S[n, l, m] = rng.normal(0, sig)
                    if m > 0:
                        T[n, l, m] = rng.normal(0, sig)
        S[0, 0, 0] = 1.0
        H = gb.Hamiltonian(gb.SCFPotential(m=1e12, r_s=20.0, Snlm=S, Tnlm=T))
        t = np.arange(1001, dtype=float)
        N = n_orbits or 1_250_000           # 10^7 orbits / 8 GPUs
        desc = f"C5: SCFPotential(nmax=10,lmax=6) leapfrog dt=1Myr 1000 steps final-state-only, {N} orbits/GPU"
        run = lambda w0, tt, out=None: gb.leapfrog_integrate_hamiltonian(H, w0, tt, save_all=0, out=out)[1]
        units = lambda N_, out: N_ * 1000
    else:
        raise SystemExit(f"unknown workload {name}")
    return H, t, N, desc, run, units


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md clocks line).  Sampled
    in-process through NVML every 20 ms: a polling `nvidia-smi -lms` child serialises against this
    process's own driver calls (measured: it stretched a 34 ms C2 step to 108 ms); nvidia-smi is only the
    fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, force_smi=False):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.nv, self.h, self.stop_flag, self.thread = None, None, False, None
        self.sm, self.smax, self.reasons, self.how = [], None, set(), None
        if not force_smi:
            try:
                import pynvml
                pynvml.nvmlInit()
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
                self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
                self.nv = pynvml
            except Exception:
                self.nv = None

    def _poll_nvml(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self.how = "nvml 20ms"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        try:
            self.how = "nvidia-smi -lms 100"
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "how": self.how}


def _cpu_jobs(name, H, t, chk, n_jobs, per_job_seconds):
    """Bounded samples of workload `name` for the CPU arm: a list of n_jobs callables, each integrating its own
    sample with the reference's CPU code and returning the units it processed, in the SAME unit the GPU arm
    reports (orbit-steps; DOP853: attempted internal steps; mock streams: particle-steps).  Also returns a
    description of one sample."""
    import gala_b200 as gb
    pot = H.potential
    grad = lambda q: chk.gradient(pot, q)

    if name in ("headline", "c1", "c1x", "c4", "c5", "c2"):
        if name == "c2":
            def one(w0):
                # attempted steps of the reference's own dop853() (nbatch=1 = the per-orbit control the GPU computes),
                # observed through its right-hand-side calls: dense = 1 + 11 nstep + 4 naccpt, final = 1 + 11 nstep + naccpt
                _, _, _, calls_d = chk.dop853_nfcn(H, w0, t, save_all=True, nbatch=1)
                return calls_d
            def steps_of(w0, calls_d):      # untimed: the second observation that separates nstep from naccpt
                _, _, _, calls_f = chk.dop853_nfcn(H, w0, t, save_all=False, nbatch=1)
                nacc = (calls_d - calls_f) // 3
                return int(((calls_f - 1 - nacc) // 11).sum())
        elif name == "c4":
            one = lambda w0: chk.ruth4(H, w0, t, save_all=False) is None or w0.shape[1] * (len(t) - 1)
        else:
            sv = name in ("c1", "c1x")
            one = lambda w0: chk.leapfrog(pot, w0, t, save_all=sv) is None or w0.shape[1] * (len(t) - 1)
        probe_n = 16 if name == "c5" else 64
        w_probe = make_ic(probe_n, 99, grad)
        t0 = time.perf_counter(); one(w_probe); dt_probe = time.perf_counter() - t0
        per_job = int(max(probe_n // 4, min(200_000, probe_n * per_job_seconds / max(dt_probe, 1e-4))))
        w0s = [make_ic(per_job, 100 + k, grad) for k in range(n_jobs)]
        if name == "c2":
            def mk(w0):
                def job():
                    job.calls = one(w0)
                    return 0
                job.post = lambda: steps_of(w0, job.calls)
                return job
            return [mk(w) for w in w0s], f"{per_job} orbits x {len(t)} dense-output times"
        return [(lambda w=w: one(w)) for w in w0s], f"{per_job} orbits x {len(t) - 1} steps"

    if name in ("c3", "c3d", "c3sg", "c3sgd"):
        # the reference's mock-stream loops (dynamics/mockstream/mockstream.pyx:442-620 leapfrog, :176-303 dop853):
        # for every release time the rows [progenitor, its particles] are integrated from there to tfinal --
        # fixed step: c_leapfrog_step_nbody over the rows; adaptive: ONE dop853 system of all rows (shared step).
        # A sample = every `stride`-th release time (work per release time is linear in the remaining time, so a
        # uniform subset is unbiased); units = particle-steps, as the GPU arm counts them.
        from oracle import oracle
        n_steps, n_part = 5000, 10
        selfgrav = name in ("c3sg", "c3sgd")
        pps = [gb.PlummerPotential(m=2.5e4, b=0.05) if selfgrav else None]
        w_prog = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * gb.KMS_TO_KPC_MYR, 50.0 * gb.KMS_TO_KPC_MYR]).reshape(6, 1)
        tt = np.arange(n_steps + 1) * -1.0
        prog = chk.leapfrog(pot, w_prog, tt, save_all=True)[:, ::-1, 0]      # (6, ntimes) from the earliest time on
        tf = tt[::-1].copy()
        # groups cost ~ (remaining steps) x 21 rows; probe one mid group to size the stride
        def group_rows(i, rng):
            x, v, _ = oracle.fardal_release_numpy(chk, pot, prog[:3, i:i + 1].T, prog[3:, i:i + 1].T, tf[i:i + 1],
                                                  np.array([2.5e4]), np.array([n_part]), rng, gala_modified=True)
            return np.vstack([prog[:, i], np.hstack([x, v])])
        def run_group(i, rows):
            nst = n_steps - i
            if nst == 0:
                return 0
            if name in ("c3", "c3sg"):
                chk.nbody_leapfrog(H, pps, rows, tf[i], nst, 1.0)
            elif selfgrav:
                chk.nbody_dop853(H, pps, rows, t1=tf[i], t2=tf[-1], dt0=1.0, mode=1)
            else:
                chk.dop853_step_rows(H, rows, tf[i], tf[-1], 1.0, group=False)
            return 2 * n_part * nst
        rng = np.random.RandomState(42)
        rows_mid = group_rows(n_steps // 2, rng)
        t0 = time.perf_counter(); run_group(n_steps // 2, rows_mid); dt_mid = time.perf_counter() - t0
        n_groups = int(max(2, min(400, per_job_seconds / max(dt_mid, 1e-4))))
        jobs = []
        for k in range(n_jobs):
            idx = np.linspace(0, n_steps - 1, n_groups, dtype=int) + (k % max(1, n_steps // n_groups // 2))
            idx = np.clip(idx, 0, n_steps - 1)
            rows = [group_rows(int(i), rng) for i in idx]
            jobs.append(lambda idx=idx, rows=rows: sum(run_group(int(i), r) for i, r in zip(idx, rows)))
        return jobs, f"{n_groups} of the {n_steps + 1} release times x {2 * n_part} particles (+ the progenitor row)"
    raise ValueError(f"no CPU arm for workload {name}")


def cpu_reference_rate(name, H, t, seconds_target=12.0, threads=None, one_thread=True):
    """Times the reference's CPU implementation (oracle/_ref/libgala_ref_fast.so = the reference's own C++ built with
    its shipped flags; the plain-C port if the compiled reference is absent) on a bounded sample of the workload:
    with all host threads (`value`, `cores`) and on one thread (`value_1thread`; the reference itself is
    single-threaded).  Unit = the GPU arm's."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    if oracle.have_ref("fast"):
        chk, kind = oracle.Ref("fast"), "reference"
    else:
        chk, kind = oracle.Port(), "port"
    threads = threads or os.cpu_count() or 1

    def timed(n_jobs, seconds, repeats):
        jobs, what = _cpu_jobs(name, H, t, chk, n_jobs, seconds)
        best, units = None, 0
        for _ in range(repeats):
            t0 = time.perf_counter()
            if n_jobs == 1:
                got = [jobs[0]()]
            else:
                with ThreadPoolExecutor(n_jobs) as ex:       # ctypes releases the GIL: real parallelism
                    got = list(ex.map(lambda j: j(), jobs))
            el = time.perf_counter() - t0
            best = el if best is None else min(best, el)
            units = sum(got)
        if hasattr(jobs[0], "post"):                          # c2: step counts need a second (untimed) observation
            units = sum(j.post() for j in jobs)
        return units / best, best, what

    v_all, el_all, what = timed(threads, seconds_target, 2)
    flags = chk.build_flags() if kind == "reference" else "port -O2"
    out = {"value": v_all, "unit": "orbit-steps/s", "cores": threads, "kind": kind, "sample_seconds": el_all,
           "sample": f"{threads} threads x ({what}), best of 2, {flags}"}
    if one_thread:
        v_1, el_1, what1 = timed(1, max(2.0, seconds_target / 3.0), 1)
        out["value_1thread"] = v_1
        out["sample_1thread"] = f"1 thread x ({what1}), {el_1:.1f} s"
    return out


def config_of(desc, N, ntimes, world):
    """The `config` object of the JSON line; identical on the GPU arm and on the reference arm."""
    in_bytes = 48 * N
    out_bytes = 48 * N * (ntimes if "save_all" in desc or "dense-output" in desc else 1)
    big = max(in_bytes, out_bytes) > 126e6
    return {"workload": desc, "orbits_per_gpu": N, "ntimes": int(ntimes),
            "l2": f"inputs {in_bytes / 1e6:.0f} MB, outputs {out_bytes / 1e6:.0f} MB per launch" +
                  (" > 126 MB L2" if big else " (compute-bound kernel: inputs read once per launch, state in registers)"),
            "parallelism": f"orbit-index sharding x{world}, no collectives"}


def measured_hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-written, copy bandwidth)"
    except Exception:
        return 6500.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def single_call_strong(args, gb, torch, H, t, run, units, world, name):
    """Strong scaling through the drop-in API (north_star: "each GPU takes a contiguous slice of orbits and results
    are gathered to the host"): ONE pinned host array of TOTAL_ORBITS[name] orbits, ONE call of the public function
    with gb.set_devices(range(world)) -- the C ABI shards it (gb_launch.n_devices), every device copies its slice
    in, integrates, copies its slice of the result straight into the one output array.  Runs on rank 0 only while
    the other ranks wait on a CPU barrier (their GPUs are idle and are driven by rank 0's process here)."""
    total = args.total_orbits or TOTAL_ORBITS.get(name)
    if not total:
        return {"skipped": f"the single-call leg is defined for {sorted(TOTAL_ORBITS)}"}
    if gb._abi.device_count() < world:
        return {"skipped": f"rank 0 sees {gb._abi.device_count()} devices, needs {world}"}
    w0 = make_ic(total, 4242, lambda q: H.potential.gradient(q))
    pin_in = gb.pinned_empty(w0.shape); pin_in[...] = w0
    save_all = name in ("c1", "c1x", "c2")
    pin_out = gb.pinned_empty((6, len(t), total) if save_all else w0.shape)
    devs = list(range(world))
    gb.set_devices(devs)
    nunits = 0
    steps = args.steps if not save_all else min(args.steps, 5)       # multi-GB read-backs: a few calls are enough
    try:
        for _ in range(2 if save_all else 3):
            out = run(pin_in, t, pin_out)
            units(total, out)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = run(pin_in, t, pin_out)          # returns when every device has delivered its slice
            nunits += units(total, out)
        el = time.perf_counter() - t0
    finally:
        gb.set_devices(None)
    return {"value": nunits / el, "unit": "orbit-steps/s", "ms_per_call": el / steps * 1e3,
            "orbits": int(total), "devices": devs, "scaling": "strong",
            "h2d_bytes_per_call": int(pin_in.nbytes + t.nbytes), "d2h_bytes_per_call": int(pin_out.nbytes),
            "what": "one pinned host (6,N) array in, one C-ABI call over all devices (gb_launch.n_devices), one "
                    + ("(6,ntimes,N)" if save_all else "(6,N)") + " array out; wall clock on rank 0"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--orbits", type=int, default=0, help="orbits per GPU (0 = workload default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strict", action="store_true", help="use the strict-IEEE kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-single-call", action="store_true", help="skip the strong-scaling single-call measurement")
    ap.add_argument("--total-orbits", type=int, default=0, help="orbits of the single-call measurement (0 = workload default)")
    ap.add_argument("--smi-clocks", action="store_true", help="sample clocks with an nvidia-smi child (A/B only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "b200":
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))

    import gala_b200 as gb

    if args.impl == "reference":
        if rank != 0:
            return
        H, t, N, desc, run, units = workload(args.workload, args.orbits)
        vals = []
        for k in range(args.warmup + args.steps):
            r = cpu_reference_rate(args.workload, H, t, seconds_target=4.0, one_thread=(k == args.warmup + args.steps - 1))
            if k >= args.warmup:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals])) if vals else float("nan")
        cb = dict(vals[-1]); cb["value"] = v
        line = {"impl": "reference", "metric": "FP64 orbit-steps/sec", "value": v, "unit": cb["unit"],
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(np.mean([r["sample_seconds"] for r in vals])) * 1e3 if vals else None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_of(desc, N, len(t), args.gpus),
                "cpu_baseline": cb,
                "e2e": {"value": v, "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if gb._abi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")      # a barrier that does not put a spinning kernel on the GPUs

    H, t, N, desc, run, units = workload(args.workload, args.orbits)
    H.strict_math = args.strict
    # per-rank slice of the global orbit index: seeds differ per rank, work per GPU is fixed (weak scaling)
    grad = lambda q: H.potential.gradient(q)
    w0_host = make_ic(N, 1000 + rank, grad)
    w0_dev = torch.as_tensor(w0_host, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 peak measured live (burst figure, kernel alone)
    peak_tf = gb._abi.lib().gb_fp64_peak_tflops(5)

    for _ in range(args.warmup):
        out = run(w0_dev, t)
        units(N, out)      # also warms the unit count (C2 sums per-orbit step counts with a lazily loaded torch kernel)
    torch.cuda.synchronize()

    clocks = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else
                          int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]), force_smi=args.smi_clocks)
    if rank == 0:
        clocks.start()
    n0 = gb._abi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    tot_units = 0
    barrier()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    trace = []
    for k in range(args.steps):
        ta = time.perf_counter()
        ev[k][0].record()
        out = run(w0_dev, t)
        ev[k][1].record()
        tb = time.perf_counter()
        tot_units += units(N, out)
        trace.append((tb - ta, time.perf_counter() - tb))
    e_stop.record()
    barrier()
    launches = gb._abi.launch_count() - n0
    if os.environ.get("BENCH_TRACE"):
        print("trace (call s, units s):", [(round(a * 1e3, 2), round(b * 1e3, 2)) for a, b in trace], file=sys.stderr)
    clk = clocks.stop() if rank == 0 else None
    ms_total = e_start.elapsed_time(e_stop)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    uu = torch.tensor([float(tot_units)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(uu, op=dist.ReduceOp.SUM)
    ms_max, units_all = tt.item(), uu.item()
    value = units_all / (ms_max * 1e-3)

    # end-to-end through the host-buffer API (pinned numpy in, numpy out)
    e2e = None
    if not args.no_e2e:
        pin = torch.empty(w0_host.shape, dtype=torch.float64).pin_memory()
        pin.numpy()[...] = w0_host
        w0_pinned = pin.numpy()
        out_shape = (6, len(t), N) if args.workload == "c2" else np.asarray(run(w0_pinned, t)).shape
        # the result lands in a caller-provided page-locked array (gb.pinned_empty), as the inputs do:
        # with pageable arrays the driver stages every copy and page-faults the fresh result buffer
        out_pin = gb.pinned_empty(out_shape)
        out_h = run(w0_pinned, t, out_pin)
        e2e_units = 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_h = run(w0_pinned, t, out_pin)   # returns after the D2H copy completed
            e2e_units += units(N, out_h)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        ue = torch.tensor([float(e2e_units)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ue, op=dist.ReduceOp.SUM)
        e2e = {"value": ue.item() / te.item(), "unit": "orbit-steps/s",
               "h2d_bytes_per_step": int(w0_host.nbytes + t.nbytes), "d2h_bytes_per_step": int(np.asarray(out_h).nbytes)}

    # strong scaling through ONE call over all devices, rank 0 only (the other ranks idle on a CPU barrier)
    single = None
    if not args.no_single_call:
        if world > 1:
            dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                single = single_call_strong(args, gb, torch, H, t, run, units, world, args.workload)
            except Exception as e:
                single = {"error": f"{type(e).__name__}: {e}"}
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank == 0:
        flops = FLOPS.get(args.workload, 0)
        per_launch_units = tot_units / max(args.steps, 1)
        mean_kern_s = float(np.mean(kern_ms)) * 1e-3
        per_launch_flops = units.flops_of(N, out) if hasattr(units, "flops_of") else flops * per_launch_units
        achieved_tf = per_launch_flops / mean_kern_s / 1e12
        traffic = None
        tr_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr_file):
            traffic = json.load(open(tr_file)).get(args.workload)
        line = {
            "metric": "FP64 orbit-steps/sec", "value": value, "unit": "orbit-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(desc, N, len(t), world), "math": "strict" if args.strict else "fast",
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf > 0 else None, "traffic": traffic,
                         "flops_per_orbit_step": flops if not hasattr(units, "flops_of") else per_launch_flops / max(per_launch_units, 1),
                         "peak_source": "DFMA microbenchmark measured live (gala_b200/csrc/peak.cu); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "clocks": clk, "gpu_launches": launches, "e2e": e2e, "single_call": single,
        }
        if args.workload in HBM_BOUND:
            # trajectory output dominates: 48 B written per orbit-step (+ 48 B read per orbit), against the measured HBM peak
            hbm_peak, hbm_src = measured_hbm_peak()
            alg_bytes = 48.0 * per_launch_units + 48.0 * N
            gbs = alg_bytes / mean_kern_s / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                "traffic": traffic, "bytes_per_orbit_step": 48, "peak_source": hbm_src,
                                "fp64_frac": achieved_tf / peak_tf if peak_tf > 0 else None}
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_reference_rate(args.workload, H, t)
            except Exception as e:       # the oracle is test infrastructure; its absence must not kill the bench
                line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
