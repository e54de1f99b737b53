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
FLOPS = {"headline": 146, "c1": 46, "c4": 666, "c2": 3970, "c5": 3400, "c3": 146, "c3d": 0,
         "c3sg": 2 * 146 + 20, "c3sgd": 0}      # self-gravity lane: progenitor + particle gradients + one Plummer term


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
            return res[1]

        def units(N_, out):
            ns = stats["nstep"]
            return int(ns.sum().item() if hasattr(ns, "cpu") else ns.sum())
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
            return stream.w()

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


def cpu_reference_rate(name, H, t, seconds_target=12.0, threads=None):
    """Times the reference's CPU implementation (oracle/_ref/libgala_ref_fast.so; the plain-C port if the
    compiled reference is absent) with all host threads on a bounded sample of the workload."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    import gala_b200 as gb
    if oracle.have_ref("fast"):
        chk, kind = oracle.Ref("fast"), "reference"
    else:
        chk, kind = oracle.Port(), "port"
    threads = threads or os.cpu_count() or 1
    pot = H.potential
    # per-thread sample sized from a quick probe
    probe_n = 64
    w_probe = make_ic(probe_n, 99, lambda q: chk.gradient(pot, q))

    def one(w0):
        if name in ("headline", "c1"):
            chk.leapfrog(pot, w0, t, save_all=(name == "c1")); return w0.shape[1] * (len(t) - 1)
        if name == "c4":
            chk.ruth4(H, w0, t, save_all=False); return w0.shape[1] * (len(t) - 1)
        if name == "c2":
            # reference default nbatch=100 couples the orbits of a batch; count output samples instead of
            # internal steps is not comparable, so use nbatch=1 semantics == what the GPU computes
            out, st, rc = chk.dop853(H, w0, t, save_all=True, nbatch=1); return None
        raise ValueError(name)
    t0 = time.perf_counter(); one(w_probe); dt_probe = time.perf_counter() - t0
    per_thread = int(max(probe_n, min(200_000, probe_n * seconds_target / max(dt_probe, 1e-4))))
    w0s = [make_ic(per_thread, 100 + k, lambda q: chk.gradient(pot, q)) for k in range(threads)]
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:       # ctypes releases the GIL: real parallelism
            units = list(ex.map(one, w0s))
        el = time.perf_counter() - t0
        best = el if best is None else min(best, el)
    if name == "c2":
        # orbit-steps of the adaptive run: take the GPU definition (attempted steps) from a strict rerun is
        # not available on the CPU side without instrumenting the reference; report orbits*outputs instead
        total = threads * per_thread * len(t)
        unit = "orbit-output-samples/s"
    else:
        total = sum(units); unit = "orbit-steps/s"
    return {"value": total / best, "unit": unit, "cores": threads, "kind": kind, "sample_seconds": best,
            "sample": f"{threads} threads x {per_thread} orbits x {len(t) - 1} steps, best of 2, "
                      f"{chk.build_flags() if kind == 'reference' else 'port -O2'}"}


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
            r = cpu_reference_rate(args.workload, H, t, seconds_target=4.0)
            if k >= args.warmup:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals])) if vals else float("nan")
        cb = dict(vals[-1]); cb["value"] = v
        line = {"impl": "reference", "metric": "FP64 orbit-steps/sec", "value": v, "unit": cb["unit"],
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(np.mean([r["sample_seconds"] for r in vals])) * 1e3 if vals else None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": {"workload": desc},
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

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

    if rank == 0:
        flops = FLOPS.get(args.workload, 0)
        per_launch_units = tot_units / max(args.steps, 1)
        mean_kern_s = float(np.mean(kern_ms)) * 1e-3
        achieved_tf = flops * per_launch_units / mean_kern_s / 1e12
        traffic = None
        tr_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr_file):
            traffic = json.load(open(tr_file)).get(args.workload)
        line = {
            "metric": "FP64 orbit-steps/sec", "value": value, "unit": "orbit-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "orbits_per_gpu": N, "ntimes": int(len(t)),
                       "math": "strict" if args.strict else "fast",
                       "l2": f"inputs {w0_host.nbytes / 1e6:.0f} MB per launch" +
                             (" > 126 MB L2" if w0_host.nbytes > 126e6 else " (compute-bound kernel: inputs read once per 1000 steps)"),
                       "parallelism": f"orbit-index sharding x{world}, no collectives"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf > 0 else None, "traffic": traffic,
                         "flops_per_orbit_step": flops,
                         "peak_source": "DFMA microbenchmark measured live (gala_b200/csrc/peak.cu); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "clocks": clk, "gpu_launches": launches, "e2e": e2e,
        }
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
