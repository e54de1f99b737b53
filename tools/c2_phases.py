"""Where does one C2 step (MW2022 DOP853, 1000 dense-output times) spend its time?  Host wall clock and
CUDA-event time of the whole call, of the statistics read-back, and of the result allocation."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gala_b200 as gb
from bench import make_ic

N = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 8
save_all = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H = gb.Hamiltonian(gb.MilkyWayPotential2022())
w0 = torch.as_tensor(make_ic(N, 1000, lambda q: H.potential.gradient(q)), device="cuda")
t = np.linspace(0, 1000, 1000)
rows = []
for k in range(6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    res = gb.dop853_integrate_hamiltonian(H, w0, t, save_all=save_all, return_status=True)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ns = int(res[2]["nstep"].sum().item())
    t3 = time.perf_counter()
    del res
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    rows.append({"call_wall_ms": (t1 - t0) * 1e3, "sync_after_ms": (t2 - t1) * 1e3, "event_ms": e0.elapsed_time(e1),
                 "stats_ms": (t3 - t2) * 1e3, "free_ms": (t4 - t3) * 1e3, "nstep_total": ns})
print(json.dumps({"N": N, "save_all": save_all, "lib": os.environ.get("GALA_B200_LIB", "default"),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("GB_")}, "rows": rows}))
