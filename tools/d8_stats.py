"""Distribution of per-orbit DOP853 step counts on the C2 bench initial conditions (GPU)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gala_b200 as gb
from bench import make_ic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
H = gb.Hamiltonian(gb.MilkyWayPotential2022())
w0 = make_ic(N, 1000, lambda q: H.potential.gradient(q))
t = np.linspace(0, 1000, 1000)
res = gb.dop853_integrate_hamiltonian(H, w0, t, save_all=1, return_status=True)
st = res[2]
out = {}
for k in ("nstep", "naccpt", "nrejct", "nfcn"):
    v = np.asarray(st[k]).astype(float)
    out[k] = {"min": v.min(), "mean": v.mean(), "median": float(np.median(v)), "p90": float(np.quantile(v, .9)),
              "p99": float(np.quantile(v, .99)), "p999": float(np.quantile(v, .999)), "max": v.max()}
ns = np.asarray(st["nstep"]).astype(float)
w = ns[: (N // 32) * 32].reshape(-1, 32)
out["warp_max_over_mean"] = float(w.max(1).sum() * 32 / w.sum())
out["status_counts"] = {int(k): int(v) for k, v in zip(*np.unique(np.asarray(st["status"]), return_counts=True))}
r0 = np.sqrt((w0[:3] ** 2).sum(0))
i = np.argsort(ns)[-5:]
out["worst"] = [{"nstep": float(ns[j]), "r0": float(r0[j])} for j in i]
print(json.dumps(out, indent=1))
