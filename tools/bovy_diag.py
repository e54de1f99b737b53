"""BovyMWPotential2014 leapfrog: device-timed (CUDA events) with and without the gamma* table, and per component."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gala_b200 as gb
from bench import make_ic

N = 3031040
t = np.arange(1001.0)


def run(name, pot, w0):
    H = gb.Hamiltonian(pot)
    for _ in range(2):
        gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0)
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gb.leapfrog_integrate_hamiltonian(H, w0, t, save_all=0); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    print(f"{name:40s} {min(ms):7.2f} ms (min of 3; max {max(ms):.2f})", flush=True)


bovy = gb.BovyMWPotential2014()
w0 = torch.as_tensor(make_ic(N, 1, lambda q: bovy.gradient(q)), device="cuda")
r = torch.linalg.norm(w0[:3], dim=0)
print("IC radius kpc: min %.3f median %.3f max %.3f" % (r.min().item(), r.median().item(), r.max().item()))
run("bovy2014 (table)", bovy, w0)
if "--one" in sys.argv:
    sys.exit(0)
os.environ["GB_PLC_NO_TABLE"] = "1"
run("bovy2014 (series / continued fraction)", bovy, w0)
del os.environ["GB_PLC_NO_TABLE"]
run("disk only (MiyamotoNagai)", bovy["disk"], w0)
run("bulge only (PowerLawCutoff, table)", bovy["bulge"], w0)
run("halo only (NFW)", bovy["halo"], w0)
run("disk + halo", gb.CCompositePotential(disk=bovy["disk"], halo=bovy["halo"]), w0)
run("mw2022 on the same ICs", gb.MilkyWayPotential2022(), w0)
