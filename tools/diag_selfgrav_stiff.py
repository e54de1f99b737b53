"""Diagnostic (test infrastructure, uses the oracle): C3 with progenitor self-gravity, DOP853: which stream particles end with the stiffness code -4 on the GPU,
and does the compiled reference (lane by lane) end the same particles the same way?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gala_b200 as gb
from gala_b200.mockstream import DirectNBody, _nbody_dop853
from oracle import oracle

b = float(sys.argv[1]) if len(sys.argv) > 1 else 0.004
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
ref = oracle.Ref("strict")
pot = gb.MilkyWayPotential2022(); H = gb.Hamiltonian(pot)
prog = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * gb.KMS_TO_KPC_MYR, 50.0 * gb.KMS_TO_KPC_MYR])
pp = gb.PlummerPotential(m=2.5e4, b=b)
t = np.arange(n_steps + 1) * -1.0
nb = DirectNBody(prog, [pp], external_potential=pot)
orb = nb.integrate_orbit(t=t, Integrator="dopri853")
po = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
s0 = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)).sample(po, 2.5e4, n_particles=10)
w0 = np.ascontiguousarray(np.vstack([s0.pos, s0.vel]).T)
time = np.asarray(po.t)
body0 = np.concatenate([po.pos[:, 0], po.vel[:, 0]])[None, :]
_, _, traj, _ = _nbody_dop853(H, [pp], body0, time, time[-1], time[1] - time[0], 0, save_all=True)
unq, nstream = np.unique(s0.release_time, return_counts=True)
group = np.repeat(np.arange(len(unq), dtype=np.int32), nstream)
t1 = np.repeat(unq, nstream)
out_p, out_b, _, status = _nbody_dop853(H, [pp], traj, None, time[-1], time[1] - time[0], 1, w0_rows=w0, t1=t1, group=group,
                                        err_if_fail=0)
bad = np.nonzero(status < 0)[0]
print(f"b={b}: {bad.size} of {status.size} particles end with a negative code; codes {np.unique(status, return_counts=True)}")
for p in bad[:8]:
    rows = np.vstack([traj[group[p]], w0[p:p + 1]])
    fin, _, rc = ref.nbody_dop853(H, [pp], rows, t1=t1[p], t2=time[-1], dt0=time[1] - time[0], mode=1)
    d0 = np.sqrt(((w0[p, :3] - traj[group[p], 0, :3]) ** 2).sum())
    print(f"  particle {p}: gpu {status[p]}  reference {rc}  released {d0 * 1e3:.1f} pc from the progenitor at t={t1[p]:.0f}")
ok = np.nonzero(status > 0)[0][:200]
worst = 0.0
for p in ok[::10]:
    rows = np.vstack([traj[group[p]], w0[p:p + 1]])
    fin, _, rc = ref.nbody_dop853(H, [pp], rows, t1=t1[p], t2=time[-1], dt0=time[1] - time[0], mode=1)
    worst = max(worst, np.abs(fin[1] - out_p[p]).max() / np.abs(fin[1]).max())
print(f"  agreement with the reference on 20 successful particles: max rel {worst:.2e}")
