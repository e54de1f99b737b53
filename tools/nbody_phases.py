"""Wall time of the three device calls of a self-gravitating C3 stream (DOP853): progenitor orbit (one
lane, dense), bodies at the release times (one lane, dense), the stream particles (one lane each)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gala_b200 as gb
from gala_b200.mockstream import DirectNBody, _nbody_dop853, _nbody_leapfrog

pot = gb.MilkyWayPotential2022(); H = gb.Hamiltonian(pot)
prog = np.array([13.0, 0.0, 20.0, 0.0, 130.0 * gb.KMS_TO_KPC_MYR, 50.0 * gb.KMS_TO_KPC_MYR])
pp = gb.PlummerPotential(m=2.5e4, b=float(sys.argv[1]) if len(sys.argv) > 1 else 0.004)
n_steps = 5000
t = np.arange(n_steps + 1) * -1.0


def timed(label, fn, n=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    print(f"{label:60s} {(time.perf_counter() - t0) / n * 1e3:9.2f} ms")
    return r


nb = DirectNBody(prog, [pp], external_potential=pot)
orb = timed("DirectNBody.integrate_orbit (1 body, dense 5001, dop853)", lambda: nb.integrate_orbit(t=t, Integrator="dopri853"))
timed("same through the plain n=6 orbit kernel (massless body)", lambda: DirectNBody(prog, [None], external_potential=pot).integrate_orbit(t=t, Integrator="dopri853"))
po = gb.Orbit(pos=orb.pos[:, ::-1, 0], vel=orb.vel[:, ::-1, 0], t=t[::-1], hamiltonian=H)
s0 = gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)).sample(po, 2.5e4, n_particles=10)
w0 = np.ascontiguousarray(np.vstack([s0.pos, s0.vel]).T)
tm = np.asarray(po.t)
body0 = np.concatenate([po.pos[:, 0], po.vel[:, 0]])[None, :]
traj = timed("bodies at the release times (1 lane, dense)", lambda: _nbody_dop853(H, [pp], body0, tm, tm[-1], tm[1] - tm[0], 0, save_all=True)[2])
unq, nstream = np.unique(s0.release_time, return_counts=True)
group = np.repeat(np.arange(len(unq), dtype=np.int32), nstream)
t1 = np.repeat(unq, nstream)
timed("stream particles, DOP853, lane = [body, particle] (100020 lanes)", lambda: _nbody_dop853(H, [pp], traj, None, tm[-1], tm[1] - tm[0], 1, w0_rows=w0, t1=t1, group=group, err_if_fail=0))
full = timed("bodies over the full grid, leapfrog (1 lane, 5000 steps)", lambda: _nbody_leapfrog(H, [pp], body0, tm[0], tm[-1], n_steps, 1.0, save_all=True)[2])
timed("stream particles, leapfrog, lane = [body, particle]", lambda: _nbody_leapfrog(H, [pp], full, 0.0, tm[-1], 0, 1.0, w0_rows=w0, t1=t1, group=group))
