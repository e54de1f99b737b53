"""cProfile of one MockStreamGenerator.run of the C3 workload (host orchestration vs device time)."""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gala_b200 as gb

integ = sys.argv[1] if len(sys.argv) > 1 else "leapfrog"
sg = len(sys.argv) > 2 and sys.argv[2] == "selfgrav"
H = gb.Hamiltonian(gb.MilkyWayPotential2022())
prog = gb.PhaseSpacePosition(pos=[13.0, 0.0, 20.0], vel=np.array([0.0, 130.0, 50.0]) * gb.KMS_TO_KPC_MYR)


def run():
    gen = gb.MockStreamGenerator(gb.FardalStreamDF(gala_modified=True, random_state=np.random.RandomState(42)), H,
                                 progenitor_potential=gb.PlummerPotential(m=2.5e4, b=0.004) if sg else None)
    return gen.run(prog, 2.5e4, dt=-1.0, n_steps=5000, n_particles=10, release_every=1, Integrator=integ,
                   Integrator_kwargs={"err_if_fail": 0} if integ != "leapfrog" else None)


for _ in range(3):
    run()
t0 = time.perf_counter()
for _ in range(5):
    run()
print(f"{integ} selfgrav={sg}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per run")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    run()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
