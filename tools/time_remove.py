"""Development timing: simulation_t::remove_constraint at the bench size (config 3, 1M tets) through
sbsb200_remove_constraints (in place) against what a rebuild of the device scene costs (finalize)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sbs = importlib.import_module("soft-body-simulator_b200")
sc = importlib.import_module("soft-body-simulator_b200.scenes")
scene = sc.config3()
t0 = time.time()
sim = sbs.Simulation(0, 32)
scene.instantiate(sim)
t1 = time.time()
print("scene %s: finalize + upload (what a rebuild costs) %.2f s" % (scene.name, t1 - t0))
for _ in range(2):
    sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
sim.synchronize()
rng = np.random.default_rng(1)
n = sim.constraint_count()
gone = rng.choice(n, 1001, replace=False).astype(np.uint32)
for label, ids in (("1 constraint", gone[:1]), ("1000 constraints", gone[1:])):
    t0 = time.time()
    sim.remove_constraints(ids)
    t1 = time.time()
    sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    sim.synchronize()
    t2 = time.time()
    print("remove %-16s %.3f ms, the frame after it %.3f ms wall (%.3f ms on the device), %d constraints left"
          % (label, 1e3 * (t1 - t0), 1e3 * (t2 - t1), sim.stats()["last_step_ms"], sim.constraint_count()))
print("non-finite values:", sim.count_non_finite() if hasattr(sim, "count_non_finite") else "n/a")
