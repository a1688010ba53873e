"""Development aid: what a substep costs besides the colour steps (config 3): iterations = 0 leaves predict + commit
(+ launch), detection on/off."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sbs = importlib.import_module("soft-body-simulator_b200")
sc = importlib.import_module("soft-body-simulator_b200.scenes")
scene = sc.config3()
for label, K, detect, bp in (("full frame", 10, True, 1), ("no iterations (predict + commit + detection)", 0, True, 1),
                             ("no iterations, detection once per frame", 0, False, 1), ("1 iteration", 1, True, 1),
                             ("full frame, no BVH broadphase", 10, True, 0), ("full frame, detection once per frame", 10, False, 1)):
    scene.broadphase = bp
    sim = sbs.Simulation(0, 32)
    scene.instantiate(sim)
    for f in range(4):
        sim.step(scene.dt, scene.substeps, K, detect)
        sim.synchronize()
    st0 = sim.stats()
    for f in range(5):
        sim.step(scene.dt, scene.substeps, K, detect)
        sim.synchronize()
    st = sim.stats()
    print("%-48s frame %.3f ms, substep kernel %.1f us per launch, contacts %d" % (
        label, st["last_step_ms"], 1e3 * (st["kernel_ms"] - st0["kernel_ms"]) / (st["kernel_launches"] - st0["kernel_launches"]), sim.contact_count()))
    sim.close()
