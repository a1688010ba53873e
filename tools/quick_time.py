"""Ad-hoc timing helper (development only): python tools/quick_time.py <config> [precision] [schedule]"""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sbs = importlib.import_module("soft-body-simulator_b200")
sc = importlib.import_module("soft-body-simulator_b200.scenes")
if os.environ.get("SBSB200_LIB"):          # a variant build of the library (development A/B)
    sbs.LIB_PATH = os.environ["SBSB200_LIB"]
cfg = sys.argv[1] if len(sys.argv) > 1 else "config1"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sched = int(sys.argv[3]) if len(sys.argv) > 3 else 0
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 5
t0 = time.time()
scene = {"config1": sc.config1, "config2": sc.config2, "config3": sc.config3, "config5": sc.config5,
         "config4": lambda: sc.config4(int(os.environ.get("NB", "512")))}[cfg]()
t1 = time.time()
sim = sbs.Simulation(0, prec, schedule=sched,
                     region_shape=int(os.environ["REGION_SHAPE"]) if "REGION_SHAPE" in os.environ else None)
scene.instantiate(sim)
t2 = time.time()
print("scene %s: build %.2fs finalize+upload %.2fs" % (scene.name, t1 - t0, t2 - t1), sim.stats())
for f in range(frames):
    sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    sim.synchronize()
    st = sim.stats()
    proj = scene.n_tets * scene.substeps * scene.iterations
    print("frame %d: %.3f ms  %.3f Gproj/s  contacts=%d" % (f, st["last_step_ms"], proj / st["last_step_ms"] / 1e6, len(sim.contacts()[0])))
