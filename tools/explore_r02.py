"""Development aid (round 2): contact counts per frame of config-3 variants.  python tools/explore_r02.py"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sbs = importlib.import_module("soft-body-simulator_b200")
sc = importlib.import_module("soft-body-simulator_b200.scenes")


def run(scene, frames, label):
    sim = sbs.Simulation(0, 32)
    ids = scene.instantiate(sim)
    out = []
    for f in range(frames):
        sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
        sim.synchronize()
        st = sim.stats()
        out.append((st["last_step_ms"], len(sim.contacts()[0])))
    x, v = sim.download(ids[0])
    print("%s: ms %s  contacts %s  |v|max %.2f  ymin %.3f finite %s" % (
        label, " ".join("%.2f" % a for a, _ in out), " ".join(str(b) for _, b in out),
        np.abs(v).max(), x[:, 1].min(), np.isfinite(x).all()), flush=True)
    sim.close()


for kw in (dict(), dict(vy=0.0), dict(vy=-5.0), dict(vy=-2.0, floor_gap=0.1), dict(vy=-1.0, floor_gap=0.05),
           dict(vy=0.0, floor_gap=0.0), dict(vy=-2.0, floor_gap=60.0), dict(vy=-2.0, prestrain=(1.02, 0.99, 1.0))):
    run(sc.config3(**kw), 24, "config3 %s" % kw)
