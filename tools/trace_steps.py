"""Development aid: per-phase cycle breakdown of the persistent kernel's colour steps.
SBSB200_TRACE_STEPS=N python tools/trace_steps.py <config>"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N = int(os.environ.get("TRACE_STEPS", "80"))
sbs = importlib.import_module("soft-body-simulator_b200")
sc = importlib.import_module("soft-body-simulator_b200.scenes")
cfg = sys.argv[1] if len(sys.argv) > 1 else "config3"
scene = getattr(sc, cfg)()
sim = sbs.Simulation(0, 32, schedule=2, trace_steps=N,
                     region_shape=int(os.environ["REGION_SHAPE"]) if "REGION_SHAPE" in os.environ else None)
scene.instantiate(sim)
for _ in range(3):
    sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
sim.synchronize()
print(sim.stats())
full = sim.debug_trace().reshape(-1, N, 16)
t = full[:, :, :8]
t = t[:, 8:, :]                      # skip the first sweep (cold)
valid = full[:, 8:, 0] > 0
def show(name, x):
    x = x[np.isfinite(x)]
    if x.size:
        print("  %-14s mean %8.0f  p50 %8.0f  p95 %8.0f  max %8.0f" % (name, x.mean(), np.percentile(x, 50), np.percentile(x, 95), x.max()))
f = full[:, 8:, :].astype(np.float64)
f[f == 0] = np.nan
print("regions %d, steps %d; cycles per colour step of thread 0 (over regions and steps)" % f.shape[:2])
show("start->tets", f[:, :, 8] - f[:, :, 0])
for j in range(1, 6):
    show("tet%d" % j, f[:, :, 8 + j] - f[:, :, 7 + j])
last = np.nanmax(f[:, :, 8:14], axis=2)
show("write back", f[:, :, 15] - last)
show("->gather next", f[:, :, 4] - f[:, :, 15])
show("first poll", f[:, :, 2] - f[:, :, 4])
show("poll rounds", f[:, :, 1])
show("polls+barrier", f[:, :, 7] - f[:, :, 2])
show("whole step", f[:, :, 7] - f[:, :, 0])
C = sim.stats()["n_green_colours"]
whole = f[:, :, 7] - f[:, :, 0]
print("whole step by position in the sweep (mean over regions; max over regions = what the slowest region needs):")
for c in range(C):
    w = whole[:, c::C]
    print("  colour %d: mean %7.0f  max-over-regions mean %7.0f  tets %6.0f  push %5.0f" % (
        c, np.nanmean(w), np.nanmean(np.nanmax(w, axis=0)),
        np.nanmean((np.nanmax(f[:, c::C, 8:14], axis=2) - f[:, c::C, 8])), np.nanmean(f[:, c::C, 15] - np.nanmax(f[:, c::C, 8:14], axis=2))))
# time between the barriers of consecutive steps of one region = the step as the region experiences it
gap = np.diff(f[:, :, 7], axis=1)
print("barrier-to-barrier: mean %.0f p50 %.0f p95 %.0f" % (np.nanmean(gap), np.nanpercentile(gap, 50), np.nanpercentile(gap, 95)))
for c in range(C):
    print("  into colour %d: %.0f" % ((c + 1) % C, np.nanmean(gap[:, c::C])))
