"""Development aid: per-warp cycle breakdown of the resident kernel's colour steps (sbsb200_debug_trace_steps).
TRACE_STEPS=N REGION_SHAPE=0|1 python tools/trace_steps.py <config>"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N = int(os.environ.get("TRACE_STEPS", "80"))
W = 12
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
st = sim.stats()
print(st)
C = st["n_green_colours"]
full = sim.debug_trace().reshape(-1, N, W, 16).astype(np.float64)
full[full == 0] = np.nan
f = full[:, C:, :, :]                      # skip the first sweep (cold)
nw = int(np.isfinite(f[0, 0, :, 0]).sum())
print("regions %d, steps %d, warps %d" % (f.shape[0], f.shape[1], nw))
f = f[:, :, :nw, :]
t0 = np.nanmin(f[:, :, :, 0], axis=2, keepdims=True)          # step start of the region = first warp out of the barrier
rel = lambda slot: f[:, :, :, slot] - t0[:, :, :]
last_tet = np.nanmax(f[:, :, :, 8:14], axis=3) - t0[:, :, :]
def by_colour(x, name):
    # x: [regions, steps, warps] -> per colour (position in sweep): mean over regions & sweeps, per warp
    print(name)
    for c in range(C):
        v = x[:, c::C, :]
        print("  colour %d: per warp %s | mean %6.0f  max-over-regions %6.0f" % (
            c, " ".join("%6.0f" % np.nanmean(v[:, :, w]) for w in range(nw)), np.nanmean(v),
            np.nanmean(np.nanmax(np.nanmax(v, axis=2), axis=0))))
by_colour(rel(8), "first tet starts (prepare after the barrier + collision constraints of first touches)")
by_colour(last_tet, "last tet done")
by_colour(rel(15), "pushes issued")
by_colour(rel(2), "first poll round answered (pulls for the NEXT step; empty when the warp's lane 0 pulls nothing)")
by_colour(f[:, :, :, 1], "poll rounds")
by_colour(rel(6) , "ready for the barrier... (before the pulls)")
by_colour(rel(7), "out of the barrier = step done")
gap = np.diff(np.nanmin(f[:, :, :, 7], axis=2), axis=1)
print("barrier-to-barrier: mean %.0f p50 %.0f p95 %.0f" % (np.nanmean(gap), np.nanpercentile(gap, 50), np.nanpercentile(gap, 95)))
for c in range(C):
    print("  into colour %d: mean %.0f  max-over-regions %.0f" % ((c + 1) % C, np.nanmean(gap[:, c::C]), np.nanmean(np.nanmax(gap[:, c::C], axis=0))))
g = np.nanmin(f[:, :, :, 3], axis=2)      # globaltimer (ns) at the start of a step, per region
skew = g - np.nanmin(g, axis=0, keepdims=True)
print("start skew between regions (ns): per colour mean of max %s" % [int(np.nanmean(np.nanmax(skew[:, c::C], axis=0))) for c in range(C)])
step_ns = np.diff(np.nanmin(g, axis=0))
print("global step (first region to first region, ns): by colour %s" % [int(np.nanmean(step_ns[c::C])) for c in range(C)])
late = np.nanmean(skew, axis=1)
o = np.argsort(late)
print("regions that start latest on average (ns): %s ; earliest: %s" % ([(int(r), int(late[r])) for r in o[-8:]], [(int(r), int(late[r])) for r in o[:4]]))
