"""GPU parity on meshes that are NOT lattices: every tet has its own rest shape (no rest-shape
dictionary), clusters have varying sizes and up to 16 distinct vertices (the NVC4 = 4 kernels), regions
are irregular.  Plus BASELINE config 2 at its full size (100k tets)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {64: 1e-9, 32: 1e-4}


def delaunay_body(scenes, seed=11, dims=(13, 7, 7), h=0.5):
    """Delaunay tetrahedralisation of a jittered point grid: irregular valences and shapes, but no slivers
    (a near-degenerate rest tet has a huge DmInv and makes the reference itself amplify rounding noise, so
    no tolerance statement is possible on it — same remark as scenes.config2)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(*[np.arange(d) for d in dims], indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    pts = (g + rng.uniform(-0.3, 0.3, size=g.shape)) * h
    tri = Delaunay(pts)
    tets = tri.simplices.astype(np.int64)
    e = pts[tets[:, :3]] - pts[tets[:, 3:4]]
    vol = np.abs(np.linalg.det(e)) / 6.0
    tets = tets[vol > 0.15 * np.median(vol)]
    used = np.unique(tets)
    remap = -np.ones(len(pts), np.int64)
    remap[used] = np.arange(len(used))
    x0 = pts[used]
    tets = remap[tets].astype(np.uint32)
    x = x0 * np.array([1.06, 0.97, 1.0]) + np.array([0.0, 0.15, 0.0])   # mild pre-strain, just above the floor
    return scenes.TetBody(x0=x0, tets=tets, x=x)


def run_pair(sbs, oracle, scene, precision, schedule, frames=1):
    sim = sbs.Simulation(0, precision, schedule=schedule)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    order = sim.constraint_order()
    assert np.array_equal(np.sort(order), np.arange(len(order)))
    ref.set_constraint_order(order)
    for _ in range(frames):
        sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
        ref.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    return sim, ids, ref


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
def test_delaunay_body_on_a_floor(sbs, scenes, oracle, precision, schedule):
    body = delaunay_body(scenes)
    floor = scenes.Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.3, 0.0), (-100.0, -5.0, -100.0, 100.0, 5.0, 100.0))
    # A short window: on this irregular, stiff mesh the reference's own Gauss-Seidel sweep amplifies a
    # 1e-16 difference about threefold per sweep (measured: 1e-14 after 3 sweeps, 5e-4 after 20 in fp64
    # on both sides), so parity can only be stated over a few sweeps — same remark as scenes.config2.
    scene = scenes.Scene("delaunay", [body, floor], substeps=2, iterations=3)
    sim, ids, ref = run_pair(sbs, oracle, scene, precision, schedule)
    st = sim.stats()
    assert st["schedule"] == schedule, sim.schedule_note()
    assert st["n_green_colours"] > 8                       # not the lattice pattern
    xg, _ = sim.download(ids[0])
    xr, _ = ref.download(0)
    assert np.isfinite(xg).all()
    dev = np.abs(xg - xr).max() / scene.bbox_diagonal()
    assert dev <= TOL[precision], dev
    assert len(ref.contacts()[0]) > 0
    assert np.abs(xg - body.x).max() > 1e-3


@pytest.mark.parametrize("schedule", [1, 2])
def test_lattice_with_jittered_rest_shape_has_no_dictionary(sbs, scenes, oracle, schedule):
    """Same topology as the lattices, but every tet's DmInv differs: the streaming kernels run."""
    body = scenes.prestrained_bar(9, 9, 25, seed=5, prestrain=(1.05, 0.97, 1.0))
    rng = np.random.default_rng(3)
    body.x0 = body.x0 + rng.uniform(-0.05, 0.05, size=body.x0.shape)
    floor = scenes.Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.1, 0.0), (-100.0, -5.0, -100.0, 100.0, 5.0, 100.0))
    scene = scenes.Scene("jittered_rest", [body, floor], substeps=5, iterations=5)
    sim, ids, ref = run_pair(sbs, oracle, scene, 64, schedule)
    xg, _ = sim.download(ids[0])
    xr, _ = ref.download(0)
    assert np.abs(xg - xr).max() <= 1e-9 * scene.bbox_diagonal()


def test_config2_at_full_size_fp32(sbs, scenes, oracle):
    """BASELINE configs[1]: the 100k-tet cantilever, fp32 build, against the reference algorithm run in the
    exported colour order (one frame = 10 substeps x 10 iterations = 1e7 projections on the CPU)."""
    scene = scenes.config2()
    sim, ids, ref = run_pair(sbs, oracle, scene, 32, 0)
    assert sim.stats()["n_tets"] == 100_000
    xg, _ = sim.download(ids[0])
    xr, _ = ref.download(0)
    dev = np.abs(xg - xr).max() / scene.bbox_diagonal()
    assert dev <= 1e-4, dev


def test_config3_at_full_size_one_substep_fp32(sbs, scenes, oracle):
    """BASELINE configs[2], the bench workload (1M tets on sphere + floor, BVH broadphase, 148 regions, six
    warps per region with the clusters that exchange vertices on the warps that have a sub-partition to
    themselves): ONE substep (detection + 10 iterations = 1e7 projections on the CPU) against the reference
    algorithm run in the exported colour order, fp32 build."""
    scene = scenes.config3()
    scene.dt = scene.dt / scene.substeps
    scene.substeps = 1
    sim, ids, ref = run_pair(sbs, oracle, scene, 32, 0)
    st = sim.stats()
    assert st["n_tets"] == 1_000_000 and st["schedule"] == sbs.SCHED_PERSISTENT and st["n_regions"] > 100
    xg, vg = sim.download(ids[0])
    xr, vr = ref.download(0)
    dev = np.abs(xg - xr).max() / scene.bbox_diagonal()
    assert dev <= 1e-4, dev
    assert len(sim.contacts()[0]) == len(ref.contacts()[0]) > 0
    assert np.abs(xg - scene.items[0].x).max() > 1e-3        # something moved

