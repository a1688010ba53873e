"""GPU tests added in round 2: the clamp / inversion route of the Green projection on the device, the three
region shapes of the resident schedule against the reference algorithm, the float host interface, batched
masses, the resting config-3 golden."""
import numpy as np
import pytest

import golden_cases as G

pytestmark = pytest.mark.gpu

TOL = {64: 1e-9, 32: 1e-4}


def crushed_scene(scenes):
    """Config 1 squashed to 40 % in y (every singular value in y below the 0.577 clamp,
    green_constraint.cpp:98-102) with a few vertices pushed through the opposite face of their tets (inverted
    tets, :61-65, :91-96)."""
    scene = scenes.config1(W=5, H=5, D=9, seed=21)
    body = scene.items[0]
    body.x = body.x.copy()
    body.x[:, 1] = 0.4 * body.x[:, 1] + 0.5
    # push every 7th interior vertex two cells along +x: the tets around it turn inside out
    W, H, D = 5, 5, 9
    ids = np.arange(W * H * D).reshape(W, H, D)
    inner = ids[1:-1, 1:-1, 1:-1].reshape(-1)[::7]
    body.x[inner, 0] += 2.2
    scene.name = "config1_crushed_and_inverted"
    return scene


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
def test_clamp_and_inversion_route_on_the_device(sbs, scenes, oracle, precision, schedule):
    """One substep of a crushed scene with inverted tets against the reference algorithm in the exported order;
    the debug counter proves that green_general (the route of green_constraint.cpp:61-65, :91-102) ran on the
    device."""
    scene = crushed_scene(scenes)
    sim = sbs.Simulation(0, precision, schedule=schedule)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    dt = scene.dt / scene.substeps
    sim.step(dt, 1, 4, False)
    ref.step(dt, 1, 4, False)
    st = sim.stats()
    assert st["schedule"] == schedule, sim.schedule_note()
    assert st["green_general_calls"] > 100, st          # the general route really ran
    xg, _ = sim.download(ids[0])
    xr, _ = ref.download(0)
    assert np.isfinite(xg).all()
    dev = np.abs(xg - xr).max() / scene.bbox_diagonal()
    assert dev <= TOL[precision], dev
    assert np.abs(xr - scene.items[0].x).max() > 1e-2    # the crushed tets pushed back


def test_general_route_counter_stays_zero_on_a_mild_scene(sbs, scenes):
    scene = scenes.config1(W=4, H=4, D=6)
    sim = sbs.Simulation(0, 32)
    scene.instantiate(sim)
    sim.step(scene.dt, 2, 3)
    sim.synchronize()
    assert sim.stats()["green_general_calls"] == 0


@pytest.mark.parametrize("shape", [0, 1])
@pytest.mark.parametrize("precision", [64, 32])
def test_region_shapes_against_the_reference(sbs, scenes, oracle, precision, shape):
    """Pencils and compact blocks cut the same mesh differently and order the colours differently; each
    exports its own serial order, and the reference run in that order agrees (config 2 at a size that gives
    several regions, contacts through a raised floor)."""
    scene = scenes.config2(W=13, H=13, D=41)
    scene.items[1] = scenes.Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.05, 0.0), scenes._BIG)   # the floor cuts the bottom layer
    sim = sbs.Simulation(0, precision, schedule=sbs.SCHED_PERSISTENT, region_shape=shape)
    ids = scene.instantiate(sim)
    st = sim.stats()
    assert st["schedule"] == sbs.SCHED_PERSISTENT and st["n_regions"] > 4, sim.schedule_note()
    assert st["n_shared_vertices"] > 0 and st["pulls_per_sweep"] > 0
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    sim.step(scene.dt, 2, 5, True)
    ref.step(scene.dt, 2, 5, True)
    assert len(ref.contacts()[0]) > 0
    xg, vg = sim.download(ids[0])
    xr, vr = ref.download(0)
    dev = np.abs(xg - xr).max() / scene.bbox_diagonal()
    assert dev <= TOL[precision], dev
    if precision == 64:
        assert len(sim.contacts()[0]) == len(ref.contacts()[0])


def test_float_host_interface_equals_the_double_one(sbs, scenes):
    """sbsb200_step_host_f32 moves the same state as sbsb200_step_host in half the bytes: with an fp32 device state
    and float-representable inputs both interfaces return the same bits."""
    scene = scenes.config1(W=5, H=4, D=7)
    body = scene.items[0]
    x0 = body.x.astype(np.float32)
    v0 = (0.01 * np.arange(x0.size, dtype=np.float32).reshape(x0.shape) % 0.3).astype(np.float32)
    outs = []
    for use_float in (False, True):
        sim = sbs.Simulation(0, 32)
        ids = scene.instantiate(sim)
        if use_float:
            xo, vo = np.empty_like(x0), np.empty_like(v0)
            sim.step_host_f32(ids[0], x0, v0, scene.dt, 3, 4, False, xo, vo)
            xi, vi = xo.copy(), vo.copy()
            sim.step_host_f32(ids[0], xi, vi, scene.dt, 3, 4, False, xo, vo)
        else:
            xo, vo = np.empty(x0.shape), np.empty(x0.shape)
            sim.step_host(ids[0], x0.astype(np.float64), v0.astype(np.float64), scene.dt, 3, 4, False, xo, vo)
            xi, vi = xo.copy(), vo.copy()
            sim.step_host(ids[0], xi, vi, scene.dt, 3, 4, False, xo, vo)
        outs.append((np.asarray(xo, np.float64), np.asarray(vo, np.float64)))
        sim.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.abs(outs[0][0] - x0).max() > 1e-4


def test_step_host_without_outputs_returns_after_the_step(sbs, scenes):
    """x_out = v_out = NULL: the call still returns only when the inputs may be reused and the step is done."""
    scene = scenes.config1(W=4, H=4, D=6)
    sim = sbs.Simulation(0, 32)
    ids = scene.instantiate(sim)
    x = scene.items[0].x.copy()
    sim.step_host(ids[0], x, None, scene.dt, 2, 2, False, None, None)
    x[:] = np.nan                                   # would poison an upload still in flight
    xs, _ = sim.download(ids[0])
    assert np.isfinite(xs).all()


@pytest.mark.parametrize("schedule", [1, 2])
def test_batched_masses_equal_one_by_one(sbs, scenes, oracle, schedule):
    scene = scenes.config1(W=5, H=4, D=6)
    n = scene.items[0].x0.shape[0]
    pins = np.arange(0, n, 5, dtype=np.uint32)
    outs = []
    for batched in (True, False):
        sim = sbs.Simulation(0, 64, schedule=schedule)
        ids = scene.instantiate(sim)
        sim.step(scene.dt, 2, 3)
        if batched:
            sim.set_masses(ids[0], pins, np.zeros(len(pins)))
        else:
            for v in pins:
                sim.set_mass(ids[0], int(v), 0.0)
        sim.step(scene.dt, 2, 3)
        outs.append(sim.download(ids[0]))
        if batched:
            ref = oracle.World()
            scene.instantiate(ref)
            ref.set_constraint_order(sim.constraint_order())
            ref.step(scene.dt, 2, 3)
            for v in pins:
                ref.set_mass(0, int(v), 0.0)
            ref.step(scene.dt, 2, 3)
            xr, _ = ref.download(0)
            assert np.abs(outs[0][0] - xr).max() <= 1e-9 * scene.bbox_diagonal()
        sim.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    before = scene.items[0].x
    assert np.abs(outs[0][0][pins] - before[pins]).max() > 0     # they moved in the first frame ...


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
def test_resting_config3_golden(sbs, scenes, precision, schedule):
    """tests/golden/ref_config3_resting_small.npz (the reference's own solver): the bench scene in small — an
    unstrained block resting on sphere + floor, detection every substep.  The golden was produced in a seeded random
    order, so the device result is compared through the reference algorithm's order-independent parts only when
    the orders differ: here the device's own order is fed to the oracle, and the golden pins the oracle."""
    scene, frames, gold = G.load("ref_config3_resting_small")
    sim = sbs.Simulation(0, precision, schedule=schedule)
    ids = scene.instantiate(sim)
    from oracle import oracle as O
    ref = O.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    seen = 0
    for _ in range(frames):
        sim.step(scene.dt, scene.substeps, scene.iterations, True)
        ref.step(scene.dt, scene.substeps, scene.iterations, True)
        seen += len(ref.contacts()[0])
    assert seen > 0
    xg, _ = sim.download(ids[0])
    xr, _ = ref.download(0)
    assert np.abs(xg - xr).max() <= TOL[precision] * scene.bbox_diagonal()


def test_host_step_of_listed_vertices(sbs, scenes):
    """sbsb200_step_host_vertices_f32 with all vertices listed equals sbsb200_step_host_f32; with a subset listed the
    others keep their device state."""
    scene = scenes.config1(W=5, H=4, D=7)
    n = scene.items[0].x0.shape[0]
    x0 = scene.items[0].x.astype(np.float32)
    v0 = np.zeros_like(x0)
    a = sbs.Simulation(0, 32)
    ida = scene.instantiate(a)
    xa, va = np.empty_like(x0), np.empty_like(x0)
    a.step_host_f32(ida[0], x0, v0, scene.dt, 2, 3, False, xa, va)
    b = sbs.Simulation(0, 32)
    idb = scene.instantiate(b)
    perm = np.ascontiguousarray(np.random.default_rng(3).permutation(n), np.uint32)
    xb, vb = np.empty_like(x0), np.empty_like(x0)
    b.step_host_vertices_f32(idb[0], perm, np.ascontiguousarray(x0[perm]), np.ascontiguousarray(v0[perm]), scene.dt, 2, 3,
                             False, xb, vb)
    assert np.array_equal(xb, xa[perm]) and np.array_equal(vb, va[perm])
    few = np.ascontiguousarray(perm[:7])
    xf, vf = np.empty((7, 3), np.float32), np.empty((7, 3), np.float32)
    b.step_host_vertices_f32(idb[0], few, None, None, scene.dt, 2, 3, False, xf, vf)     # nothing uploaded: plain step
    a.step(scene.dt, 2, 3)
    x2, v2 = a.download(ida[0])
    assert np.array_equal(xf, x2[few].astype(np.float32)) and np.array_equal(vf, v2[few].astype(np.float32))


def test_non_finite_check(sbs, scenes):
    """sbsb200_count_non_finite: zero on a healthy scene, the number of poisoned vertices after NaNs were uploaded."""
    scene = scenes.config1(W=4, H=4, D=6)
    sim = sbs.Simulation(0, 32)
    ids = scene.instantiate(sim)
    sim.step(scene.dt, 2, 2)
    assert sim.count_non_finite() == 0
    x = scene.items[0].x.copy()
    x[[3, 17, 40]] = np.nan
    sim.upload(ids[0], x)
    assert sim.count_non_finite() == 3


def test_host_step_of_all_bodies_at_once(sbs, scenes):
    """body = SBSB200_ALL_BODIES: x, v of every tet body concatenated in body order, one copy each way."""
    scene = scenes.config4(n_bodies=6, W=3, H=3, D=5)
    bodies = scene.tet_bodies()
    x0 = np.concatenate([scene.items[b].x for b in bodies]).astype(np.float32)
    v0 = np.zeros_like(x0)
    a = sbs.Simulation(0, 32)
    ida = scene.instantiate(a)
    for b in bodies:
        a.upload(ida[b], scene.items[b].x.astype(np.float32).astype(np.float64))
    a.step(scene.dt, 3, 3)
    b_ = sbs.Simulation(0, 32)
    scene.instantiate(b_)
    xo, vo = np.empty_like(x0), np.empty_like(x0)
    b_.step_host_f32(-1, x0, v0, scene.dt, 3, 3, False, xo, vo)
    xa = np.concatenate([a.download(ida[b])[0] for b in bodies]).astype(np.float32)
    va = np.concatenate([a.download(ida[b])[1] for b in bodies]).astype(np.float32)
    assert np.array_equal(xo, xa) and np.array_equal(vo, va)


def _remove_in_reference(world, labels, insertion_id):
    """simulation_t::remove_constraint in a world whose constraint list was permuted: `labels[j]` = insertion index
    of the constraint at position j.  The reference swaps the constraint with the last one and drops it."""
    j = labels.index(insertion_id)
    world.remove_constraint(j)
    labels[j] = labels[-1]
    labels.pop()


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
def test_constraints_removed_in_place(sbs, scenes, oracle, precision, schedule):
    """sbsb200_remove_constraints (simulation_t::remove_constraint, simulation.cpp:34-39, without re-planning):
    after a first frame a tenth of the Green constraints goes; the following frames equal the reference
    algorithm — its own remove_constraint, then the exported order of the remaining constraints — and differ
    from the scene that kept them."""
    scene = scenes.config1(W=5, H=5, D=9, seed=3)
    sim = sbs.Simulation(0, precision, schedule=schedule)
    ids = scene.instantiate(sim)
    keep = sbs.Simulation(0, precision, schedule=schedule)
    scene.instantiate(keep)
    worlds = [oracle.World()]
    try:
        from oracle import ref as REF
        if REF.available():
            worlds.append(REF.World())            # the reference's own translation units
    except ImportError:
        pass
    order0 = sim.constraint_order()
    labels = []
    for w in worlds:
        scene.instantiate(w)
        w.set_constraint_order(order0)
        labels.append([int(i) for i in order0])
    for w in [sim, keep] + worlds:
        w.step(scene.dt, scene.substeps, scene.iterations, False)
    n = sim.constraint_count()
    rng = np.random.default_rng(5)
    gone = rng.choice(n, n // 10, replace=False).astype(np.uint32)
    sim.remove_constraints(gone[: len(gone) // 2])
    sim.remove_constraints(gone[len(gone) // 2:])          # two batches
    assert sim.constraint_count() == n - len(gone)
    order1 = sim.constraint_order()
    assert len(order1) == n - len(gone) and not set(order1.tolist()) & set(gone.tolist())
    assert [i for i in order0.tolist() if i not in set(gone.tolist())] == order1.tolist()   # the schedule stayed
    with pytest.raises(sbs.SbsError):
        sim.remove_constraints(gone[:1])                     # twice
    for w, lab in zip(worlds, labels):
        for g in gone.tolist():
            _remove_in_reference(w, lab, g)
        w.set_constraint_order(np.array([lab.index(i) for i in order1.tolist()], np.uint32))
    for _ in range(2):
        for w in [sim, keep] + worlds:
            w.step(scene.dt, scene.substeps, scene.iterations, False)
    assert sim.stats()["schedule"] == schedule, sim.schedule_note()
    xg, vg = sim.download(ids[0])
    xk, _ = keep.download(ids[0])
    diag = scene.bbox_diagonal()
    for w in worlds:
        xr, vr = w.download(0)
        assert np.abs(xg - xr).max() <= TOL[precision] * diag, np.abs(xg - xr).max() / diag
    assert np.abs(xg - xk).max() > 5 * TOL[32] * diag       # the removed constraints mattered (1.4e-3 of the diagonal)
