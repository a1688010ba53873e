"""GPU parity: the CUDA path through the C ABI against the CPU oracle run in the exported
colour order (BASELINE.json north_star: fp64 build <= 1e-9 of the bounding-box diagonal,
fp32 build <= 1e-4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {64: 1e-9, 32: 1e-4}


def run_pair(sbs, oracle, scene, precision, frames=1, schedule=0, substeps=None, iterations=None):
    sim = sbs.Simulation(0, precision, schedule=schedule)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    S = scene.substeps if substeps is None else substeps
    K = scene.iterations if iterations is None else iterations
    ref.contact_history = []
    for _ in range(frames):
        sim.step(scene.dt, S, K, scene.detect_every_substep)
        ref.step(scene.dt, S, K, scene.detect_every_substep)
        ref.contact_history.append(len(ref.contacts()[0]))
    out = []
    for b in scene.tet_bodies():
        xg, vg = sim.download(ids[b])
        xr, vr = ref.download(b)
        out.append((xg, vg, xr, vr))
    return sim, ref, out


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
def test_config1_beam_on_floor(sbs, scenes, oracle, precision, schedule):
    scene = scenes.config1()
    sim, ref, out = run_pair(sbs, oracle, scene, precision, frames=2, schedule=schedule)
    assert sim.stats()["schedule"] == schedule, sim.schedule_note()
    diag = scene.bbox_diagonal()
    (xg, vg, xr, vr), = out
    assert np.isfinite(xg).all()
    dev = np.abs(xg - xr).max() / diag
    assert dev <= TOL[precision], "max position deviation %.3e of bbox diagonal" % dev
    # something must actually have happened: projections ran and contacts existed
    projected, early = ref.counters()
    assert projected > 0
    assert max(ref.contact_history) > 0


@pytest.mark.parametrize("precision", [64, 32])
def test_contact_set_matches(sbs, scenes, oracle, precision):
    scene = scenes.config1(W=5, H=4, D=6)
    sim, ref, _ = run_pair(sbs, oracle, scene, precision, frames=1)
    assert len(ref.contacts()[0]) > 0
    gb, gv, gs, gp, gn = sim.contacts()
    rb, rv, rs, rp, rn = ref.contacts()
    gkey = sorted(zip(gb.tolist(), gv.tolist(), gs.tolist()))
    rkey = sorted(zip(rb.tolist(), rv.tolist(), rs.tolist()))
    if precision == 64:
        assert gkey == rkey
        go = np.lexsort((gs, gv, gb))
        ro = np.lexsort((rs, rv, rb))
        np.testing.assert_allclose(gp[go], rp[ro], atol=1e-12)
        np.testing.assert_allclose(gn[go], rn[ro], atol=1e-12)
    else:
        # fp32 may flip vertices within rounding of the surface
        assert len(set(gkey) ^ set(rkey)) <= max(2, len(rkey) // 20)


@pytest.mark.parametrize("schedule", [1, 2])
def test_predict_commit_exact_fp64(sbs, scenes, oracle, schedule):
    """Order-independent stages must match exactly: zero iterations leaves predict + commit."""
    scene = scenes.config1(W=4, H=4, D=5)
    sim, ref, out = run_pair(sbs, oracle, scene, 64, frames=3, iterations=0, schedule=schedule)
    (xg, vg, xr, vr), = out
    assert np.array_equal(xg, xr)
    assert np.array_equal(vg, vr)


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
def test_config2_cantilever_small(sbs, scenes, oracle, precision, schedule):
    scene = scenes.config2(W=7, H=7, D=15)
    sim, ref, out = run_pair(sbs, oracle, scene, precision, frames=1, schedule=schedule)
    (xg, vg, xr, vr), = out
    dev = np.abs(xg - xr).max() / scene.bbox_diagonal()
    assert dev <= TOL[precision], dev
    pinned = np.arange(7 * 7 * 15) % 15 == 0
    assert np.array_equal(xg[pinned], scene.items[0].x[pinned].astype(np.float32 if precision == 32 else np.float64))


import golden_cases as G  # noqa: E402


def _checkers(oracle):
    """The C restatement and, when it travelled with the repo, the reference's own sources."""
    out = [("oracle", oracle.World)]
    from oracle import ref as R
    if R.available():
        out.append(("reference", R.World))
    return out


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("name", G.case_names())
def test_golden_scenes_against_reference_in_gpu_colour_order(sbs, oracle, name, precision, schedule):
    """Every scene behind tests/golden (floor, box and sphere contacts, two bodies joined by
    damped springs, damped Green constraints, pinned vertices): the GPU result vs the reference
    algorithm run with constraints permuted into the GPU's exported colour order."""
    scene, frames, _ = G.load(name)
    sim = sbs.Simulation(0, precision, schedule=schedule)
    ids = scene.instantiate(sim)
    order = sim.constraint_order()
    assert np.array_equal(np.sort(order), np.arange(len(order)))
    for _ in range(frames):
        sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    diag = scene.bbox_diagonal()
    for label, World in _checkers(oracle):
        ref = World()
        scene.instantiate(ref)
        ref.set_constraint_order(order)
        for _ in range(frames):
            ref.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
        for b in scene.tet_bodies():
            xg, vg = sim.download(ids[b])
            xr, vr = ref.download(b)
            dev = np.abs(xg - xr).max() / diag
            assert dev <= TOL[precision], "%s: %s body %d deviates %.3e" % (label, name, b, dev)


def test_surface_map_matches_reference_numbering(sbs, scenes, oracle):
    scene = scenes.config1(W=6, H=5, D=4)
    sim = sbs.Simulation(0, 32)
    ids = scene.instantiate(sim)
    s2t, _ = oracle.boundary_surface(scene.items[0].x0.shape[0], scene.items[0].tets)
    assert np.array_equal(sim.surface_map(ids[0]), s2t)


def test_set_mass_pins_a_vertex_between_frames(sbs, scenes, oracle):
    """main.cpp:158-165 toggles mass 1 <-> 0 between frames."""
    scene = scenes.config1(W=4, H=4, D=6)
    sim = sbs.Simulation(0, 64)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    for w in (sim, ref):
        w.step(scene.dt, 10, 10)
    x1, _ = sim.download(ids[0])
    for w, b in ((sim, ids[0]), (ref, 0)):
        w.set_mass(b, 17, 0.0)
        w.step(scene.dt, 10, 10)
    xg, _ = sim.download(ids[0])
    xr, _ = ref.download(0)
    assert np.abs(xg - xr).max() <= 1e-9 * scene.bbox_diagonal()


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("precision", [64, 32])
def test_dragging_a_pinned_vertex_between_frames(sbs, scenes, oracle, precision, schedule):
    """What a picker does (main.cpp:158-165, src/rendering/pick.cpp): pin a vertex (mass 0) and move it a little
    every frame (sbsb200_set_vertices), with and without a velocity override, in both schedules."""
    scene = scenes.config1(W=4, H=5, D=7, seed=4)
    sim = sbs.Simulation(0, precision, schedule=schedule)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    picked = [3, 88]      # a corner vertex and an interior-face vertex
    for w, b in ((sim, ids[0]), (ref, 0)):
        for v in picked:
            w.set_mass(b, v, 0.0)
    for f in range(4):
        xs, _ = sim.download(ids[0])
        target = xs[picked] + np.array([[0.05, 0.1, 0.0], [0.0, 0.08, -0.04]])
        vel = None if f % 2 == 0 else np.array([[0.5, 0.0, 0.0], [0.0, 0.0, 0.25]])
        sim.set_vertices(ids[0], picked, target, vel)
        xr, vr = ref.download(0)
        xr[picked] = target
        if vel is not None:
            vr[picked] = vel
        ref.upload(0, xr, vr)
        for w in (sim, ref):
            w.step(scene.dt, 5, 6)
    xg, vg = sim.download(ids[0])
    xr, vr = ref.download(0)
    tol = TOL[precision]
    assert np.abs(xg - xr).max() <= tol * scene.bbox_diagonal()
    # a pinned vertex is not moved by constraints or gravity, only by its own velocity (timestep.cpp:35-43)
    assert np.abs(xg[picked] - (target + vel * scene.dt)).max() <= 1e-6
    with pytest.raises(sbs.SbsError):
        sim.set_vertices(ids[0], [10 ** 6], [[0, 0, 0]])
    with pytest.raises(sbs.SbsError):
        sim.set_vertices(ids[1], [0], [[0, 0, 0]])     # the floor is not a tetrahedral body


def test_step_host_round_trip_equals_resident_stepping(sbs, scenes):
    scene = scenes.config1(W=5, H=5, D=7)
    a = sbs.Simulation(0, 32)
    b = sbs.Simulation(0, 32)
    ia = scene.instantiate(a)
    ib = scene.instantiate(b)
    nV = scene.items[0].x0.shape[0]
    xin = scene.items[0].x.copy()
    vin = np.zeros((nV, 3))
    xo = np.empty((nV, 3))
    vo = np.empty((nV, 3))
    for _ in range(2):
        a.step(scene.dt, 10, 10)
        b.step_host(ib[0], xin, vin, scene.dt, 10, 10, False, xo, vo)
        xin[:] = xo
        vin[:] = vo
    xa, va = a.download(ia[0])
    # the host round trip goes through fp64 host arrays of fp32 device values: exact
    assert np.array_equal(xa, xo) and np.array_equal(va, vo)


def test_error_behaviour(sbs, scenes):
    sim = sbs.Simulation(0, 32)
    with pytest.raises(sbs.SbsError):
        sim.step(0.016, 10, 10)                      # step before finalize
    with pytest.raises(sbs.SbsError):
        sim.add_tet_body(np.zeros((4, 3)), np.array([[0, 1, 2, 7]]))   # vertex index out of range
    b = sim.add_tet_body(np.eye(4, 3), np.array([[0, 1, 2, 3]]))
    with pytest.raises(sbs.SbsError):
        sim.add_distance_constraints(b, 5, np.array([[0, 1]]))          # no such body
    sim.finalize()
    with pytest.raises(sbs.SbsError):
        sim.finalize()                               # twice
    with pytest.raises(sbs.SbsError):
        sim.step(0.016, 0, 10)                       # substeps must be positive
    with pytest.raises(sbs.SbsError):
        sim.add_sdf_plane((0, 1, 0), (0, 0, 0), (-1, -1, -1, 1, 1, 1))  # after finalize
    sim.step(0.016, 2, 2)
    x, v = sim.download(b)
    assert np.isfinite(x).all()


def test_empty_scene_and_body_without_tets(sbs):
    sim = sbs.Simulation(0, 32)
    sim.finalize()
    sim.step(0.016, 2, 2)
    sim.synchronize()
    sim2 = sbs.Simulation(0, 32)
    b = sim2.add_tet_body(np.array([[0.0, 1.0, 0.0]]), np.zeros((0, 4), np.uint32))
    sim2.finalize()
    sim2.step(0.016, 1, 1)
    x, v = sim2.download(b)
    np.testing.assert_allclose(x[0], [0.0, 1.0 - 9.81 * 0.016 ** 2, 0.0], rtol=1e-6)


def test_ensemble_of_independent_bodies(sbs, scenes, oracle):
    """config4 in small: 400 bodies -> whole bodies per region (a few of them, so that a colour step fills its
    warps), no vertex shared between regions, no synchronisation between regions at all."""
    scene = scenes.config4(n_bodies=400, W=3, H=3, D=5)
    sim = sbs.Simulation(0, 64, schedule=2)
    ids = scene.instantiate(sim)
    st = sim.stats()
    assert st["schedule"] == 2 and 1 < st["n_regions"] <= 400 and st["n_shared_vertices"] == 0, sim.schedule_note()
    assert st["pulls_per_sweep"] == 0
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    sim.step(scene.dt, 10, 10)
    ref.step(scene.dt, 10, 10)
    assert len(ref.contacts()[0]) > 0
    worst = 0.0
    for b in scene.tet_bodies():
        worst = max(worst, np.abs(sim.download(ids[b])[0] - ref.download(b)[0]).max())
    assert worst <= 1e-9 * scene.bbox_diagonal()


@pytest.mark.parametrize("precision", [32])
def test_full_size_properties_config3(sbs, scenes, precision):
    """BASELINE size (1M tets): properties that need no oracle run — finite state, exported
    order is a permutation, determinism (two contexts give identical bits), momentum sanity."""
    scene = scenes.config3()
    out = []
    for _ in range(2):
        sim = sbs.Simulation(0, precision)
        ids = scene.instantiate(sim)
        order = sim.constraint_order()
        sim.step(scene.dt, scene.substeps, scene.iterations, True)
        x, v = sim.download(ids[0])
        out.append((x, v))
        st = sim.stats()
        sim.close()
    assert np.array_equal(np.sort(order), np.arange(scene.n_tets))
    assert np.isfinite(out[0][0]).all() and np.isfinite(out[0][1]).all()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert st["n_tets"] == 1_000_000 and st["n_surface_vertices"] == 22_002
    # the 10% pre-strain of a 40-wide block relaxes by a few units in one frame, not more
    assert np.abs(out[0][0] - scene.items[0].x).max() < 6.0


def test_surface_output_for_rendering(sbs, scenes, oracle):
    """Boundary positions + normals straight from the device: triangles equal the
    reference's boundary extraction, normals equal the reference's formula on the downloaded positions."""
    scene = scenes.config1(W=5, H=4, D=6)
    sim = sbs.Simulation(0, 32)
    ids = scene.instantiate(sim)
    sim.step(scene.dt, 4, 4)
    s2t, tris = oracle.boundary_surface(scene.items[0].x0.shape[0], scene.items[0].tets)
    assert np.array_equal(sim.surface_triangles(ids[0]), tris)
    out = sim.download_surface(ids[0])
    x, _ = sim.download(ids[0])
    assert np.array_equal(out[:, :3], x[s2t].astype(np.float32))
    p = x[s2t]
    n = np.zeros_like(p)
    fn = np.cross(p[tris[:, 1]] - p[tris[:, 0]], p[tris[:, 2]] - p[tris[:, 0]])
    for k in range(3):
        np.add.at(n, tris[:, k], fn)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    assert np.abs(out[:, 3:] - n).max() < 1e-4
    assert np.abs(np.linalg.norm(out[:, 3:], axis=1) - 1).max() < 1e-5


def test_surface_output_with_colours(sbs, scenes):
    """The reference's 9-float render vertex (prepare_vertices_for_surface_rendering,
    tetrahedral_mesh_boundary.cpp:170-193): position, normal, colour — one colour for the body (geometry_t::set_color)
    or one per surface vertex."""
    scene = scenes.config1(W=4, H=4, D=6)
    sim = sbs.Simulation(0, 32)
    ids = scene.instantiate(sim)
    sim.step(scene.dt, 2, 3)
    plain = sim.download_surface(ids[0])
    one = sim.download_surface_rgb(ids[0], (1.0, 1.0, 0.0))          # main.cpp:21: beam_geometry.set_color(255, 255, 0)
    assert one.shape == (len(plain), 9)
    # (the normals are summed with float atomics: the last bit depends on the order of the triangles)
    same = lambda a: np.array_equal(a[:, :3], plain[:, :3]) and np.abs(a[:, 3:6] - plain[:, 3:]).max() < 1e-5
    assert same(one) and np.array_equal(one[:, 6:], np.tile(np.float32([1, 1, 0]), (len(plain), 1)))
    col = np.random.default_rng(0).random((len(plain), 3)).astype(np.float32)
    each = sim.download_surface_rgb(ids[0], col)
    assert same(each) and np.array_equal(each[:, 6:], col)
    with pytest.raises(sbs.SbsError):
        sim.download_surface_rgb(ids[0], col[:5])
