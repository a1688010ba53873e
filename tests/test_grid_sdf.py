"""Discrete-grid SDFs: sdf_model_t's non-analytic path (sdf_model.cpp:71-74) and
environment_body_t's triangle-mesh constructor (environment_body.cpp:12-78).

Discregrid is an un-vendored, un-pinned dependency of the reference: PARITY UNPINNED at that boundary.
What pins the restatements instead:
  * mathematics — the 32 shape functions are a nodal basis, sum to one, and reproduce every polynomial of
    degree <= 3 exactly; the mesh distance equals the analytic distance of a box;
  * two independent restatements (oracle/xpbd_oracle.c and oracle/ref_shim) agreeing, the second one driven
    by the REFERENCE'S OWN environment_body_t / sdf_model_t code (oracle/_ref);
  * a golden scene produced through the reference's own constructor (tests/golden/ref_config1_on_mesh.npz).
CPU tests check the oracle and the product's host-compiled sampling code; GPU tests check the device path
through the C ABI against the oracle (bit-level for the bake's fp64 arithmetic up to rounding, 1e-12 / 1e-5
for the sampling)."""
import ctypes as C

import numpy as np
import pytest

import golden_cases as G

dp = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)


def _poly(co, p):
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    m = np.stack([np.ones_like(x), x, y, z, x * x, y * y, z * z, x * y, y * z, x * z, x ** 3, y ** 3, z ** 3, x * x * y,
                  x * x * z, y * y * x, y * y * z, z * z * x, z * z * y, x * y * z], -1)
    return m @ co


def _poly_grad(co, p, h=1e-6):
    return np.stack([(_poly(co, p + h * np.eye(3)[d]) - _poly(co, p - h * np.eye(3)[d])) / (2 * h) for d in range(3)], -1)


DOMAIN = (np.array([-1.0, 0.0, 2.0]), np.array([2.0, 1.5, 4.0]))


# ---------------------------------------------------------------------------------------------- CPU
def test_shape_functions_are_a_nodal_basis(oracle):
    """N_i(x_j) = delta_ij at the 32 nodes of the reference cell, sum N = 1, sum dN = 0."""
    nodes = oracle.grid_node_positions([-1, -1, -1], [1, 1, 1], [1, 1, 1])
    assert len(nodes) == 32
    # map node index -> shape function index through interpolation of indicator data
    for j in range(32):
        val = np.zeros(32)
        val[j] = 1.0
        v, _ = oracle.grid_interpolate([-1, -1, -1], [1, 1, 1], [1, 1, 1], val, nodes)
        np.testing.assert_allclose(v, val, atol=1e-14)
    rng = np.random.default_rng(3)
    for xi in rng.uniform(-1, 1, size=(50, 3)):
        N, dN = oracle.grid_shape(xi)
        assert abs(N.sum() - 1) < 1e-14 and np.abs(dN.sum(0)).max() < 1e-13


@pytest.mark.parametrize("res", [(1, 1, 1), (3, 2, 4), (5, 7, 2)])
def test_oracle_grid_reproduces_cubic_polynomials(oracle, res):
    rng = np.random.default_rng(0)
    co = rng.normal(size=20)
    lo, hi = DOMAIN
    P = oracle.grid_node_positions(lo, hi, res)
    assert len(P) == oracle.grid_node_count(res) == len(np.unique(np.round(P, 9), axis=0))
    pts = rng.uniform(lo, hi, size=(300, 3))
    pts[:3] = [lo, hi, 0.5 * (lo + hi)]  # corners of the domain are inside (AlignedBox::contains is closed)
    v, g = oracle.grid_interpolate(lo, hi, res, _poly(co, P), pts)
    np.testing.assert_allclose(v, _poly(co, pts), atol=1e-11)
    np.testing.assert_allclose(g, _poly_grad(co, pts), atol=1e-6)
    out, gout = oracle.grid_interpolate(lo, hi, res, _poly(co, P), [hi + 1e-9, lo - [1, 0, 0]])
    assert (out == np.finfo(float).max).all() and (gout == 0).all()


def test_oracle_mesh_distance_of_a_box_is_analytic(oracle, scenes):
    x, f = scenes.box_mesh((0, 0, 0), (1, 2, 0.5))
    rng = np.random.default_rng(1)
    pts = rng.uniform(-1, 2.5, size=(4000, 3))
    c, h = np.array([0.5, 1.0, 0.25]), np.array([0.5, 1.0, 0.25])
    q = np.abs(pts - c) - h
    exact = np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(q.max(1), 0)
    np.testing.assert_allclose(oracle.mesh_signed_distance(x, f, pts), exact, atol=1e-14)


def test_oracle_and_reference_constructor_agree_on_a_baked_mesh(oracle, scenes):
    """environment_body_t(sim, id, geometry, domain, resolution) run by the reference's own code over the
    shim vs the C restatement: extended domain (the reference inflates once per mesh vertex), node values,
    interpolated values and gradients."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    x, f = scenes.octahedron((0.1, -0.2, 0.3), (1.5, 1.0, 0.75))
    dom, res = (-2, -2, -2, 2, 2, 2), (5, 4, 6)
    w = R.World()
    b = w.add_sdf_mesh(x, f, dom, res)
    vol = w.volume(b)
    d2, nodes = oracle.bake_mesh_sdf(x, f, dom, res)
    assert np.array_equal(vol, d2)
    assert vol[3] - vol[0] > 4.0 * (1 + 6 * 2e-3 * np.sqrt(3) * 0.9)  # six inflations, not one
    P = oracle.grid_node_positions(d2[:3], d2[3:], res)
    np.testing.assert_allclose(w.sdf_evaluate(b, P)[0], nodes, atol=1e-13)
    pts = np.random.default_rng(2).uniform(vol[:3], vol[3:], size=(2000, 3))
    sr, gr = w.sdf_evaluate(b, pts)
    so, go = oracle.grid_interpolate(d2[:3], d2[3:], res, nodes, pts)
    np.testing.assert_allclose(sr, so, atol=1e-13)
    np.testing.assert_allclose(gr, go, atol=1e-12)
    inside = (np.abs(pts - [0.1, -0.2, 0.3]) / [1.5, 1.0, 0.75]).sum(1) < 1
    exact_sign = oracle.mesh_signed_distance(x, f, pts) < 0
    assert np.array_equal(exact_sign, inside)


def test_product_node_layout_matches_oracle(sbs, oracle):
    for res in ((3, 2, 4), (1, 1, 1), (5, 7, 2), (10, 10, 10)):
        assert sbs.grid_node_count(res) == oracle.grid_node_count(res)
        lo, hi = DOMAIN
        if res != (10, 10, 10):
            assert np.array_equal(sbs.grid_node_positions(lo, hi, res), oracle.grid_node_positions(lo, hi, res))


@pytest.mark.parametrize("precision,tol", [(64, 1e-12), (32, 2e-5)])
def test_product_sampling_code_on_host_matches_oracle(hostmath, oracle, precision, tol):
    """csrc/grid_sdf.cuh grid_interpolate<R> compiled for the host (tests/host_math.cu)."""
    rng = np.random.default_rng(5)
    res = np.array([4, 3, 5], np.uint32)
    lo, hi = DOMAIN
    nodes = rng.normal(size=oracle.grid_node_count(res))
    pts = rng.uniform(lo - 0.1, hi + 0.1, size=(1500, 3))
    val = np.empty(len(pts))
    grad = np.empty((len(pts), 3))
    fn = hostmath.hostmath_grid_sample_f64 if precision == 64 else hostmath.hostmath_grid_sample_f32
    n_in = fn(res.ctypes.data_as(u32p), lo.ctypes.data_as(dp), hi.ctypes.data_as(dp), nodes.ctypes.data_as(dp),
              len(nodes), len(pts), pts.ctypes.data_as(dp), val.ctypes.data_as(dp), grad.ctypes.data_as(dp))
    vo, go = oracle.grid_interpolate(lo, hi, res, nodes, pts)
    inside = vo < 1e300
    if precision == 64:
        assert n_in == inside.sum() and np.array_equal(val >= 1e300, ~inside)
    both = inside & (val < 1e300)
    assert both.sum() > 1000
    np.testing.assert_allclose(val[both], vo[both], atol=tol * 10)
    np.testing.assert_allclose(grad[both], go[both], atol=tol * 100)


def test_golden_mesh_obstacle_scene_oracle_vs_reference(oracle):
    """The C restatement reproduces the golden produced by the reference's own constructor and solver."""
    scene, frames, gold = G.load("ref_config1_on_mesh")
    state, contacts = G.run_backend(oracle.World(), scene, frames, gold["order"])
    for f in range(frames):
        assert np.array_equal(contacts[f][0], gold["contacts_f%d" % f])
        np.testing.assert_allclose(contacts[f][2], gold["contact_normals_f%d" % f], atol=1e-11)
    assert (gold["contacts_f1"][:, 2] == 2).sum() >= 4  # the grid obstacle is actually hit
    np.testing.assert_allclose(state[0][0], gold["x_b0"], atol=1e-11)


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [(64, 1e-12), (32, 2e-5)])
def test_gpu_grid_sampling_matches_oracle(sbs, scenes, oracle, precision, tol):
    """sbsb200_eval_sdf = sdf_model_t::evaluate on the device, for a grid with random node values."""
    rng = np.random.default_rng(11)
    res = (6, 5, 7)
    lo, hi = DOMAIN
    nodes = rng.normal(size=oracle.grid_node_count(res))
    sim = sbs.Simulation(0, precision)
    bar = scenes.prestrained_bar(3, 3, 3, 1)
    sim.add_tet_body(bar.x0, bar.tets)
    g = sim.add_sdf_grid(lo, hi, res, nodes)
    pl = sim.add_sdf_plane((0, 1, 0), (0, -5, 0), scenes._BIG)
    sim.finalize()
    pts = rng.uniform(lo - 0.05, hi + 0.05, size=(5000, 3))
    pts[:2] = [lo, hi]
    sd, gr = sim.eval_sdf(g, pts)
    so, go = oracle.grid_interpolate(lo, hi, res, nodes, pts)
    inside = so < 1e300
    if precision == 64:
        assert np.array_equal(sd >= 1e300, ~inside)
    both = inside & (sd < 1e300)
    assert both.sum() > 4000
    np.testing.assert_allclose(sd[both], so[both], atol=tol * 10)
    np.testing.assert_allclose(gr[both], go[both], atol=tol * 100)
    assert (gr[~both] == 0).all()
    # the analytic kinds go through the same entry point
    sp, gp = sim.eval_sdf(pl, pts[:10])
    np.testing.assert_allclose(sp, pts[:10, 1] + 5, atol=1e-5)
    assert np.allclose(gp, [0, 1, 0])
    d, r, n = sim.sdf_grid(g)
    assert np.array_equal(d, np.concatenate([lo, hi])) and tuple(r) == res and np.array_equal(n, nodes)
    with pytest.raises(sbs.SbsError):
        sim.eval_sdf(0, pts[:1])  # a tet body


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["octahedron", "box"])
def test_gpu_bake_matches_oracle(sbs, scenes, oracle, shape):
    """sbsb200_add_sdf_mesh: domain extension on the host, mesh distance at every node on the device."""
    if shape == "octahedron":
        x, f = scenes.octahedron((0.1, -0.2, 0.3), (1.5, 1.0, 0.75))
    else:
        x, f = scenes.box_mesh((-1, -0.5, 0), (1, 0.25, 2))
    dom, res = (-2, -2, -2, 2, 2, 2.5), (7, 5, 6)
    sim = sbs.Simulation(0, 32)
    b = sim.add_sdf_mesh(x, f, dom, res)
    d, r, nodes = sim.sdf_grid(b)
    d2, n2 = oracle.bake_mesh_sdf(x, f, dom, res)
    assert np.array_equal(d, d2) and tuple(r) == res
    np.testing.assert_allclose(nodes, n2, atol=1e-13)
    assert (nodes < 0).any() and (nodes > 0).any()
    # default resolution of the reference (environment_body.h:24)
    b2 = sim.add_sdf_mesh(x, f, dom)
    assert tuple(sim.sdf_grid(b2)[1]) == (10, 10, 10)


@pytest.mark.gpu
def test_gpu_grid_rejects_bad_input(sbs, scenes):
    sim = sbs.Simulation(0, 32)
    with pytest.raises(sbs.SbsError):
        sim.add_sdf_grid((0, 0, 0), (1, 1, 1), (2, 2, 2), np.zeros(5))          # wrong node count
    with pytest.raises(sbs.SbsError):
        sim.add_sdf_grid((0, 0, 0), (1, 0, 1), (2, 2, 2), np.zeros(sbs.grid_node_count((2, 2, 2))))  # empty domain
    x, f = scenes.octahedron((0, 0, 0), (1, 1, 1))
    with pytest.raises(sbs.SbsError):
        sim.add_sdf_mesh(x, np.array([[0, 1, 9]], np.uint32), (-1, -1, -1, 1, 1, 1))  # index out of range
    with pytest.raises(sbs.SbsError):
        sim.add_sdf_mesh(x, np.array([[0, 0, 1]], np.uint32), (-1, -1, -1, 1, 1, 1))  # degenerate triangle


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [(64, 1e-9), (32, 1e-4)])
@pytest.mark.parametrize("broadphase", [0, 1])
def test_gpu_body_on_baked_mesh_obstacle(sbs, scenes, oracle, precision, tol, broadphase):
    """A beam landing on a triangle-mesh obstacle (grid SDF baked on the device) and the floor, detection every
    substep, against the oracle in the exported colour order: contact sets equal, positions within tolerance."""
    scene = scenes.config1_on_mesh(W=5, H=4, D=7, seed=3)
    scene.detect_every_substep = True
    scene.broadphase = broadphase
    sim = sbs.Simulation(0, precision)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    hit_grid = 0
    for _ in range(2):
        sim.step(scene.dt, scene.substeps, scene.iterations, True)
        ref.step(scene.dt, scene.substeps, scene.iterations, True)
        bg, vg, sg, pg, ng = sim.contacts()
        br, vr, sr, pr, nr = ref.contacts()
        if precision == 64:
            assert sorted(zip(vg.tolist(), sg.tolist())) == sorted(zip(vr.tolist(), sr.tolist()))
        hit_grid += int((sr == 2).sum())
    assert hit_grid > 0
    xg, _ = sim.download(ids[0])
    xr, _ = ref.download(0)
    assert np.abs(xg - xr).max() / scene.bbox_diagonal() <= tol
