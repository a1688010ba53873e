"""CPU tests (no GPU): the oracle against every pin we have —
the reference's own PLY fixtures, numpy's SVD, the hand-computed single-projection vectors of
SURVEY.md Appendix C, and golden vectors produced by the reference's own solver sources
(tests/golden/ref_*.npz, see tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_cases as G


def test_bar_model_matches_reference_ply_fixtures(oracle, scenes):
    kats = np.load(G.GOLDEN + "/mesh_kats.npz")
    # the fixtures were written with (W,H,D) = (2,2,2), (5,2,2), (12,4,4)
    for name, dims in (("cube_tet", (2, 2, 2)), ("tet_bar_5x2x2", (5, 2, 2)), ("bar_tet", (12, 4, 4))):
        pos, tets = oracle.bar_model(*dims)
        assert np.array_equal(pos, kats[name + "_pos"]), name
        assert np.array_equal(tets, kats[name + "_tets"]), name
        pos2, tets2 = scenes.bar_model(*dims)
        assert np.array_equal(pos2, pos) and np.array_equal(tets2, tets), name


def test_bar_model_first_tet_of_bar_fixture(oracle):
    # data/meshes/bar_tet.ply: first tet "4 4 16 1 0" = (p3,p1,p4,p0) of cell (0,0,0)
    _, tets = oracle.bar_model(12, 4, 4)
    assert tets[0].tolist() == [4, 16, 1, 0]


def test_svd_against_numpy(oracle):
    rng = np.random.default_rng(0)
    for i in range(500):
        F = rng.normal(size=(3, 3))
        if i % 5 == 0:
            F = np.eye(3) + 1e-3 * rng.normal(size=(3, 3))
        if i % 7 == 0:
            F[:, 2] = 2 * F[:, 1]
        U, s, V = oracle.svd3(F)
        assert np.all(np.diff(s) <= 1e-15) and s[2] >= 0
        np.testing.assert_allclose(s, np.linalg.svd(F, compute_uv=False), atol=1e-13)
        np.testing.assert_allclose(U @ np.diag(s) @ V.T, F, atol=1e-13)
        np.testing.assert_allclose(U.T @ U, np.eye(3), atol=1e-13)
        np.testing.assert_allclose(V.T @ V, np.eye(3), atol=1e-13)


# rest shape of data/meshes/tetrahedron.ply, one vertex displaced.
APPENDIX_C = {
    "stretched": ((1.0, 2.0, 0.2), -3.076795489933365e-07,
                  [[0.035304255942184, 0.041871300447051, -0.975531002249217],
                   [1.0, 0.062704683014217, 0.976225691829209],
                   [1.964695744057816, 0.041871300447051, -0.975531002249217],
                   [1.0, 1.853552716091681, 0.174836312669225]]),
    "inverted": ((1.0, -0.9, 0.1), -1.0113244315883548e-06,
                 [[-0.096292665816897, -0.033314194020719, -1.052817596864343],
                  [1.0, -0.105195190532349, 1.081874699691428],
                  [2.096292665816897, -0.033314194020719, -1.052817596864343],
                  [1.0, -0.728176421426213, 0.123760494037259]]),
}


@pytest.mark.parametrize("case", sorted(APPENDIX_C))
def test_single_projection_known_answers(oracle, case):
    kats = np.load(G.GOLDEN + "/mesh_kats.npz")
    x0 = kats["tetrahedron_pos"].astype(np.float64)
    assert kats["tetrahedron_tets"].tolist() == [[0, 1, 2, 3]]
    DmInv, V0 = oracle.green_rest_state(x0)
    assert V0 == pytest.approx(-1.0)
    p3, lam, xi_out = APPENDIX_C[case]
    xi = x0.copy()
    xi[3] = p3
    out, lam_out, ran, diag = oracle.green_project(xi, x0, np.ones(4), DmInv, V0, 1e6, 0.3, 1e-4, 0.0, 0.0016)
    assert ran
    assert lam_out == pytest.approx(lam, rel=1e-12)
    np.testing.assert_allclose(out, np.array(xi_out), atol=2e-15 * 1e1 + 1e-14)
    assert bool(diag[6]) == (case == "inverted")


def test_rest_state_is_an_early_out(oracle):
    # every tet at rest hits S < 1e-20 (green_constraint.cpp:130-131) and must not move
    pos, tets = oracle.bar_model(3, 3, 3)
    w = oracle.World()
    b = w.add_tet_body(pos.astype(np.float64), tets.astype(np.uint32))
    x_before, _ = w.download(b)
    w.step(0.016, 1, 3)
    projected, early = w.counters()
    assert projected == 0 and early == 3 * tets.shape[0]
    x_after, v = w.download(b)
    # free fall only: dy = -9.81 * dt^2
    np.testing.assert_allclose(x_after - x_before, np.tile([0, -9.81 * 0.016 ** 2, 0], (len(pos), 1)), atol=1e-15)


@pytest.mark.parametrize("name", G.case_names())
def test_oracle_matches_reference_solver_golden(oracle, name):
    """The C restatement against outputs of the reference's own sources (oracle/_ref)."""
    scene, frames, gold = G.load(name)
    state, contacts = G.run_backend(oracle.World(), scene, frames, gold["order"])
    diag = scene.bbox_diagonal()
    for b, (x, v) in state.items():
        assert np.abs(x - gold["x_b%d" % b]).max() <= 1e-12 * diag
        assert np.abs(v - gold["v_b%d" % b]).max() <= 1e-9 * max(1.0, np.abs(gold["v_b%d" % b]).max())
        s2t, _ = oracle.boundary_surface(scene.items[b].x0.shape[0], scene.items[b].tets)
        assert np.array_equal(s2t, gold["surface_map_b%d" % b])
    for f, (keys, pts, nrm) in enumerate(contacts):
        assert np.array_equal(keys, gold["contacts_f%d" % f]), "contact set, frame %d" % f
        np.testing.assert_allclose(pts, gold["contact_points_f%d" % f], atol=1e-12)
        np.testing.assert_allclose(nrm, gold["contact_normals_f%d" % f], atol=1e-12)


def test_oracle_matches_live_reference_build_when_present(oracle, scenes):
    """If oracle/_ref/libsbsref.so travelled with the repo, compare live on a fresh seed."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    scene = scenes.config1(W=6, H=5, D=7, seed=4242)
    a, b = oracle.World(), R.World()
    scene.instantiate(a)
    scene.instantiate(b)
    order = np.random.default_rng(1).permutation(a.constraint_count()).astype(np.uint32)
    a.set_constraint_order(order)
    b.set_constraint_order(order)
    for _ in range(2):
        a.step(scene.dt, 10, 10)
        b.step(scene.dt, 10, 10)
    assert np.abs(a.download(0)[0] - b.download(0)[0]).max() <= 1e-12 * scene.bbox_diagonal()


def test_constraint_order_must_be_a_permutation(oracle, scenes):
    scene = scenes.config1(W=3, H=3, D=3)
    w = oracle.World()
    scene.instantiate(w)
    n = w.constraint_count()
    with pytest.raises(RuntimeError):
        w.set_constraint_order(np.zeros(n, np.uint32))
    with pytest.raises(RuntimeError):
        w.set_constraint_order(np.arange(n - 1, dtype=np.uint32))


def test_detect_every_substep_equals_repeated_single_substep_frames(oracle, scenes):
    scene = scenes.config3(W=5, H=4, D=5, radius=4.0, gap=-0.2)
    a, b = oracle.World(), oracle.World()
    scene.instantiate(a)
    scene.instantiate(b)
    a.step(0.016, 4, 3, True)
    for _ in range(4):
        b.step(0.004, 1, 3, False)
    assert np.array_equal(a.download(0)[0], b.download(0)[0])


def test_remove_constraint_matches_the_reference(oracle, scenes):
    """simulation_t::remove_constraint (simulation.cpp:34-39: swap with the last, drop it) in the C restatement against the
    reference's own method through oracle/_ref: a tenth of the constraints goes after the first frame, two more frames."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    scene = scenes.config1(W=5, H=5, D=9, seed=3)
    worlds = [oracle.World(), R.World()]
    for w in worlds:
        scene.instantiate(w)
        w.step(scene.dt, scene.substeps, scene.iterations, False)
    n = worlds[0].constraint_count()
    assert n == worlds[1].constraint_count()
    gone = np.random.default_rng(5).choice(n, n // 10, replace=False)
    for w in worlds:
        for k, g in enumerate(sorted(gone.tolist(), reverse=True)):     # positions stay valid when removed from the back
            w.remove_constraint(g)
        assert w.constraint_count() == n - len(gone)
        for _ in range(2):
            w.step(scene.dt, scene.substeps, scene.iterations, False)
    xa, va = worlds[0].download(0)
    xb, vb = worlds[1].download(0)
    assert np.abs(xa - xb).max() <= 1e-12 * scene.bbox_diagonal()
    keep = oracle.World()
    scene.instantiate(keep)
    for _ in range(3):
        keep.step(scene.dt, scene.substeps, scene.iterations, False)
    assert np.abs(keep.download(0)[0] - xa).max() > 1e-4 * scene.bbox_diagonal()     # the constraints mattered
