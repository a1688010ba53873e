"""One scene decomposed over two GPUs (SURVEY §8e, config 5 in small): every rank runs its block of
regions, shared vertices travel through mailboxes in peer memory (stores over NVLink issued by the
substep kernel).  Needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


def run_decomposed(sbs, scene, precision, world, frames, devices):
    sims = []
    for r in range(world):
        sim = sbs.Simulation(devices[r], precision, schedule=sbs.SCHED_PERSISTENT)
        ids = scene.instantiate(sim, partition=(r, world))
        sims.append((sim, ids))
    for r, (sim, _) in enumerate(sims):
        for q, (peer, _) in enumerate(sims):
            if q != r:
                sim.connect_peer_context(q, peer)
    for _ in range(frames):
        for sim, _ in sims:                      # asynchronous: the kernels of all ranks run together
            sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    for sim, _ in sims:
        sim.synchronize()
    return sims


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
@pytest.mark.parametrize("precision", [64, 32])
def test_decomposed_body_matches_the_reference_in_the_exported_order(sbs, scenes, oracle, precision):
    scene = scenes.config1(W=9, H=9, D=41, bottom=0.0)     # 12 800 tets -> 2 regions per rank
    world, frames = 2, 1
    sims = run_decomposed(sbs, scene, precision, world, frames, devices=[0, 1])
    orders = [sim.constraint_order() for sim, _ in sims]
    assert np.array_equal(orders[0], orders[1])
    st = sims[0][0].stats()
    assert st["schedule"] == 2 and st["n_regions"] >= 2 * world, sims[0][0].schedule_note()
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(orders[0])
    for _ in range(frames):
        ref.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    xr, vr = ref.download(0)
    ranks = sims[0][0].vertex_ranks(sims[0][1][0])
    assert set(np.unique(ranks)) == {0, 1}
    x = np.empty_like(xr)
    v = np.empty_like(vr)
    for r, (sim, ids) in enumerate(sims):
        xs, vs = sim.download(ids[0])
        x[ranks == r] = xs[ranks == r]
        v[ranks == r] = vs[ranks == r]
    tol = 1e-9 if precision == 64 else 1e-4
    dev = np.abs(x - xr).max() / scene.bbox_diagonal()
    assert dev <= tol, dev
    assert len(ref.contacts()[0]) > 0


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_stepping_before_the_peers_are_connected_is_an_error(sbs, scenes):
    scene = scenes.config1(W=9, H=9, D=41)
    sim = sbs.Simulation(0, 32, schedule=sbs.SCHED_PERSISTENT)
    scene.instantiate(sim, partition=(0, 2))
    with pytest.raises(sbs.SbsError):
        sim.step(scene.dt, 1, 1)


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_owned_vertices_round_trip_through_the_host(sbs, scenes):
    """sbsb200_step_host_vertices_f32 on the ranks of a decomposed body: every rank uploads and downloads only the
    vertices it owns; two frames that way equal two frames with the state left on the devices."""
    scene = scenes.config1(W=9, H=9, D=41)
    ref = run_decomposed(sbs, scene, 32, 2, 2, devices=[0, 1])
    sims = []
    for r in range(2):
        sim = sbs.Simulation(r, 32, schedule=sbs.SCHED_PERSISTENT)
        sims.append((sim, scene.instantiate(sim, partition=(r, 2))))
    sims[0][0].connect_peer_context(1, sims[1][0])
    sims[1][0].connect_peer_context(0, sims[0][0])
    ranks = sims[0][0].vertex_ranks(sims[0][1][0])
    owned = [np.ascontiguousarray(np.nonzero(ranks == r)[0], np.uint32) for r in range(2)]
    x = [scene.items[0].x.astype(np.float32)[o] for o in owned]
    v = [np.zeros_like(a) for a in x]
    import threading
    for _ in range(2):
        out = [(np.empty_like(x[r]), np.empty_like(v[r])) for r in range(2)]
        # the call blocks until the rank's frame is done, and the ranks wait for each other: one host thread per rank
        ts = [threading.Thread(target=lambda r=r: sims[r][0].step_host_vertices_f32(
            sims[r][1][0], owned[r], x[r], v[r], scene.dt, scene.substeps, scene.iterations, False, out[r][0], out[r][1]))
            for r in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        x = [o[0] for o in out]
        v = [o[1] for o in out]
    for r in range(2):
        xr, vr = ref[r][0].download(ref[r][1][0])
        assert np.array_equal(x[r], xr[owned[r]].astype(np.float32))
        assert np.array_equal(v[r], vr[owned[r]].astype(np.float32))


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_constraints_removed_in_place_on_every_rank(sbs, scenes, oracle):
    """sbsb200_remove_constraints on a decomposed body: every rank takes the same constraints out of its copy of the
    scene; the exchange plan stays as it is.  Against the reference algorithm with its own remove_constraint."""
    scene = scenes.config1(W=9, H=9, D=41, bottom=0.0)
    sims = run_decomposed(sbs, scene, 32, 2, 1, devices=[0, 1])
    order0 = sims[0][0].constraint_order()
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(order0)
    labels = [int(i) for i in order0]
    ref.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    n = sims[0][0].constraint_count()
    gone = np.random.default_rng(9).choice(n, n // 20, replace=False).astype(np.uint32)
    for sim, _ in sims:
        sim.remove_constraints(gone)
    order1 = sims[0][0].constraint_order()
    assert np.array_equal(order1, sims[1][0].constraint_order()) and len(order1) == n - len(gone)
    for g in gone.tolist():
        j = labels.index(g)
        ref.remove_constraint(j)
        labels[j] = labels[-1]
        labels.pop()
    ref.set_constraint_order(np.array([labels.index(i) for i in order1.tolist()], np.uint32))
    for sim, _ in sims:
        sim.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    for sim, _ in sims:
        sim.synchronize()
    ref.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
    xr, _ = ref.download(0)
    ranks = sims[0][0].vertex_ranks(sims[0][1][0])
    x = np.empty_like(xr)
    for r, (sim, ids) in enumerate(sims):
        xs, _ = sim.download(ids[0])
        x[ranks == r] = xs[ranks == r]
    assert np.abs(x - xr).max() <= 1e-4 * scene.bbox_diagonal()
