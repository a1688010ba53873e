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
