"""GPU-built BVH broadphase (soft-body-simulator_b200/csrc/bvh.cuh) against the plain narrowphase
and against the reference's own point_bvh_model_t::collide (bvh_model.cpp:30-100, via oracle/_ref)."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def contacts_key(sim):
    b, v, s, p, n = sim.contacts()
    return sorted(zip(b.tolist(), v.tolist(), s.tolist()))


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("schedule", [1, 2])
def test_bvh_finds_exactly_the_contacts_of_the_plain_narrowphase_inside_the_volume(sbs, scenes, precision, schedule):
    scene = scenes.config3(W=9, H=9, D=13, radius=6.0, gap=-0.4)   # sphere + floor, detection every substep
    out = []
    for mode in (sbs.BROADPHASE_NONE, sbs.BROADPHASE_BVH):
        sim = sbs.Simulation(0, precision, schedule=schedule)
        sim.set_broadphase(mode)
        ids = scene.instantiate(sim)
        sim.step(scene.dt, 3, 4, True)
        out.append((contacts_key(sim), sim.download(ids[0])))
    assert len(out[0][0]) > 0
    assert out[0][0] == out[1][0]
    assert np.array_equal(out[0][1][0], out[1][1][0]) and np.array_equal(out[0][1][1], out[1][1][1])


@pytest.mark.parametrize("schedule", [1, 2])
def test_bvh_leaf_order_kept_over_several_frames_loses_no_contact(sbs, scenes, schedule):
    """The leaves are re-sorted every 8 frames only (eager frames of the resident schedule); the spheres are
    refitted at every detection, so frames that reuse an older order still find exactly the contacts of the plain
    narrowphase while the body moves and deforms."""
    scene = scenes.config3(W=9, H=9, D=13, radius=6.0, gap=-0.4)
    sims = []
    for mode in (sbs.BROADPHASE_NONE, sbs.BROADPHASE_BVH):
        sim = sbs.Simulation(0, 32, schedule=schedule)
        sim.set_broadphase(mode)
        sims.append((sim, scene.instantiate(sim)))
    seen = 0
    for frame in range(5):
        keys, states = [], []
        for sim, ids in sims:
            sim.step(scene.dt, 3, 4, True)
            keys.append(contacts_key(sim))
            states.append(sim.download(ids[0]))
        assert keys[0] == keys[1], "frame %d" % frame
        assert np.array_equal(states[0][0], states[1][0]) and np.array_equal(states[0][1], states[1][1])
        seen += len(keys[0])
    assert seen > 0


def test_bvh_culls_like_the_reference_when_the_body_is_outside_the_sdf_volume(sbs, scenes, oracle):
    """A beam dips below the floor plane, but the plane's englobing volume() box is far away: the root
    sphere neither has its centre inside the SDF nor reaches the box, so the reference's traversal stops at
    the root and reports nothing (bvh_model.cpp:47-64) — whatever the tree looks like."""
    scene = scenes.config1(W=4, H=4, D=6, bottom=-0.3)
    floor = copy.deepcopy(scene.items[1])
    floor.volume = (100.0, -5.0, 100.0, 110.0, 5.0, 110.0)
    scene.items[1] = floor
    found = {}
    for mode in (sbs.BROADPHASE_NONE, sbs.BROADPHASE_BVH):
        sim = sbs.Simulation(0, 64)
        sim.set_broadphase(mode)
        scene.instantiate(sim)
        sim.step(scene.dt, 1, 1, False)
        found[mode] = len(contacts_key(sim))
    assert found[sbs.BROADPHASE_NONE] > 0          # vertices do penetrate the plane
    assert found[sbs.BROADPHASE_BVH] == 0          # ... but the broadphase never reaches them
    from oracle import ref as R
    if R.available():
        w = R.World()
        scene.instantiate(w)
        w.step(scene.dt, 1, 1, False)
        assert len(w.contacts()[0]) == 0


def test_bvh_keeps_the_bodies_of_an_ensemble_apart(sbs, scenes):
    """One radix tree over all bodies, one subtree per body: contacts equal the plain narrowphase."""
    scene = scenes.config4(n_bodies=40, W=3, H=3, D=5)
    keys = []
    for mode in (sbs.BROADPHASE_NONE, sbs.BROADPHASE_BVH):
        sim = sbs.Simulation(0, 32)
        sim.set_broadphase(mode)
        scene.instantiate(sim)
        sim.step(scene.dt, 2, 2, True)
        keys.append(contacts_key(sim))
    assert len(keys[0]) > 0 and keys[0] == keys[1]
