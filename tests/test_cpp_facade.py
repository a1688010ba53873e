"""The C++ facade (soft-body-simulator_b200/cpp/sbs): the reference's class names over the C ABI.
tests/cpp/facade_demo.cpp is main.cpp:17-130 of the reference written against it."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
DEMO = os.path.join(BUILD, "facade_demo")
LIBDIR = os.path.join(ROOT, "soft-body-simulator_b200", "lib")


def build_demo():
    src = os.path.join(HERE, "cpp", "facade_demo.cpp")
    hdr = os.path.join(ROOT, "soft-body-simulator_b200", "cpp", "sbs", "b200", "facade.hpp")
    if os.path.exists(DEMO) and all(os.path.getmtime(DEMO) >= os.path.getmtime(f) for f in (src, hdr)):
        return DEMO
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I",
                           os.path.join(ROOT, "soft-body-simulator_b200", "cpp"), src, "-o", DEMO,
                           "-L", LIBDIR, "-lsbsb200", "-Wl,-rpath," + LIBDIR])
    return DEMO


def test_facade_compiles_against_the_reference_include_paths_and_fails_loudly_without_a_gpu(sbs):
    sbs.load_library()
    demo = build_demo()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is covered by the gpu test")
    r = subprocess.run([demo, "4", "4", "12", "1", "1", "5", os.path.join(BUILD, "never.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 10
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    assert not os.path.exists(os.path.join(BUILD, "never.bin"))


@pytest.mark.gpu
@pytest.mark.parametrize("obstacle", ["floor", "mesh"])
@pytest.mark.parametrize("precision", [64, 32])
def test_facade_demo_matches_the_c_abi_path_and_the_oracle(sbs, scenes, oracle, precision, obstacle):
    demo = build_demo()
    out = os.path.join(BUILD, "facade_%d_%s.bin" % (precision, obstacle))
    W, H, D, frames, S, K = 4, 4, 12, 3, 2, 5
    subprocess.check_call([demo, str(W), str(H), str(D), str(frames), str(S), str(K), out, str(precision)] +
                          (["mesh"] if obstacle == "mesh" else []))
    rows = np.fromfile(out, np.float64).reshape(-1, 9)
    x0, xd, vd = rows[:, 0:3], rows[:, 3:6], rows[:, 6:9]
    pos, tets = scenes.bar_model(W, H, D)
    assert len(x0) == len(pos)
    # the same scene through the ctypes binding and through the oracle (reference algorithm)
    body = scenes.TetBody(x0=x0.copy(), tets=tets.astype(np.uint32), x=x0.copy())
    floor = scenes.Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (-200.0, -5.0, -200.0, 200.0, 5.0, 200.0))
    scene = scenes.Scene("facade_demo", [body, floor], substeps=S, iterations=K)
    if obstacle == "mesh":  # the octahedron of tests/cpp/facade_demo.cpp, baked by environment_body_t's mesh constructor
        rock_x, rock_f = scenes.octahedron((0.0, 0.0, 3.0), (1.5, 0.75, 3.0))
        scene.items.append(scenes.Sdf("mesh", rock_x, rock_f, (-6.0, -3.0, -6.0, 8.0, 8.0, 30.0), (8, 6, 12)))
    sim = sbs.Simulation(0, precision)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    for f in range(frames):
        for w, b in ((sim, ids[0]), (ref, 0)):
            w.step(scene.dt, S, K, False)
            if f == 0:
                w.set_mass(b, 0, 0.0)
    xs, vs = sim.download(ids[0])
    assert np.array_equal(xs, xd) and np.array_equal(vs, vd)
    xr, _ = ref.download(0)
    tol = 1e-9 if precision == 64 else 1e-4
    assert np.abs(xd - xr).max() <= tol * scene.bbox_diagonal()
    assert np.abs(xd - x0).max() > 1e-3   # it moved


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [64, 32])
def test_facade_constraints_added_and_removed_between_frames(sbs, scenes, oracle, precision):
    """simulation_t::remove_constraint (swap with the last, simulation.cpp:34-39) and add_constraint after the first
    frame: the facade rebuilds the device scene (colouring included) and carries the state over."""
    demo = build_demo()
    out = os.path.join(BUILD, "facade_%d_dynamic.bin" % precision)
    W, H, D, frames, S, K = 4, 4, 12, 3, 2, 5
    subprocess.check_call([demo, str(W), str(H), str(D), str(frames), str(S), str(K), out, str(precision), "dynamic"])
    rows = np.fromfile(out, np.float64).reshape(-1, 9)
    x0, xd, vd = rows[:, 0:3], rows[:, 3:6], rows[:, 6:9]
    _, tets = scenes.bar_model(W, H, D)
    tets = tets.astype(np.uint32)
    floor = scenes.Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (-200.0, -5.0, -200.0, 200.0, 5.0, 200.0))

    def first_frame(world):
        scene = scenes.Scene("facade_demo", [scenes.TetBody(x0=x0.copy(), tets=tets, x=x0.copy()), floor],
                             substeps=S, iterations=K)
        ids = scene.instantiate(world)
        return scene, ids

    def later_frames(world, x, v):
        t2 = tets.copy()
        t2[5] = t2[-1]
        mass = np.ones(len(x0))
        mass[0] = 0.0
        scene = scenes.Scene("facade_demo", [scenes.TetBody(x0=x0.copy(), tets=t2[:-1], x=x.copy(), mass=mass), floor],
                             substeps=S, iterations=K)
        scene.distance.append((0, 0, np.array([[1, len(x0) - 1]], np.uint32), 1e-4, 0.0))
        ids = scene.instantiate(world)
        world.upload(ids[0], x, v)
        return scene, ids

    sim = sbs.Simulation(0, precision)
    scene, ids = first_frame(sim)
    ref = oracle.World()
    first_frame(ref)
    ref.set_constraint_order(sim.constraint_order())
    for w in (sim, ref):
        w.step(scene.dt, S, K, False)
    xs, vs = sim.download(ids[0])
    xr, vr = ref.download(0)
    sim2 = sbs.Simulation(0, precision)
    scene2, ids2 = later_frames(sim2, xs, vs)
    ref2 = oracle.World()
    later_frames(ref2, xr, vr)
    ref2.set_constraint_order(sim2.constraint_order())
    for _ in range(frames - 1):
        for w in (sim2, ref2):
            w.step(scene.dt, S, K, False)
    xs2, vs2 = sim2.download(ids2[0])
    xr2, _ = ref2.download(0)
    assert np.array_equal(xs2, xd) and np.array_equal(vs2, vd)       # facade == C ABI path, bit for bit
    tol = 1e-9 if precision == 64 else 1e-4
    assert np.abs(xd - xr2).max() <= tol * scene.bbox_diagonal()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [64, 32])
def test_facade_constraints_removed_in_place(sbs, scenes, oracle, precision):
    """simulation_t::remove_constraint of Green constraints between frames with nothing added: the facade patches the
    device scene in place (sbsb200_remove_constraints) and keeps simulation.cpp:34-39's numbering (the last constraint
    takes the place of the removed one).  Checked against the reference algorithm run with its own remove_constraint."""
    demo = build_demo()
    out = os.path.join(BUILD, "facade_%d_remove.bin" % precision)
    W, H, D, frames, S, K = 4, 4, 12, 3, 2, 5
    subprocess.check_call([demo, str(W), str(H), str(D), str(frames), str(S), str(K), out, str(precision), "remove"])
    rows = np.fromfile(out, np.float64).reshape(-1, 9)
    x0, xd = rows[:, 0:3], rows[:, 3:6]
    raw = np.fromfile(out + ".order", np.uint32)
    n0, n1 = int(raw[0]), int(raw[1])
    before, after = raw[2:2 + n0], raw[2 + n0:2 + n0 + n1]
    _, tets = scenes.bar_model(W, H, D)
    assert n0 == len(tets) and n1 == n0 - 4 and sorted(after.tolist()) == list(range(n1))
    floor = scenes.Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (-200.0, -5.0, -200.0, 200.0, 5.0, 200.0))
    scene = scenes.Scene("facade_demo", [scenes.TetBody(x0=x0.copy(), tets=tets.astype(np.uint32), x=x0.copy()), floor],
                         substeps=S, iterations=K)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(before)
    labels = before.tolist()                     # position in simulation.constraints() of the constraint in slot j
    ref.step(scene.dt, S, K, False)
    ref.set_mass(0, 0, 0.0)
    for gone in (5, 17, 40, 17):
        last = len(labels) - 1
        j = labels.index(gone)
        ref.remove_constraint(j)                 # the oracle's list: its last slot moves into slot j
        labels[j] = labels[-1]
        labels.pop()
        if gone != last:                         # the facade's list: position `last` became position `gone`
            labels[labels.index(last)] = gone
    ref.set_constraint_order(np.array([labels.index(q) for q in after.tolist()], np.uint32))
    for _ in range(frames - 1):
        ref.step(scene.dt, S, K, False)
    xr, _ = ref.download(0)
    tol = 1e-9 if precision == 64 else 1e-4
    assert np.abs(xd - xr).max() <= tol * scene.bbox_diagonal()


@pytest.mark.gpu
def test_facade_visual_model_for_a_renderer(sbs, scenes):
    """tetrahedral_body_t::visual_model() / update_visual_model() / prepare_*_for_rendering (tetrahedral_body.cpp:85-119,
    :157-165; tetrahedral_mesh_boundary.cpp:170-208) on the facade: the 9-float vertex buffer and the index buffer a
    renderer consumes, filled from the device, equal what the C ABI hands out for the same scene."""
    demo = build_demo()
    out = os.path.join(BUILD, "facade_32_surface.bin")
    W, H, D, frames, S, K = 4, 4, 12, 24, 2, 5          # long enough for the beam to reach the floor
    subprocess.check_call([demo, str(W), str(H), str(D), str(frames), str(S), str(K), out, "32", "surface"])
    rows = np.fromfile(out, np.float64).reshape(-1, 9)
    x0 = rows[:, 0:3]
    raw = np.fromfile(out + ".surface", np.uint32)
    nv, nt = int(raw[0]), int(raw[1])
    vbuf = raw[2:2 + 9 * nv].view(np.float32).reshape(nv, 9)
    ibuf = raw[2 + 9 * nv:2 + 9 * nv + 3 * nt].reshape(nt, 3)
    s2t = raw[2 + 9 * nv + 3 * nt:2 + 9 * nv + 3 * nt + nv]
    _, tets = scenes.bar_model(W, H, D)
    floor = scenes.Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (-200.0, -5.0, -200.0, 200.0, 5.0, 200.0))
    scene = scenes.Scene("facade_demo", [scenes.TetBody(x0=x0.copy(), tets=tets.astype(np.uint32), x=x0.copy()), floor],
                         substeps=S, iterations=K)
    sim = sbs.Simulation(0, 32)
    ids = scene.instantiate(sim)
    for f in range(frames):
        sim.step(scene.dt, S, K, False)
        if f == 0:
            sim.set_mass(ids[0], 0, 0.0)
    assert np.array_equal(s2t, sim.surface_map(ids[0])) and np.array_equal(ibuf, sim.surface_triangles(ids[0]).reshape(-1, 3))
    ref = sim.download_surface_rgb(ids[0], (1.0, 1.0, 0.0))       # the demo's beam_geometry.set_color(255, 255, 0)
    assert np.array_equal(vbuf[:, :3], ref[:, :3]) and np.array_equal(vbuf[:, 6:], ref[:, 6:])
    assert np.abs(vbuf[:, 3:6] - ref[:, 3:6]).max() < 1e-5          # normals: float atomics, order-dependent last bit
    assert np.abs(np.linalg.norm(vbuf[:, 3:6], axis=1) - 1).max() < 1e-5
    # simulation_t::contacts(): what contact_handler_t::handle would have seen at the last detection (contact.h:11-58)
    rows = np.fromfile(out + ".contacts", np.float64).reshape(-1, 9)
    tail, rows = rows[-1], rows[:-1]
    body, vertex, sdf_body, point, normal = sim.contacts()
    assert len(rows) == len(body) > 0
    mine = sorted((int(s2t[int(r[2])]), tuple(np.round(r[3:9], 12))) for r in rows)
    theirs = sorted((int(v), tuple(np.round(np.concatenate([p, n]), 12))) for v, p, n in zip(vertex, point, normal))
    assert mine == theirs
    assert set(rows[:, 0].astype(int)) == {0} and set(rows[:, 1].astype(int)) == {1}     # beam, floor: simulation body indices
    # sdf_model_t::evaluate on the host (plane) and as the device samples it: y = 1.25 above the floor, gradient +y
    assert tail[0] == -1 and abs(tail[1] - 1.25) < 1e-12 and abs(tail[2] - 1.25) < 1e-12
    assert np.allclose(tail[3:6], (0, 1, 0)) and np.allclose(tail[6:9], (0, 1, 0))
