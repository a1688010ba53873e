"""Scene files as the front door of the GPU solver: sbs::io::load_scene of the C++
host side (soft-body-simulator_b200/cpp/sbs/io/load_scene.h) against the reference's own loader.

tests/cpp/scene_dump.cpp is ONE program compiled twice — against the reference's headers and sources
(oracle/_ref/ref_scene_dump, built by oracle/build_ref.sh where /root/reference is mounted) and against this
repo's headers.  The reference's output on tests/golden/scenes/*.json is committed next to them (*.dump,
regenerate with `oracle/_ref/ref_scene_dump <scene> x > <scene>.dump`); the product must reproduce it byte
for byte: lights, node order, ids, flags, float-rounded rescale/translate of every position, indices, colours,
skipped bodies (missing asset, not a .ply)."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
SCENES = os.path.join(HERE, "golden", "scenes")
CPP = os.path.join(ROOT, "soft-body-simulator_b200", "cpp")
LIBDIR = os.path.join(ROOT, "soft-body-simulator_b200", "lib")
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_scene_dump")


def _build(name, link):
    src = os.path.join(HERE, "cpp", name + ".cpp")
    exe = os.path.join(BUILD, name)
    deps = [src] + [os.path.join(dp, f) for dp, _, fs in os.walk(CPP) for f in fs]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return exe
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", CPP, src, "-o", exe]
    if link:
        cmd += ["-L", LIBDIR, "-lsbsb200", "-Wl,-rpath," + LIBDIR]
    subprocess.check_call(cmd)
    return exe


@pytest.fixture(scope="module")
def scene_dump():
    return _build("scene_dump", link=False)


@pytest.mark.parametrize("name", ["beam_on_floor", "lights_only"])
def test_loader_reproduces_the_reference_loader_byte_for_byte(scene_dump, name):
    scene = os.path.join(SCENES, name + ".json")
    mine = subprocess.run([scene_dump, scene, "x"], capture_output=True, check=True).stdout
    gold = open(os.path.join(SCENES, name + ".dump"), "rb").read()
    assert mine == gold
    if os.path.exists(REF_DUMP):  # the reference's loader itself, when it was built here
        assert subprocess.run([REF_DUMP, scene, "x"], capture_output=True, check=True).stdout == gold
    if name == "beam_on_floor":
        text = gold.decode()
        assert text.startswith("nodes 5\n")                       # 2 of 7 bodies skipped
        assert 'node rock "one" environment - collideable' in text  # escaped quotes in an id
        assert "node decoration environment - -" in text            # "collideable" absent = false
        assert "node loose cube - physical -" in text
        assert "missing asset" not in text and "not a ply" not in text


def test_loader_returns_an_empty_scene_for_a_bad_path(scene_dump, tmp_path):
    for path in ("/nonexistent/scene.json", os.path.join(HERE, "golden", "mesh_kats.npz")):
        r = subprocess.run([scene_dump, path], capture_output=True, check=True)
        assert r.stdout == b"nodes 0\n"
        if os.path.exists(REF_DUMP):
            assert subprocess.run([REF_DUMP, path], capture_output=True, check=True).stdout == r.stdout


def test_loader_fails_on_malformed_json(scene_dump, tmp_path):
    for i, text in enumerate(['{"lights": ', '{"lights": {"directional": {"direction": {"x": "a"}}}}', "[1, 2,, 3]"]):
        bad = tmp_path / ("bad%d.json" % i)
        bad.write_text(text)
        r = subprocess.run([scene_dump, str(bad)], capture_output=True)
        assert r.returncode != 0 and b"json" in r.stderr  # an exception, as nlohmann::json throws in the reference


def _parse_dump(path):
    """[(id, physical, collideable, positions[n,3], indices, (mass, v) or None)] from a *.dump file."""
    nodes = []
    cur = None
    for line in open(path):
        t = line.split()
        if line.startswith("node "):
            body = line[len("node "):].rstrip("\n").rsplit(" ", 3)
            cur = {"id": body[0], "physical": body[2] == "physical", "collideable": body[3] == "collideable"}
            nodes.append(cur)
        elif t and t[0] == "geometry":
            cur["tet"] = t[1] == "tetrahedron"
        elif t and t[0] == "physics":
            cur["mass"], cur["v"] = float(t[1]), np.array([float(x) for x in t[2:5]])
        elif t and t[0] == "positions":
            cur["x"] = np.array([float.fromhex(x) for x in t[1:]]).reshape(-1, 3)
        elif t and t[0] == "indices":
            cur["idx"] = np.array([int(x) for x in t[1:]], np.uint32)
    return nodes


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [64, 32])
def test_scene_file_end_to_end_on_the_gpu(sbs, scenes, oracle, precision):
    """tests/cpp/scene_demo.cpp: load_scene + factories creating tetrahedral_body_t / environment_body_t (mesh
    constructor: grid SDFs baked on the device) + timestep_t::step, against the oracle stepping the same scene
    built from the reference loader's committed output; the non-collideable cube must fall through the floor."""
    demo = _build("scene_demo", link=True)
    out = os.path.join(BUILD, "scene_demo_%d.bin" % precision)
    frames, S, K = 12, 4, 5
    subprocess.check_call([demo, os.path.join(SCENES, "beam_on_floor.json"), str(frames), str(S), str(K), out,
                           str(precision)])
    rows = np.fromfile(out, np.float64).reshape(-1, 9)
    nodes = _parse_dump(os.path.join(SCENES, "beam_on_floor.dump"))
    items = []
    for n in nodes:   # body order = node order (environment bodies first, then objects)
        if n["physical"]:
            it = scenes.TetBody(x0=n["x"].copy(), tets=n["idx"].reshape(-1, 4), x=n["x"].copy(),
                                mass=np.full(len(n["x"]), n["mass"]))
            it.v0 = np.tile(n["v"], (len(n["x"]), 1))
        else:
            lo, hi = n["x"].min(0) - 1.0, n["x"].max(0) + 1.0
            it = scenes.Sdf("mesh", n["x"], n["idx"].reshape(-1, 3), tuple(lo) + tuple(hi), (8, 8, 8))
        it.collideable = n["collideable"]
        items.append(it)
    scene = scenes.Scene("beam_on_floor", items, substeps=S, iterations=K)
    sim = sbs.Simulation(0, precision)
    ids = scene.instantiate(sim)
    ref = oracle.World()
    scene.instantiate(ref)
    ref.set_constraint_order(sim.constraint_order())
    for w, bodies in ((sim, ids), (ref, list(range(len(items))))):
        for b, it in zip(bodies, items):
            if isinstance(it, scenes.TetBody):
                w.upload(b, it.x, it.v0)
    bar, cube = [i for i, it in enumerate(items) if isinstance(it, scenes.TetBody)]
    hits = 0
    for _ in range(frames):
        sim.step(scene.dt, S, K, False)
        ref.step(scene.dt, S, K, False)
        for w, bodies in ((sim, ids), (ref, list(range(len(items))))):
            cb, _, cs, _, _ = w.contacts()
            assert set(cb.tolist()) <= {bodies[bar]}                 # the cube is not handed to the cd system
            assert set(cs.tolist()) <= {bodies[0], bodies[1]}        # nor is the decoration
        hits += len(ref.contacts()[0])
    assert hits > 0
    tol = (1e-9 if precision == 64 else 1e-4) * scene.bbox_diagonal()
    at = 0
    for i, (b, it) in enumerate(zip(ids, items)):
        if not isinstance(it, scenes.TetBody):
            continue
        xs, vs = sim.download(b)
        xr, _ = ref.download(i)
        xd = rows[at:at + len(xs)]
        at += len(xs)
        assert np.array_equal(xd[:, 0:3], it.x0)
        assert np.array_equal(xd[:, 3:6], xs) and np.array_equal(xd[:, 6:9], vs)   # facade == C ABI path
        assert np.abs(xs - xr).max() <= tol
    assert at == len(rows)
    # the bar (collideable) is held up by the rock it starts in contact with; the loose cube (not collideable)
    # straddles the floor at the start and falls freely through it
    xb, _ = sim.download(ids[bar])
    xc, _ = sim.download(ids[cube])
    fall = 0.5 * 9.81 * (frames * scene.dt) ** 2
    assert xc[:, 1].max() < 0.5 - 0.8 * fall and xb[:, 1].max() > items[bar].x[:, 1].max() - 0.5 - 0.8 * fall
