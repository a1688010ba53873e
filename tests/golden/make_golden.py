"""Generate the committed golden fixtures (run in the authoring container only, where
/root/reference is mounted):

  python tests/golden/make_golden.py

1. mesh_kats.npz  — vertex/tet tables parsed from the reference's own PLY fixtures
   (/root/reference/data/meshes/{cube_tet,tet_bar_5x2x2,bar_tet,tetrahedron,2tets,3tets}.ply);
   the first three equal get_simple_bar_model(2,2,2), (5,2,2), (12,4,4).
2. ref_*.npz      — outputs of the REFERENCE'S OWN solver sources (oracle/_ref/libsbsref.so,
   built by oracle/build_ref.sh from /root/reference/src against oracle/ref_shim) on small
   seeded scenes: positions/velocities after N frames, contacts of every detection, the
   surface-vertex map, for a stated constraint order.
Nothing in the tests reads /root/reference; they read these files.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_MESHES = "/root/reference/data/meshes"


def parse_tet_ply(path):
    with open(path) as f:
        tok = f.read().split("\n")
    nv = nt = 0
    i = 0
    while tok[i].strip() != "end_header":
        p = tok[i].split()
        if p[:2] == ["element", "vertex"]:
            nv = int(p[2])
        if p[:2] == ["element", "tet"]:
            nt = int(p[2])
        i += 1
    body = tok[i + 1:]
    pos = np.array([[float(x) for x in body[k].split()[:3]] for k in range(nv)], np.float32)
    tets = np.array([[int(x) for x in body[nv + k].split()[1:5]] for k in range(nt)], np.int32)
    return pos, tets


def mesh_kats():
    out = {}
    for name in ("cube_tet", "tet_bar_5x2x2", "bar_tet", "tetrahedron", "2tets", "3tets"):
        pos, tets = parse_tet_ply(os.path.join(REF_MESHES, name + ".ply"))
        out[name + "_pos"] = pos
        out[name + "_tets"] = tets
    np.savez_compressed(os.path.join(HERE, "mesh_kats.npz"), **out)


def cases(sc):
    """name -> (scene, frames, order_seed or None)"""
    two = sc.config1(W=3, H=3, D=4)
    b2 = sc.prestrained_bar(3, 3, 4, 77, translate=(4.0, 0.3, 0.0))
    two.items.insert(1, b2)                      # bodies: tet, tet, floor
    pairs = np.array([[35, 0], [34, 1], [31, 4], [30, 5]], np.uint32)  # springs between the bodies
    two.distance.append((0, 1, pairs, 1e-4, 1e-3))
    two.name = "two_bodies_springs"
    damped = sc.config1(W=4, H=3, D=5, seed=11)
    damped.items[0].beta = 1e-6
    damped.name = "config1_damped"
    box = sc.config1(W=4, H=4, D=6, seed=5)
    box.items.append(sc.Sdf("box", (0.5, -1.0, 1.0), (2.5, 0.25, 3.5), sc._BIG))
    box.name = "config1_plus_box"
    # the pre-strained block of round 1, 0.25 inside the sphere, floor far below (the fixtures predate the
    # resting-contact default of scenes.config3 and are kept bit for bit)
    old3 = dict(W=9, H=7, D=11, radius=6.0, gap=-0.25, prestrain=tuple(sc._PRESTRAIN), vy=0.0, floor_gap=12.0)
    c3one = sc.config3(**old3)
    c3one.substeps = 1
    c3one.name = "config3_small_single_substep"
    rest = sc.config3(W=9, H=7, D=11, radius=6.0)
    rest.substeps = 2                            # detection at every substep of a short frame: the block is still down
    rest.dt = 0.004
    rest.name = "config3_resting_small"
    return {
        "ref_config3_small_1substep": (c3one, 1, 17),
        "ref_config1_small": (sc.config1(W=5, H=4, D=6), 2, None),
        "ref_config1_small_permuted": (sc.config1(W=5, H=4, D=6), 2, 123),
        "ref_config1_full": (sc.config1(), 1, 9),
        "ref_config2_small": (sc.config2(W=5, H=5, D=9), 2, 5),
        "ref_config3_small": (sc.config3(**old3), 2, 17),
        # the bench scene in small: unstrained block resting on sphere + floor, contacts at every detection
        "ref_config3_resting_small": (rest, 2, 19),
        "ref_two_bodies_springs": (two, 2, 3),
        "ref_config1_damped": (damped, 2, 21),
        "ref_config1_plus_box": (box, 2, 8),
        # triangle-mesh obstacle baked into a grid SDF by the reference's own environment_body_t constructor
        # (environment_body.cpp:12-78) over the shim's Discregrid stand-ins
        "ref_config1_on_mesh": (sc.config1_on_mesh(), 2, 31),
    }


def run_case(World, scene, frames, order_seed):
    w = World()
    scene.instantiate(w)
    n = w.constraint_count()
    order = np.arange(n, dtype=np.uint32)
    if order_seed is not None:
        order = np.random.default_rng(order_seed).permutation(n).astype(np.uint32)
        w.set_constraint_order(order)
    out = {"order": order, "frames": np.int32(frames)}
    for f in range(frames):
        w.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
        b, v, s, p, nr = w.contacts()
        o = np.lexsort((s, v, b))
        out["contacts_f%d" % f] = np.stack([b[o], v[o].astype(np.int64), s[o]], 1).astype(np.int64) if len(b) else np.zeros((0, 3), np.int64)
        out["contact_points_f%d" % f] = p[o]
        out["contact_normals_f%d" % f] = nr[o]
    for b in scene.tet_bodies():
        x, v = w.download(b)
        out["x_b%d" % b] = x
        out["v_b%d" % b] = v
        if hasattr(w, "surface_map"):
            out["surface_map_b%d" % b] = w.surface_map(b)
    return out


def main():
    sc = importlib.import_module("soft-body-simulator_b200.scenes")
    from oracle import ref as R
    R.build()
    only = sys.argv[1:]
    if not only:
        mesh_kats()
    for name, (scene, frames, seed) in cases(sc).items():
        if only and name not in only:
            continue
        out = run_case(R.World, scene, frames, seed)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith(("x_", "contacts"))})


if __name__ == "__main__":
    main()
