"""Shared by CPU and GPU tests: rebuild the scenes behind tests/golden/ref_*.npz (the scene
definitions live in tests/golden/make_golden.py, the script that produced the files)."""
import importlib
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def scenes_module():
    return importlib.import_module("soft-body-simulator_b200.scenes")


def case_names():
    return sorted(make_golden.cases(scenes_module()).keys())


def load(name):
    scene, frames, seed = make_golden.cases(scenes_module())[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    return scene, frames, gold


def run_backend(world, scene, frames, order):
    """Step `world` like make_golden.run_case did, with the golden's constraint order."""
    scene.instantiate(world)
    if hasattr(world, "set_constraint_order"):
        world.set_constraint_order(order)
    contacts = []
    for _ in range(frames):
        world.step(scene.dt, scene.substeps, scene.iterations, scene.detect_every_substep)
        b, v, s, p, n = world.contacts()
        o = np.lexsort((s, v, b))
        contacts.append((np.stack([b[o], v[o].astype(np.int64), s[o]], 1).astype(np.int64) if len(b)
                         else np.zeros((0, 3), np.int64), p[o], n[o]))
    state = {b: world.download(b) for b in scene.tet_bodies()}
    return state, contacts
