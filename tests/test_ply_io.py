"""sbs/io/ply.h of the C++ facade against the reference's own PLY fixtures (copied as golden vectors
under tests/golden/ply/ from /root/reference/data/meshes) and its mesh generator."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "ply")
BUILD = os.path.join(HERE, "_build")
TOOL = os.path.join(BUILD, "ply_tool")


@pytest.fixture(scope="module")
def tool():
    src = os.path.join(HERE, "cpp", "ply_tool.cpp")
    hdrs = [os.path.join(ROOT, "soft-body-simulator_b200", "cpp", "sbs", "io", "ply.h"),
            os.path.join(ROOT, "soft-body-simulator_b200", "cpp", "sbs", "b200", "facade.hpp")]
    os.makedirs(BUILD, exist_ok=True)
    if not os.path.exists(TOOL) or any(os.path.getmtime(TOOL) < os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I",
                               os.path.join(ROOT, "soft-body-simulator_b200", "cpp"), src, "-o", TOOL])
    return TOOL


def dump(tool, path):
    out = subprocess.check_output([tool, "dump", path], text=True).split("\n")
    if out[0] == "none":
        return None
    kind, nv, ni, nc, nn = out[0].split()
    pos = np.array(out[1].split(), float).reshape(-1, 3)
    idx = np.array(out[2].split(), int)
    col = np.array(out[3].split(), int)
    return kind, int(nv), pos, idx, col


@pytest.mark.parametrize("dims,fixture", [((2, 2, 2), "cube_tet.ply"), ((5, 2, 2), "tet_bar_5x2x2.ply")])
def test_generator_plus_ascii_writer_reproduce_the_reference_fixtures_byte_for_byte(tool, tmp_path, dims, fixture):
    out = tmp_path / "bar.ply"
    subprocess.check_call([tool, "bar", *map(str, dims), str(out)])
    assert out.read_bytes() == open(os.path.join(GOLD, fixture), "rb").read()


def test_reader_on_the_reference_fixtures(tool):
    kats = np.load(os.path.join(HERE, "golden", "mesh_kats.npz"))
    for name in ("cube_tet", "tet_bar_5x2x2", "tetrahedron", "2tets"):
        kind, nv, pos, idx, col = dump(tool, os.path.join(GOLD, name + ".ply"))
        assert kind == "tet"
        assert np.array_equal(pos.astype(np.float32), kats[name + "_pos"])
        assert np.array_equal(idx.reshape(-1, 4), kats[name + "_tets"])
    kind, nv, pos, idx, col = dump(tool, os.path.join(GOLD, "cube.ply"))       # ascii, per-vertex colours
    assert kind == "tri" and nv == 8 and len(idx) == 36 and len(col) == 24 and col[0] == 150
    kb, nvb, posb, idxb, colb = dump(tool, os.path.join(GOLD, "cube_bin.ply"))  # binary little endian
    assert kb == "tri" and nvb == 8 and len(idxb) == 36
    assert idxb.min() == 0 and idxb.max() == 7 and np.array_equal(np.unique(posb), [0.0, 1.0])


@pytest.mark.parametrize("src", ["cube_tet.ply", "cube.ply"])
def test_round_trips_through_all_three_formats(tool, tmp_path, src):
    a = os.path.join(GOLD, src)
    le, be, back = tmp_path / "le.ply", tmp_path / "be.ply", tmp_path / "back.ply"
    subprocess.check_call([tool, "convert", a, str(le), "binary_little_endian"])
    subprocess.check_call([tool, "convert", str(le), str(be), "binary_big_endian"])
    subprocess.check_call([tool, "convert", str(be), str(back), "ascii"])
    ref = dump(tool, a)
    for p in (le, be, back):
        got = dump(tool, str(p))
        assert got[0] == ref[0] and got[1] == ref[1]
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3]) and np.array_equal(got[4], ref[4])
    assert le.read_bytes() != be.read_bytes()


def test_quirks(tool, tmp_path):
    # a path that does not end in .ply reads as nothing (src/io/ply.cpp:303-320) ...
    other = tmp_path / "cube_tet.txt"
    other.write_bytes(open(os.path.join(GOLD, "cube_tet.ply"), "rb").read())
    assert dump(tool, str(other)) is None
    # ... header lines that are not understood are skipped, `vertex_indices` is accepted ...
    text = open(os.path.join(GOLD, "cube_tet.ply")).read().replace(
        "format ascii 1.0\n", "format ascii 1.0\ncomment made by hand\nobj_info whatever 1 2 3\n")
    odd = tmp_path / "odd.ply"
    odd.write_text(text.replace("uchar int indices", "uchar int vertex_indices"))
    assert dump(tool, str(odd))[1] == 8
    # ... and a truncated body is an error, not garbage
    cut = tmp_path / "cut.ply"
    cut.write_text("\n".join(text.split("\n")[:14]) + "\n")
    assert dump(tool, str(cut)) is None
