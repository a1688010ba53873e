"""CPU tests of the product's host side: C-ABI surface, host scene compiler, and the device
math compiled for the host (tests/host_math.cu) against the oracle."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
dp = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


def _newer(target, sources):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in sources)


@pytest.fixture(scope="session")
def hostscene():
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "libhostscene.so")
    src = [os.path.join(HERE, "host_scene.cpp"), os.path.join(ROOT, "soft-body-simulator_b200/csrc/scene_build.cpp"),
           os.path.join(ROOT, "soft-body-simulator_b200/csrc/scene_build.h")]
    if not _newer(lib, src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", lib, src[0], src[1]])
    L = C.CDLL(lib)
    L.hs_boundary.argtypes = [C.c_int64, C.c_int64, u32p, u32p, u32p, i64p]
    L.hs_colour.argtypes = [C.c_int64, C.c_int64, C.c_int, u32p, dp, u32p, i64p, C.c_int, i32p]
    L.hs_regions.argtypes = [C.c_int64, C.c_int64, u32p, dp, C.c_int, i32p, i32p, i32p]
    L.hs_regions.restype = C.c_int64
    L.hs_cluster_plan.argtypes = [C.c_int64, C.c_int64, u32p, dp, C.c_int, C.c_int, C.c_int, C.c_int64, u32p, u32p, i32p, i64p,
                                  i64p]
    L.hs_exchange_emulate.argtypes = [C.c_int64, C.c_int64, u32p, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, i64p]
    return L


def test_c_abi_exports_every_declared_symbol(sbs):
    header = open(os.path.join(ROOT, "include", "sbs_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sbsb200_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 20
    assert sorted(sbs.EXPORTS) == declared
    L = C.CDLL(sbs.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback_create_fails_loudly_without_gpu(sbs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sbs.SbsError) as e:
        sbs.Simulation(0, sbs.FP32)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "soft-body-simulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("test oracle", "").replace("the oracle", "").lower() \
                    or f == "scenes.py", os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "lib", "libsbsb200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "sbsref" not in out


@pytest.mark.parametrize("dims", [(2, 2, 2), (4, 3, 5), (9, 8, 7)])
def test_boundary_matches_oracle_numbering(hostscene, oracle, dims):
    pos, tets = oracle.bar_model(*dims)
    tets = np.ascontiguousarray(tets, np.uint32)
    nV, nT = len(pos), len(tets)
    s2t = np.empty(nV, np.uint32)
    tris = np.empty((4 * nT, 3), np.uint32)
    ntri = C.c_int64(0)
    nvs = hostscene.hs_boundary(nV, nT, tets.ctypes.data_as(u32p), s2t.ctypes.data_as(u32p),
                                tris.ctypes.data_as(u32p), C.byref(ntri))
    exp_s2t, exp_tris = oracle.boundary_surface(nV, tets)
    assert np.array_equal(s2t[:nvs], exp_s2t)
    assert np.array_equal(tris[:ntri.value], exp_tris)
    W, H, D = dims
    assert nvs == W * H * D - max(W - 2, 0) * max(H - 2, 0) * max(D - 2, 0)


def test_boundary_of_irregular_mesh_with_dangling_face(hostscene, oracle):
    # two tets sharing a face + one tet sharing only an edge: non-manifold input
    tets = np.array([[0, 1, 2, 3], [1, 2, 3, 4], [0, 1, 5, 6]], np.uint32)
    s2t = np.empty(7, np.uint32)
    ntri = C.c_int64(0)
    nvs = hostscene.hs_boundary(7, 3, tets.ctypes.data_as(u32p), s2t.ctypes.data_as(u32p), None, C.byref(ntri))
    exp, _ = oracle.boundary_surface(7, tets)
    assert np.array_equal(s2t[:nvs], exp) and ntri.value == 10


@pytest.mark.parametrize("dims", [(3, 3, 3), (8, 8, 16), (21, 21, 51)])
def test_colouring_is_conflict_free_and_a_permutation(hostscene, oracle, dims):
    pos, tets = oracle.bar_model(*dims)
    tets = np.ascontiguousarray(tets, np.uint32)
    x0 = pos.astype(np.float64)
    n = len(tets)
    order = np.empty(n, np.uint32)
    offsets = np.zeros(300, np.int64)
    valid = C.c_int32(0)
    nc = hostscene.hs_colour(len(pos), n, 4, tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp),
                             order.ctypes.data_as(u32p), offsets.ctypes.data_as(i64p), 256, C.byref(valid))
    assert 0 < nc <= 48 and valid.value == 1
    assert np.array_equal(np.sort(order), np.arange(n))
    assert offsets[0] == 0 and offsets[nc] == n and np.all(np.diff(offsets[:nc + 1]) > 0)
    # independent check of vertex-disjointness per colour
    for c in range(nc):
        vs = tets[order[offsets[c]:offsets[c + 1]]].reshape(-1)
        assert len(np.unique(vs)) == len(vs)


@pytest.mark.parametrize("smem", [0, 220 * 1024, 24 * 1024])
@pytest.mark.parametrize("dims,bodies,regions,per_body", [((8, 8, 16), 1, 1, 0), ((8, 8, 16), 1, 148, 0),
                                                          ((21, 21, 51), 1, 148, 0), ((6, 6, 17), 40, 0, 1),
                                                          ((2, 2, 2), 1, 1, 0), ((3, 2, 2), 3, 5, 0)])
def test_clustered_colouring_of_lattices(hostscene, oracle, dims, bodies, regions, per_body, smem):
    """Clusters = lattice cells (5 tets), exactly 8 colours (2x2x2 parity), conflict-free (checked
    by cluster_plan_is_valid inside hs_cluster_plan), serial and storage orders are permutations,
    and the tets of one cell are consecutive in the exported serial order."""
    pos, tets = oracle.bar_model(*dims)
    tets = np.ascontiguousarray(tets, np.uint32)
    x0 = pos.astype(np.float64)
    T = len(tets) * bodies
    serial = np.empty(T, np.uint32)
    storage = np.empty(T, np.uint32)
    treg = np.empty(T, np.int32)
    ncl, mch = C.c_int64(0), C.c_int64(0)
    nc = hostscene.hs_cluster_plan(len(pos), len(tets), tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp), bodies,
                                   regions, per_body, smem, serial.ctypes.data_as(u32p), storage.ctypes.data_as(u32p),
                                   treg.ctypes.data_as(i32p), C.byref(ncl), C.byref(mch))
    n_cells = (dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1)
    assert nc == min(8, 2 ** sum(d > 2 for d in dims))
    assert ncl.value == n_cells * bodies
    assert np.array_equal(np.sort(serial), np.arange(T)) and np.array_equal(np.sort(storage), np.arange(T))
    cells = serial.reshape(-1, 5) // 5
    assert (cells == cells[:, :1]).all()                      # a cell's 5 tets run back to back ...
    assert (np.diff(serial.reshape(-1, 5), axis=1) == 1).all()  # ... in insertion order
    if per_body:   # ensembles: whole bodies per region, a few consecutive ones so that a colour step fills its warps
        body_of_tet = np.repeat(np.arange(bodies), len(tets))
        group = body_of_tet[treg == 0].max() + 1
        assert np.array_equal(treg, body_of_tet // group)
        if smem:
            # 6x6x17: 50 clusters in every colour step of a body once the classes are balanced -> three bodies fill 160
            # threads, when their vertices fit the shared memory on offer
            assert group == min(3, smem // (len(pos) * 16))
        # ensembles: colour classes of equal size inside every body (Kempe chains), so the widest colour step of a
        # region — what a CTA needs threads and registers for — is close to the mean (first-fit: 72 of 400 cells for 6x6x17)
        assert mch.value <= group * (n_cells / 8 + 3), (mch.value, group, n_cells)


@pytest.mark.parametrize("bodies,group,threads", [(4096, 5, 256), (3000, 7, 352), (2048, 7, 352), (1024, 1, 64), (512, 1, 64),
                                                   (40, 1, 64)])
def test_ensemble_regions_are_sized_for_the_least_idle_capacity(hostscene, oracle, bodies, group, threads):
    """Ensembles on 148 SMs: regions are dealt to the resident CTAs in rounds, so the planner picks the number of bodies
    per region that leaves the SMs the least idle capacity — rounds x CTAs per SM x bodies per region
    (scene_build.cpp; measured in profiles/r02_ab_bodies_per_region.txt): 4 096 bodies of 6x6x17 -> five per region
    (3 rounds x 2 CTAs x 5 = 30 for 27.7 bodies per SM), 2 048 -> seven (one 352-thread CTA per SM, 2 rounds = 14 for
    13.8), few bodies per SM -> one per region."""
    pos, tets = oracle.bar_model(6, 6, 17)
    tets = np.ascontiguousarray(tets, np.uint32)
    x0 = pos.astype(np.float64)
    nr, nt = C.c_int32(0), C.c_int32(0)
    hostscene.hs_ensemble_group.argtypes = [C.c_int64, C.c_int64, u32p, dp, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                            C.POINTER(C.c_int32)]
    g = hostscene.hs_ensemble_group(len(pos), len(tets), tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp), bodies, 148,
                                    C.byref(nr), C.byref(nt))
    assert (g, nt.value) == (group, threads), (g, nr.value, nt.value)
    assert nr.value == -(-bodies // group)


def test_clustered_colouring_of_an_irregular_mesh(hostscene):
    """Random Delaunay-like soup: a fan of tets around shared vertices, unequal cluster sizes."""
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 4, size=(200, 3))
    tets = []
    for _ in range(600):
        c = rng.integers(0, 200)
        d = np.linalg.norm(pts - pts[c], axis=1)
        near = np.argsort(d)[:8]
        tets.append(rng.choice(near, 4, replace=False))
    tets = np.ascontiguousarray(np.array(tets), np.uint32)
    T = len(tets)
    serial = np.empty(T, np.uint32)
    storage = np.empty(T, np.uint32)
    treg = np.empty(T, np.int32)
    ncl, mch = C.c_int64(0), C.c_int64(0)
    for regions, smem in ((1, 0), (7, 0), (7, 220 * 1024), (3, 20 * 1024)):
        nc = hostscene.hs_cluster_plan(200, T, tets.ctypes.data_as(u32p), pts.ctypes.data_as(dp), 1, regions, 0,
                                       smem, serial.ctypes.data_as(u32p), storage.ctypes.data_as(u32p),
                                       treg.ctypes.data_as(i32p), C.byref(ncl), C.byref(mch))
        assert nc > 0, "plan must be conflict-free"
        assert np.array_equal(np.sort(serial), np.arange(T))
        assert treg.min() >= 0 and treg.max() < regions


def test_clusters_that_exchange_vertices_start_on_the_least_loaded_sub_partitions(hostscene):
    """Warp w issues on sub-partition w % 4; with W warps the sub-partitions W % 4 .. 3 hold one warp less.
    Part A clusters are numbered first in a step, so the numbering starts at warp W % 4."""
    for nt, rot in ((192, 64), (160, 32), (224, 96), (256, 0), (384, 0), (128, 0), (64, 0), (96, 0), (320, 64)):
        assert hostscene.hs_item_rotation(nt) == rot
        warps = nt // 32
        per_smsp = [len(range(s, warps, 4)) for s in range(4)]
        first_warp = rot // 32
        assert per_smsp[first_warp % 4] == min(p for p in per_smsp if p > 0) or warps < 4


EMU_KEYS = "regions colours xclusters entries shared pulls pushes quiet max_local nt".split()


def _emulate(hostscene, oracle, dims, bodies, regions, world, per_body, iterations, collide, reverse, pencils):
    pos, tets = oracle.bar_model(*dims)
    tets = np.ascontiguousarray(tets, np.uint32)
    x0 = pos.astype(np.float64)
    st = np.zeros(26, np.int64)
    rc = hostscene.hs_exchange_emulate(len(pos), len(tets), tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp), bodies,
                                       regions, world, per_body, iterations, collide, reverse, pencils,
                                       st.ctypes.data_as(i64p))
    return rc, dict(zip(EMU_KEYS, st[:10].tolist())), st[10:18].tolist()


@pytest.mark.parametrize("dims,bodies,regions,world,per_body,collide,pencils", [
    ((9, 9, 25), 1, 8, 1, 0, 0, 1), ((9, 9, 25), 1, 8, 1, 0, 1, 1), ((9, 9, 25), 1, 8, 2, 0, 1, 1),
    ((9, 9, 25), 1, 12, 4, 0, 1, 0), ((9, 9, 25), 1, 16, 8, 0, 1, 1), ((21, 21, 51), 1, 39, 1, 0, 1, 1),
    ((5, 4, 7), 2, 3, 1, 0, 1, 1), ((6, 6, 17), 12, 0, 1, 1, 1, 1), ((4, 4, 6), 1, 1, 1, 0, 1, 1)])
def test_exchange_plan_reproduces_the_serial_sweep(hostscene, oracle, dims, bodies, regions, world, per_body, collide,
                                                   pencils):
    """The resident schedule's exchange protocol (scene_build.cpp build_exchange_plan, xpbd_resident.cuh) emulated
    on the CPU with an order-sensitive integer projection: every region keeps its own vertex table and pulls /
    pushes exactly what the plan says; a pull must find exactly the expected tag, a push must never overwrite a
    record nobody consumed, a routing word must name the rank of the reading region, and the result must equal
    the serial Gauss-Seidel sweep in the exported order — whatever order the regions run in inside a step."""
    for reverse in (0, 1):
        rc, st, by_colour = _emulate(hostscene, oracle, dims, bodies, regions, world, per_body, 3, collide, reverse, pencils)
        assert rc == 0, "emulation failed with code %d" % rc
    if per_body or regions <= 1:
        assert st["shared"] == 0 and st["pulls"] == 0 and st["xclusters"] == 0     # islands exchange nothing
    else:
        assert st["shared"] > 0 and st["pulls"] == st["pushes"] > 0


@pytest.mark.parametrize("dims,bodies,regions,per_body", [((21, 21, 51), 1, 39, 0), ((9, 9, 25), 1, 8, 0), ((6, 6, 17), 24, 0, 1)])
def test_vertex_slots_are_spread_over_the_shared_memory_banks(hostscene, oracle, dims, bodies, regions, per_body):
    """A warp's 128-bit access to the vertex table costs as many wavefronts as the fullest of the eight 16-byte bank
    columns holds distinct slots.  build_exchange_plan numbers the slots of a region so that the accesses of every warp
    instruction (lane = cluster of the step, instruction = tet m / corner) spread over the columns: vertex order of a
    lattice costs 2-3 times the ideal, the chosen order stays within 1.3 times (and the protocol emulation above runs on
    the renumbered tables)."""
    pos, tets = oracle.bar_model(*dims)
    tets = np.ascontiguousarray(tets, np.uint32)
    x0 = pos.astype(np.float64)
    st = np.zeros(26, np.int64)
    rc = hostscene.hs_exchange_emulate(len(pos), len(tets), tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp), bodies,
                                       regions, 1, per_body, 2, 1, 0, 0, st.ctypes.data_as(i64p))
    assert rc == 0
    before, after, ideal = int(st[23]), int(st[24]), int(st[25])
    assert ideal > 0 and after >= ideal
    assert after <= 1.3 * ideal, (before, after, ideal)
    assert before >= 1.6 * ideal, (before, after, ideal)       # what the renumbering is for


def test_pencil_regions_leave_half_of_the_colour_steps_without_exchange(hostscene, oracle):
    """Regions = bundles of cell columns along the shortest axis, colours ordered as a Gray code of the cell
    parities: the steps that flip the parity along the pencil axis pull (practically) nothing from other
    regions — only vertices on the mesh boundary, which fewer colours touch, still do.  Compact regions
    spread the pulls over all eight steps."""
    rc, st, by_colour = _emulate(hostscene, oracle, (21, 21, 51), 1, 39, 1, 0, 2, 0, 0, 1)
    assert rc == 0 and st["colours"] == 8
    quiet = sorted(by_colour)[:4]
    busy = sorted(by_colour)[4:]
    assert max(quiet) * 10 < min(busy), by_colour
    assert [n * 10 < min(busy) for n in by_colour] in ([True, False] * 4, [False, True] * 4)   # every other step
    rc, st2, by_colour2 = _emulate(hostscene, oracle, (21, 21, 51), 1, 39, 1, 0, 2, 0, 0, 0)
    assert rc == 0 and min(by_colour2) * 10 > max(by_colour2)


def test_colouring_reports_capacity_overflow(hostscene):
    # a fan of 40 tets around one edge needs 40 colours
    n = 40
    tets = np.array([[0, 1, 2 + i, 3 + i] for i in range(n)], np.uint32)
    x0 = np.random.default_rng(0).normal(size=(n + 3, 3))
    order = np.empty(n, np.uint32)
    offsets = np.zeros(300, np.int64)
    valid = C.c_int32(0)
    assert hostscene.hs_colour(n + 3, n, 4, tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp),
                               order.ctypes.data_as(u32p), offsets.ctypes.data_as(i64p), 16, C.byref(valid)) < 0
    assert hostscene.hs_colour(n + 3, n, 4, tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp),
                               order.ctypes.data_as(u32p), offsets.ctypes.data_as(i64p), 64, C.byref(valid)) == n


def test_empty_inputs(hostscene):
    order = np.empty(1, np.uint32)
    offsets = np.zeros(4, np.int64)
    valid = C.c_int32(0)
    x0 = np.zeros((1, 3))
    tets = np.zeros((1, 4), np.uint32)
    assert hostscene.hs_colour(0, 0, 4, tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp),
                               order.ctypes.data_as(u32p), offsets.ctypes.data_as(i64p), 256, C.byref(valid)) == 0
    s2t = np.empty(1, np.uint32)
    ntri = C.c_int64(0)
    assert hostscene.hs_boundary(0, 0, tets.ctypes.data_as(u32p), s2t.ctypes.data_as(u32p), None, C.byref(ntri)) == 0


@pytest.mark.parametrize("regions", [1, 8, 37])
def test_region_plan_partitions_tets_and_classifies_vertices(hostscene, oracle, regions):
    pos, tets = oracle.bar_model(9, 9, 17)
    tets = np.ascontiguousarray(tets, np.uint32)
    x0 = pos.astype(np.float64)
    treg = np.empty(len(tets), np.int32)
    vreg = np.empty(len(pos), np.int32)
    nnb = np.zeros(regions, np.int32)
    n_if = hostscene.hs_regions(len(pos), len(tets), tets.ctypes.data_as(u32p), x0.ctypes.data_as(dp), regions,
                                treg.ctypes.data_as(i32p), vreg.ctypes.data_as(i32p), nnb.ctypes.data_as(i32p))
    counts = np.bincount(treg, minlength=regions)
    assert counts.min() >= len(tets) // regions - 1 and counts.max() <= len(tets) // regions + 1
    # a vertex is interior to r iff all its tets are in r
    owner = np.full(len(pos), -2, np.int64)
    for t, r in zip(tets, treg):
        for v in t:
            owner[v] = r if owner[v] in (-2, r) else -1
    assert np.array_equal(np.where(owner >= 0, owner, -1), vreg)
    assert n_if == int((vreg < 0).sum())
    if regions == 1:
        assert n_if == 0 and nnb[0] == 0
    else:
        assert (nnb > 0).all()


def _host_project(fn, xi, xn, w, DmInv, V0, E, nu, alpha, beta, dt, lam):
    xi = np.ascontiguousarray(xi, float).reshape(12).copy()
    xn = np.ascontiguousarray(xn, float).reshape(12)
    w = np.ascontiguousarray(w, float)
    D = np.ascontiguousarray(DmInv, float).reshape(9)
    l = np.array([lam], float)
    mu = E / (2 * (1 + nu))
    la = E * nu / ((1 + nu) * (1 - 2 * nu))
    fn(xi.ctypes.data_as(dp), xn.ctypes.data_as(dp), w.ctypes.data_as(dp), D.ctypes.data_as(dp), V0, mu, la, alpha,
       beta, dt, l.ctypes.data_as(dp))
    return xi.reshape(4, 3), l[0]


def test_device_math_on_host_matches_oracle(hostmath, oracle):
    """The product replaces the general SVD by an eigen-decomposition of F^T F; check that
    algebra (same source the kernels compile) against the oracle's SVD-based projection over
    stretched, compressed-below-clamp, inverted, pinned and damped cases."""
    rng = np.random.default_rng(1)
    worst64 = worst32 = 0.0
    n_inv = n_clamp = 0
    for it in range(4000):
        x0 = rng.normal(size=(4, 3))
        DmInv, V0 = oracle.green_rest_state(x0)
        if abs(V0) < 0.02:
            continue
        mode = it % 4
        A = (np.eye(3) + 0.05 * rng.normal(size=(3, 3)), np.eye(3) + 0.3 * rng.normal(size=(3, 3)),
             np.diag([1.1, 0.95, 1.0]) + 0.02 * rng.normal(size=(3, 3)), rng.normal(size=(3, 3)))[mode]
        Q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        xi = (x0 @ A.T) @ Q.T + rng.normal(size=3)
        xn = xi + 0.01 * rng.normal(size=(4, 3))
        w = rng.uniform(0.5, 2, size=4)
        if it % 9 == 0:
            w[it % 4] = 0
        beta = 0.0 if it % 3 else 1e-3
        lam = -1e-7 * rng.uniform()
        ref_xi, ref_lam, ran, diag = oracle.green_project(xi, xn, w, DmInv, V0, 1e6, 0.3, 1e-4, beta, 0.0016, lam)
        n_inv += diag[6] > 0
        n_clamp += diag[2] < 0.577
        a, al = _host_project(hostmath.hostmath_green_project_f64, xi, xn, w, DmInv, V0, 1e6, 0.3, 1e-4, beta, 0.0016, lam)
        b, bl = _host_project(hostmath.hostmath_green_project_f32, xi, xn, w, DmInv, V0, 1e6, 0.3, 1e-4, beta, 0.0016, lam)
        scale = max(1.0, np.abs(ref_xi - xi).max())
        worst64 = max(worst64, np.abs(a - ref_xi).max() / scale)
        worst32 = max(worst32, np.abs(b - ref_xi).max() / scale)
        assert al == pytest.approx(ref_lam, rel=1e-9, abs=1e-18)
    assert n_inv > 100 and n_clamp > 100
    assert worst64 < 1e-11, worst64
    assert worst32 < 5e-5, worst32


def test_device_math_rest_state_early_out(hostmath, oracle):
    x0 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    DmInv, V0 = oracle.green_rest_state(x0)
    for fn in (hostmath.hostmath_green_project_f64, hostmath.hostmath_green_project_f32):
        out, lam = _host_project(fn, x0, x0, np.ones(4), DmInv, V0, 1e6, 0.3, 1e-4, 0.0, 0.0016, 0.0)
        assert np.array_equal(out, x0) and lam == 0.0
