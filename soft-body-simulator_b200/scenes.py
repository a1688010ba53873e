"""Synthetic scenes of SURVEY.md §8(d) / BASELINE.json `configs` (host-side, numpy only).

A scene is a plain description (`Scene`) that can be instantiated on any backend exposing
add_tet_body / add_sdf_* / upload — i.e. both the GPU `Simulation` and the test oracle —
so that both sides see bit-identical fp64 inputs.

`bar_model` restates get_simple_bar_model (src/geometry/get_simple_bar_model.cpp:6-122) in
vectorised numpy: vertex id = i*H*D + j*D + k, 5 tets per cell, orientation alternating on
(i+j+k) % 2.
"""
from dataclasses import dataclass, field

import numpy as np

# corner -> (di, dj, dk) of p0..p7 (get_simple_bar_model.cpp:44-51)
_CORNER = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
_ODD = np.array([[1, 0, 5, 2], [5, 2, 7, 6], [7, 0, 5, 4], [2, 0, 7, 3], [5, 0, 7, 2]])   # :61-87
_EVEN = np.array([[3, 1, 4, 0], [6, 1, 3, 2], [4, 1, 6, 5], [6, 3, 4, 7], [3, 1, 6, 4]])  # :89-113


def bar_model(W, H, D):
    """-> (positions float32 [W*H*D, 3], tets int32 [5*(W-1)*(H-1)*(D-1), 4])."""
    i, j, k = np.meshgrid(np.arange(W), np.arange(H), np.arange(D), indexing="ij")
    pos = np.stack([i, j, k], axis=-1).reshape(-1, 3).astype(np.float32)
    ci, cj, ck = np.meshgrid(np.arange(W - 1), np.arange(H - 1), np.arange(D - 1), indexing="ij")
    ci, cj, ck = ci.reshape(-1), cj.reshape(-1), ck.reshape(-1)
    corner = ((ci[:, None] + _CORNER[None, :, 0]) * H + (cj[:, None] + _CORNER[None, :, 1])) * D \
        + (ck[:, None] + _CORNER[None, :, 2])                       # [cells, 8]
    odd = ((ci + cj + ck) % 2 == 1)
    table = np.where(odd[:, None, None], _ODD[None], _EVEN[None])    # [cells, 5, 4]
    tets = np.take_along_axis(corner[:, None, :].repeat(5, axis=1), table, axis=2)
    return pos, tets.reshape(-1, 4).astype(np.int32)


def _uniform01(seed, n):
    """Counter-based uniform doubles in [0, 1): splitmix64 of (seed, counter)."""
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15) \
            + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


_PRESTRAIN = np.array([1.10, 0.95, 1.00])


@dataclass
class TetBody:
    x0: np.ndarray            # [V, 3] float64 rest positions
    tets: np.ndarray          # [T, 4] uint32
    x: np.ndarray             # [V, 3] float64 initial positions (x = xi = xn)
    mass: np.ndarray = None   # [V] float64 or None (=1)
    young: float = 1e6
    poisson: float = 0.3
    alpha: float = 1e-4
    beta: float = 0.0
    v: np.ndarray = None      # [V, 3] float64 initial velocities or None (= 0)


@dataclass
class Sdf:
    kind: str                 # "plane" | "sphere" | "box" | "mesh" (a = positions, b = triangles, volume = grid domain)
    a: tuple
    b: tuple
    volume: tuple             # (min xyz, max xyz)
    res: tuple = None         # "mesh": grid resolution (None = the reference's default 10x10x10)


@dataclass
class Scene:
    name: str
    items: list = field(default_factory=list)          # TetBody / Sdf in body order
    distance: list = field(default_factory=list)       # (b1, b2, pairs[n,2], alpha, beta)
    dt: float = 0.016
    substeps: int = 10
    iterations: int = 10
    detect_every_substep: bool = False
    collision_compliance: float = 1e-8
    broadphase: int = 0                  # 1 = GPU-built BVH (sbsb200_set_broadphase)

    @property
    def n_tets(self):
        return sum(it.tets.shape[0] for it in self.items if isinstance(it, TetBody))

    @property
    def n_vertices(self):
        return sum(it.x0.shape[0] for it in self.items if isinstance(it, TetBody))

    def bbox_diagonal(self):
        pts = np.concatenate([it.x for it in self.items if isinstance(it, TetBody)])
        return float(np.linalg.norm(pts.max(0) - pts.min(0)))

    def tet_bodies(self):
        return [i for i, it in enumerate(self.items) if isinstance(it, TetBody)]

    def instantiate(self, backend, finalize=True, partition=None):
        """Create the scene on a backend (GPU Simulation or oracle World).  partition = (rank, world):
        this backend runs its share of the scene decomposed over `world` GPUs."""
        backend.set_collision_compliance(self.collision_compliance)
        if self.broadphase and hasattr(backend, "set_broadphase"):
            backend.set_broadphase(self.broadphase)
        if partition is not None:
            backend.set_partition(*partition)
        ids = []
        for it in self.items:
            if isinstance(it, TetBody):
                ids.append(backend.add_tet_body(it.x0, it.tets, it.mass, it.young, it.poisson, it.alpha, it.beta))
            elif it.kind == "plane":
                ids.append(backend.add_sdf_plane(it.a, it.b, it.volume))
            elif it.kind == "sphere":
                ids.append(backend.add_sdf_sphere(it.a, it.b[0], it.volume))
            elif it.kind == "mesh":
                ids.append(backend.add_sdf_mesh(it.a, it.b, it.volume, it.res))
            else:
                ids.append(backend.add_sdf_box(it.a, it.b, it.volume))
        for i, it in zip(ids, self.items):
            if not getattr(it, "collideable", True):   # not handed to the cd system (main.cpp:77-86)
                backend.set_body_collideable(i, False)
        for (b1, b2, pairs, alpha, beta) in self.distance:
            backend.add_distance_constraints(b1, b2, pairs, alpha, beta)
        if finalize and hasattr(backend, "finalize"):
            backend.finalize()
        for i, it in zip(ids, self.items):
            if isinstance(it, TetBody):
                if it.v is None:
                    backend.upload(i, it.x)
                else:
                    backend.upload(i, it.x, it.v)
        return ids


_BIG = (-1e4, -1e4, -1e4, 1e4, 1e4, 1e4)


def octahedron(centre, radii):
    """Closed, outward-oriented triangle mesh (6 vertices, 8 faces); float32-representable positions."""
    c = np.asarray(centre, np.float64)
    r = np.asarray(radii, np.float64)
    x = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float64) * r + c
    f = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]], np.uint32)
    return x.astype(np.float32).astype(np.float64), f


def box_mesh(bmin, bmax):
    """Closed, outward-oriented triangle mesh of an axis-aligned box (8 vertices, 12 faces)."""
    lo, hi = np.asarray(bmin, np.float64), np.asarray(bmax, np.float64)
    u = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.float64)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [1, 2, 6], [1, 6, 5], [2, 3, 7],
                  [2, 7, 6], [3, 0, 4], [3, 4, 7]], np.uint32)
    return (lo + u * (hi - lo)).astype(np.float32).astype(np.float64), f


def config1_on_mesh(W=4, H=4, D=6, seed=7, res=(8, 8, 8)):
    """Config 1 with a triangle-mesh obstacle baked into a grid SDF (environment_body.cpp:12-78) poking into
    the beam from below, next to the floor plane."""
    sc = config1(W=W, H=H, D=D, seed=seed)
    x, f = octahedron((0.55 * (W - 1), 0.0, 0.5 * (D - 1)), (0.5 * (W - 1), 0.9, 0.4 * (D - 1)))
    dom = (-2.0, -3.0, -2.0, 1.1 * (W - 1) + 2.0, 0.95 * (H - 1) + 3.0, float(D - 1) + 2.0)
    sc.items.append(Sdf("mesh", x, f, dom, res))
    sc.name = "config1_on_mesh_%dx%dx%d" % (W, H, D)
    return sc


def prestrained_bar(W, H, D, seed, translate=(0.0, 0.0, 0.0), jitter=0.02, rotation=None, mass=None,
                    prestrain=_PRESTRAIN):
    """SURVEY §8(d) common inputs: x = A x0 + t + U(-jitter, jitter)^3, x0 unjittered, v = 0."""
    pos, tets = bar_model(W, H, D)
    x0 = pos.astype(np.float64)
    x = x0 * np.asarray(prestrain, dtype=np.float64)[None, :]
    if rotation is not None:
        x = x @ np.asarray(rotation, dtype=np.float64).T
    x = x + np.asarray(translate, dtype=np.float64)[None, :]
    if jitter:
        u = _uniform01(seed, x.size).reshape(x.shape)
        x = x + (2.0 * u - 1.0) * jitter
    return TetBody(x0=x0, tets=tets.astype(np.uint32), x=x, mass=mass)


def config1(W=8, H=8, D=16, seed=1, bottom=0.0):
    """Beam on a floor plane; reference detection semantics (once per frame)."""
    body = prestrained_bar(W, H, D, seed, translate=(0.0, bottom, 0.0))
    floor = Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (-100.0, -5.0, -100.0, 100.0, 5.0, 100.0))
    return Scene("config1_beam_%dx%dx%d" % (W, H, D), [body, floor])


def config2(W=21, H=21, D=51, seed=2):
    """Cantilever: the k = 0 plane is pinned (mass 0) AT ITS REST POSITIONS; free vertices are
    jittered; gravity bends the beam; no obstacle in reach.

    No pre-strain here: a pinned face holding a strain that can never relax makes the
    reference's energy constraint C = |V0| psi > 0 unsatisfiable, and the reference itself then
    amplifies 1e-16 rounding differences by ~10x per iteration (measured between two builds of
    the same algorithm), so no parity statement is possible on such a scene."""
    body = prestrained_bar(W, H, D, seed, prestrain=(1.0, 1.0, 1.0))
    pinned = np.arange(W * H * D) % D == 0
    body.x[pinned] = body.x0[pinned]
    mass = np.ones(W * H * D)
    mass[pinned] = 0.0
    body.mass = mass
    floor = Sdf("plane", (0.0, 1.0, 0.0), (0.0, -1e3, 0.0), _BIG)
    return Scene("config2_cantilever_%dx%dx%d" % (W, H, D), [body, floor])


def config3(W=41, H=51, D=101, seed=3, radius=30.0, gap=0.0, prestrain=(1.0, 1.0, 1.0), vy=0.0, floor_gap=0.0):
    """Block set down on an analytic sphere whose top pokes floor_gap above a floor plane (default: tangent to it);
    BVH broadphase and detection every substep.  The unstrained block starts in touching contact (dropped from
    height `gap`, initial velocity vy) and settles under gravity, so EVERY substep of a run detects and projects
    collision constraints on the whole bottom face (measured on B200: 4 142 contacts per detection at the default
    size, sustained) — the pre-strained variants of round 1 sprang off the obstacle within a few frames and the
    timed frames held no contact at all."""
    ext = (np.array([W, H, D]) - 1) * np.asarray(prestrain)
    centre = np.array([ext[0] / 2, -radius, ext[2] / 2])
    body = prestrained_bar(W, H, D, seed, translate=(0.0, gap, 0.0), prestrain=prestrain)
    if vy:
        body.v = np.zeros_like(body.x)
        body.v[:, 1] = vy
    sphere = Sdf("sphere", tuple(centre), (radius, 0.0, 0.0), _BIG)
    floor = Sdf("plane", (0.0, 1.0, 0.0), (0.0, -floor_gap, 0.0), _BIG)
    return Scene("config3_block_%dx%dx%d" % (W, H, D), [body, sphere, floor], detect_every_substep=True,
                 broadphase=1)


def _random_rotation(seed):
    u = _uniform01(seed, 4)
    z = 2.0 * u[0] - 1.0
    phi = 2.0 * np.pi * u[1]
    axis = np.array([np.sqrt(1 - z * z) * np.cos(phi), np.sqrt(1 - z * z) * np.sin(phi), z])
    ang = 2.0 * np.pi * u[2]
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K), u[3]


def config4(n_bodies=4096, W=6, H=6, D=17, first=0, grid=64, pitch=24.0):
    """Ensemble of independent bodies over one floor plane.  Bodies [first, first+n_bodies) of
    the 4096-body layout, so a rank can build only its shard."""
    items = []
    for b in range(first, first + n_bodies):
        R, h = _random_rotation(5000 + b)
        body = prestrained_bar(W, H, D, 1000 + b, rotation=R)
        gx, gz = (b % grid) * pitch, (b // grid) * pitch
        lowest = body.x[:, 1].min()
        body.x = body.x + np.array([gx, -lowest - 0.2 + 0.7 * h, gz])[None, :]
        items.append(body)
    items.append(Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), _BIG))
    return Scene("config4_ensemble_%dx(%dx%dx%d)" % (n_bodies, W, H, D), items)


def config5(W=101, H=101, D=161, seed=5):
    """Large single body over a floor plane."""
    body = prestrained_bar(W, H, D, seed, translate=(0.0, 0.0, 0.0))
    floor = Sdf("plane", (0.0, 1.0, 0.0), (0.0, 0.0, 0.0), _BIG)
    return Scene("config5_large_%dx%dx%d" % (W, H, D), [body, floor])
