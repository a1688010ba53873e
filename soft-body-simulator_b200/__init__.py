"""sbs-b200: host-side binding of the B200-native XPBD hot path (libsbsb200.so).

The product is the C-ABI library (include/sbs_b200.h) plus the C++ facade under cpp/sbs/.
This module is the thin ctypes plumbing tests and bench.py use to reach the C ABI; it holds
no physics.  It fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsbsb200.so")

FP32, FP64 = 32, 64
DETECT_PER_FRAME, DETECT_PER_SUBSTEP = 0, 1
SCHED_AUTO, SCHED_GRAPH, SCHED_PERSISTENT = 0, 1, 2
BROADPHASE_NONE, BROADPHASE_BVH = 0, 1
REGIONS_PENCILS, REGIONS_COMPACT = 0, 1

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)

EXPORTS = [
    "sbsb200_create", "sbsb200_destroy", "sbsb200_last_error", "sbsb200_set_stream", "sbsb200_set_schedule",
    "sbsb200_set_collision_compliance", "sbsb200_add_tet_body", "sbsb200_add_distance_constraints",
    "sbsb200_add_sdf_plane", "sbsb200_add_sdf_sphere", "sbsb200_add_sdf_box", "sbsb200_finalize",
    "sbsb200_grid_node_count", "sbsb200_grid_node_position", "sbsb200_add_sdf_grid", "sbsb200_add_sdf_mesh", "sbsb200_mesh_sdf_domain", "sbsb200_set_body_collideable",
    "sbsb200_get_sdf_grid", "sbsb200_eval_sdf",
    "sbsb200_constraint_count", "sbsb200_get_constraint_order", "sbsb200_get_surface_map", "sbsb200_get_stats",
    "sbsb200_schedule_note",
    "sbsb200_upload", "sbsb200_set_vertices", "sbsb200_download", "sbsb200_set_mass", "sbsb200_step", "sbsb200_step_host",
    "sbsb200_synchronize", "sbsb200_get_contacts", "sbsb200_debug_read_trace",
    "sbsb200_set_partition", "sbsb200_get_mailbox_handle", "sbsb200_connect_peers", "sbsb200_connect_peer_context",
    "sbsb200_get_vertex_ranks", "sbsb200_set_broadphase", "sbsb200_get_surface_triangles", "sbsb200_download_surface",
    "sbsb200_set_region_shape", "sbsb200_set_masses", "sbsb200_step_host_f32", "sbsb200_debug_trace_steps",
    "sbsb200_step_host_vertices_f32", "sbsb200_count_non_finite", "sbsb200_remove_constraints",
    "sbsb200_download_surface_rgb",
]


class Stats(C.Structure):
    _fields_ = [("n_bodies", C.c_int32), ("n_sdfs", C.c_int32), ("n_vertices", C.c_int64), ("n_tets", C.c_int64),
                ("n_distance", C.c_int64), ("n_surface_vertices", C.c_int64), ("n_green_colours", C.c_int32),
                ("n_distance_colours", C.c_int32), ("schedule", C.c_int32), ("n_regions", C.c_int32),
                ("n_interface_vertices", C.c_int64), ("kernels_launched", C.c_int64), ("frames", C.c_int64),
                ("last_contact_count", C.c_int64), ("last_step_ms", C.c_double), ("kernel_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("n_shared_vertices", C.c_int64), ("pulls_per_sweep", C.c_int64),
                ("pushes_per_sweep", C.c_int64), ("quiet_colours", C.c_int32), ("reserved_", C.c_int32),
                ("green_general_calls", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SbsError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libsbsb200.so and declare its prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SbsError("libsbsb200.so is not built (%s); run __graft_entry__.build(). "
                       "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.sbsb200_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.sbsb200_destroy.argtypes = [vp]
    L.sbsb200_destroy.restype = None
    L.sbsb200_last_error.argtypes = [vp]
    L.sbsb200_last_error.restype = C.c_char_p
    L.sbsb200_set_stream.argtypes = [vp, vp]
    L.sbsb200_set_schedule.argtypes = [vp, C.c_int]
    L.sbsb200_set_collision_compliance.argtypes = [vp, C.c_double]
    L.sbsb200_add_tet_body.argtypes = [vp, C.c_int64, _dp, _dp, C.c_int64, _u32p, C.c_double, C.c_double,
                                       C.c_double, C.c_double]
    L.sbsb200_add_distance_constraints.argtypes = [vp, C.c_int, C.c_int, C.c_int64, _u32p, C.c_double, C.c_double]
    L.sbsb200_add_sdf_plane.argtypes = [vp, _dp, _dp, _dp]
    L.sbsb200_add_sdf_sphere.argtypes = [vp, _dp, C.c_double, _dp]
    L.sbsb200_add_sdf_box.argtypes = [vp, _dp, _dp, _dp]
    L.sbsb200_grid_node_count.argtypes = [_u32p]
    L.sbsb200_grid_node_count.restype = C.c_int64
    L.sbsb200_grid_node_position.argtypes = [_dp, _dp, _u32p, C.c_int64, _dp]
    L.sbsb200_add_sdf_grid.argtypes = [vp, _dp, _dp, _u32p, _dp, C.c_int64, _dp]
    L.sbsb200_add_sdf_mesh.argtypes = [vp, C.c_int64, _dp, C.c_int64, _u32p, _dp, _u32p]
    L.sbsb200_set_body_collideable.argtypes = [vp, C.c_int, C.c_int]
    L.sbsb200_mesh_sdf_domain.argtypes = [C.c_int64, _dp, _dp, _dp]
    L.sbsb200_get_sdf_grid.argtypes = [vp, C.c_int, _dp, _u32p, _dp, C.c_int64]
    L.sbsb200_get_sdf_grid.restype = C.c_int64
    L.sbsb200_eval_sdf.argtypes = [vp, C.c_int, C.c_int64, _dp, _dp, _dp]
    L.sbsb200_finalize.argtypes = [vp]
    L.sbsb200_constraint_count.argtypes = [vp]
    L.sbsb200_constraint_count.restype = C.c_int64
    L.sbsb200_get_constraint_order.argtypes = [vp, _u32p, C.c_int64]
    L.sbsb200_get_surface_map.argtypes = [vp, C.c_int, _u32p, C.c_int64]
    L.sbsb200_get_surface_map.restype = C.c_int64
    L.sbsb200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.sbsb200_remove_constraints.argtypes = [vp, C.c_int64, _u32p]
    L.sbsb200_schedule_note.argtypes = [vp]
    L.sbsb200_schedule_note.restype = C.c_char_p
    L.sbsb200_upload.argtypes = [vp, C.c_int, _dp, _dp]
    L.sbsb200_set_vertices.argtypes = [vp, C.c_int, C.c_int64, _u32p, _dp, _dp]
    L.sbsb200_download.argtypes = [vp, C.c_int, _dp, _dp]
    L.sbsb200_set_mass.argtypes = [vp, C.c_int, C.c_int64, C.c_double]
    L.sbsb200_step.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_int]
    L.sbsb200_step_host.argtypes = [vp, C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp]
    L.sbsb200_synchronize.argtypes = [vp]
    L.sbsb200_get_contacts.argtypes = [vp, C.c_int64, _i32p, _u32p, _i32p, _dp, _dp]
    L.sbsb200_get_contacts.restype = C.c_int64
    L.sbsb200_set_broadphase.argtypes = [vp, C.c_int]
    L.sbsb200_get_surface_triangles.argtypes = [vp, C.c_int, _u32p, C.c_int64]
    L.sbsb200_get_surface_triangles.restype = C.c_int64
    L.sbsb200_download_surface.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.sbsb200_download_surface_rgb.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.c_int64, C.POINTER(C.c_float)]
    L.sbsb200_set_partition.argtypes = [vp, C.c_int, C.c_int]
    L.sbsb200_get_mailbox_handle.argtypes = [vp, C.c_char_p]
    L.sbsb200_connect_peers.argtypes = [vp, C.c_char_p, C.c_int]
    L.sbsb200_connect_peer_context.argtypes = [vp, C.c_int, vp]
    L.sbsb200_get_vertex_ranks.argtypes = [vp, C.c_int, _i32p, C.c_int64]
    L.sbsb200_set_region_shape.argtypes = [vp, C.c_int]
    L.sbsb200_set_masses.argtypes = [vp, C.c_int, C.c_int64, _u32p, _dp]
    _fp = C.POINTER(C.c_float)
    L.sbsb200_step_host_f32.argtypes = [vp, C.c_int, _fp, _fp, C.c_double, C.c_int, C.c_int, C.c_int, _fp, _fp]
    L.sbsb200_step_host_vertices_f32.argtypes = [vp, C.c_int, C.c_int64, _u32p, _fp, _fp, C.c_double, C.c_int, C.c_int,
                                                 C.c_int, _fp, _fp]
    L.sbsb200_count_non_finite.argtypes = [vp]
    L.sbsb200_count_non_finite.restype = C.c_int64
    L.sbsb200_debug_trace_steps.argtypes = [vp, C.c_int]
    L.sbsb200_debug_read_trace.argtypes = [vp, C.POINTER(C.c_int64), C.c_int64]
    L.sbsb200_debug_read_trace.restype = C.c_int64
    _lib = L
    return L


def grid_node_count(res):
    """Nodes of a CubicLagrangeDiscreteGrid with res cells per axis (host arithmetic only)."""
    r = np.ascontiguousarray(res, np.uint32)
    return int(load_library().sbsb200_grid_node_count(r.ctypes.data_as(_u32p)))


def grid_node_positions(dmin, dmax, res):
    """Positions of all grid nodes in node order (host arithmetic only) — where to sample an SDF for add_sdf_grid."""
    L = load_library()
    r = np.ascontiguousarray(res, np.uint32)
    lo, hi = _f64(dmin), _f64(dmax)
    n = grid_node_count(r)
    out = np.empty((n, 3))
    x = np.empty(3)
    for l in range(n):
        if L.sbsb200_grid_node_position(_d(lo), _d(hi), r.ctypes.data_as(_u32p), l, _d(x)) != 0:
            raise SbsError("bad grid")
        out[l] = x
    return out


def mesh_sdf_domain(positions, domain):
    """Grid domain environment_body_t's mesh constructor ends up with (environment_body.cpp:52-65)."""
    x = _f64(positions).reshape(-1, 3)
    out = np.empty(6)
    if load_library().sbsb200_mesh_sdf_domain(x.shape[0], _d(x), _d(_f64(domain).reshape(6)), _d(out)) != 0:
        raise SbsError("bad arguments")
    return out


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    return a.ctypes.data_as(_dp)


class Simulation:
    """One sbsb200 context == the reference's simulation_t + timestep_t on one GPU."""

    def __init__(self, device=0, precision=FP32, stream=None, schedule=SCHED_AUTO, region_shape=None, trace_steps=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.sbsb200_create(device, precision, C.byref(h))
        if rc:
            raise SbsError("sbsb200_create: %s" % self._L.sbsb200_last_error(None).decode())
        self._h = h
        self._nv = {}
        if stream is not None:
            self._ck(self._L.sbsb200_set_stream(self._h, C.c_void_p(stream)))
        if schedule != SCHED_AUTO:
            self._ck(self._L.sbsb200_set_schedule(self._h, schedule))
        if region_shape is not None:
            self._ck(self._L.sbsb200_set_region_shape(self._h, region_shape))
        if trace_steps:
            self._ck(self._L.sbsb200_debug_trace_steps(self._h, trace_steps))

    def _ck(self, rc):
        if rc < 0:
            raise SbsError("sbsb200 error %d: %s" % (rc, self._L.sbsb200_last_error(self._h).decode()))
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self._L.sbsb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_broadphase(self, mode):
        """0 = none (every surface vertex against every SDF), 1 = BVH (bvh_model.cpp:30-100)."""
        self._ck(self._L.sbsb200_set_broadphase(self._h, mode))

    def set_collision_compliance(self, alpha):
        self._ck(self._L.sbsb200_set_collision_compliance(self._h, alpha))

    def add_tet_body(self, x0, tets, mass=None, young=1e6, poisson=0.3, alpha=1e-4, beta=0.0):
        x0 = _f64(x0).reshape(-1, 3)
        tets = np.ascontiguousarray(tets, dtype=np.uint32).reshape(-1, 4)
        m = None if mass is None else _f64(mass)
        b = self._ck(self._L.sbsb200_add_tet_body(self._h, x0.shape[0], _d(x0), None if m is None else _d(m),
                                                  tets.shape[0], tets.ctypes.data_as(_u32p), young, poisson,
                                                  alpha, beta))
        self._nv[b] = x0.shape[0]
        return b

    def add_distance_constraints(self, b1, b2, pairs, alpha=1e-4, beta=0.0):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        self._ck(self._L.sbsb200_add_distance_constraints(self._h, b1, b2, pairs.shape[0],
                                                          pairs.ctypes.data_as(_u32p), alpha, beta))

    def add_sdf_plane(self, normal, point, volume):
        return self._ck(self._L.sbsb200_add_sdf_plane(self._h, _d(_f64(normal)), _d(_f64(point)),
                                                      _d(_f64(volume).reshape(6))))

    def add_sdf_sphere(self, centre, radius, volume):
        return self._ck(self._L.sbsb200_add_sdf_sphere(self._h, _d(_f64(centre)), radius,
                                                       _d(_f64(volume).reshape(6))))

    def add_sdf_box(self, bmin, bmax, volume):
        return self._ck(self._L.sbsb200_add_sdf_box(self._h, _d(_f64(bmin)), _d(_f64(bmax)),
                                                    _d(_f64(volume).reshape(6))))

    def set_body_collideable(self, body, flag):
        """Only models handed to the cd system collide (brute_force_cd_system_t(objects), main.cpp:77-86)."""
        self._ck(self._L.sbsb200_set_body_collideable(self._h, body, 1 if flag else 0))

    def add_sdf_grid(self, dmin, dmax, res, nodes, volume=None):
        """environment_body_t with a discrete-grid sdf_model_t (sdf_model.cpp:18)."""
        res = np.ascontiguousarray(res, np.uint32)
        nodes = _f64(nodes)
        vol = _d(_f64(volume).reshape(6)) if volume is not None else None
        return self._ck(self._L.sbsb200_add_sdf_grid(self._h, _d(_f64(dmin)), _d(_f64(dmax)), res.ctypes.data_as(_u32p),
                                                     _d(nodes), nodes.shape[0], vol))

    def add_sdf_mesh(self, positions, triangles, domain, res=None):
        """environment_body_t(sim, id, geometry, domain, resolution) (environment_body.cpp:12-78), baked on the device."""
        x = _f64(positions).reshape(-1, 3)
        tri = np.ascontiguousarray(triangles, np.uint32).reshape(-1, 3)
        r = None if res is None else np.ascontiguousarray(res, np.uint32)
        return self._ck(self._L.sbsb200_add_sdf_mesh(self._h, x.shape[0], _d(x), tri.shape[0], tri.ctypes.data_as(_u32p),
                                                     _d(_f64(domain).reshape(6)),
                                                     None if r is None else r.ctypes.data_as(_u32p)))

    def sdf_grid(self, body):
        """(domain[6], resolution[3], node values) of a grid sdf body."""
        dom = np.empty(6)
        res = np.empty(3, np.uint32)
        n = self._ck(self._L.sbsb200_get_sdf_grid(self._h, body, _d(dom), res.ctypes.data_as(_u32p), None, 0))
        nodes = np.empty(n)
        self._ck(self._L.sbsb200_get_sdf_grid(self._h, body, _d(dom), res.ctypes.data_as(_u32p), _d(nodes), n))
        return dom, res, nodes

    def eval_sdf(self, body, points):
        """sdf_model_t::evaluate on the device: (distances, gradients)."""
        pts = _f64(points).reshape(-1, 3)
        sd = np.empty(pts.shape[0])
        g = np.empty((pts.shape[0], 3))
        self._ck(self._L.sbsb200_eval_sdf(self._h, body, pts.shape[0], _d(pts), _d(sd), _d(g)))
        return sd, g

    def finalize(self):
        self._ck(self._L.sbsb200_finalize(self._h))

    def constraint_count(self):
        return self._ck(self._L.sbsb200_constraint_count(self._h))

    def constraint_order(self):
        n = self.constraint_count()
        order = np.empty(n, np.uint32)
        self._ck(self._L.sbsb200_get_constraint_order(self._h, order.ctypes.data_as(_u32p), n))
        return order

    def surface_map(self, body):
        n = self._ck(self._L.sbsb200_get_surface_map(self._h, body, None, 0))
        m = np.empty(max(n, 1), np.uint32)
        self._ck(self._L.sbsb200_get_surface_map(self._h, body, m.ctypes.data_as(_u32p), n))
        return m[:n]

    def surface_triangles(self, body):
        n = self._ck(self._L.sbsb200_get_surface_triangles(self._h, body, None, 0))
        t = np.empty(max(n, 1), np.uint32)
        self._ck(self._L.sbsb200_get_surface_triangles(self._h, body, t.ctypes.data_as(_u32p), n))
        return t[:n].reshape(-1, 3)

    def download_surface(self, body):
        """[n_surface_vertices, 6] float32: position and unit normal of every boundary vertex."""
        n = len(self.surface_map(body))
        out = np.empty((max(n, 1), 6), np.float32)
        self._ck(self._L.sbsb200_download_surface(self._h, body, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out[:n]

    def download_surface_rgb(self, body, colours):
        """[n_surface_vertices, 9] float32: position, unit normal and colour — the reference's render vertex
        (tetrahedral_mesh_boundary.cpp:170-193).  colours: one rgb triple or one per surface vertex."""
        n = len(self.surface_map(body))
        col = np.ascontiguousarray(colours, np.float32).reshape(-1, 3)
        out = np.empty((max(n, 1), 9), np.float32)
        fp = C.POINTER(C.c_float)
        self._ck(self._L.sbsb200_download_surface_rgb(self._h, body, col.ctypes.data_as(fp), len(col), out.ctypes.data_as(fp)))
        return out[:n]

    def stats(self):
        s = Stats()
        self._ck(self._L.sbsb200_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def remove_constraints(self, ids):
        """simulation_t::remove_constraint for tet constraints, by insertion index as counted at finalize; the scene
        is not re-planned (sbsb200_remove_constraints)."""
        a = np.ascontiguousarray(ids, np.uint32)
        self._ck(self._L.sbsb200_remove_constraints(self._h, len(a), a.ctypes.data_as(_u32p)))

    def schedule_note(self):
        return self._L.sbsb200_schedule_note(self._h).decode()

    def upload(self, body, x, v=None):
        x = _f64(x)
        vv = None if v is None else _f64(v)
        self._ck(self._L.sbsb200_upload(self._h, body, _d(x), None if vv is None else _d(vv)))

    def set_vertices(self, body, vertices, x, v=None):
        """Override x (= xi = xn) and optionally v of a few vertices between frames (dragging a picked vertex)."""
        ids = np.ascontiguousarray(vertices, np.uint32).reshape(-1)
        x = _f64(x).reshape(-1, 3)
        vv = None if v is None else _f64(v).reshape(-1, 3)
        self._ck(self._L.sbsb200_set_vertices(self._h, body, ids.shape[0], ids.ctypes.data_as(_u32p), _d(x),
                                              None if vv is None else _d(vv)))

    def download(self, body):
        n = self._nv[body]
        x = np.empty((n, 3))
        v = np.empty((n, 3))
        self._ck(self._L.sbsb200_download(self._h, body, _d(x), _d(v)))
        return x, v

    def set_mass(self, body, vertex, mass):
        self._ck(self._L.sbsb200_set_mass(self._h, body, vertex, mass))

    def set_masses(self, body, vertices, masses):
        ids = np.ascontiguousarray(vertices, np.uint32).reshape(-1)
        m = _f64(masses).reshape(-1)
        assert ids.shape == m.shape
        self._ck(self._L.sbsb200_set_masses(self._h, body, ids.shape[0], ids.ctypes.data_as(_u32p), _d(m)))

    def step_host_f32(self, body, x_in, v_in, dt, substeps, iterations, detect_every_substep, x_out, v_out):
        """The same as step_host with float32 host arrays [nV, 3] (v_in / x_out / v_out may be None)."""
        fp = C.POINTER(C.c_float)
        f = lambda a: None if a is None else a.ctypes.data_as(fp)
        for a in (x_in, v_in, x_out, v_out):
            assert a is None or (a.dtype == np.float32 and a.flags["C_CONTIGUOUS"])
        self._ck(self._L.sbsb200_step_host_f32(self._h, body, f(x_in), f(v_in), dt, substeps, iterations,
                                               DETECT_PER_SUBSTEP if detect_every_substep else DETECT_PER_FRAME,
                                               f(x_out), f(v_out)))

    def step_host_vertices_f32(self, body, vertices, x_in, v_in, dt, substeps, iterations, detect_every_substep, x_out,
                               v_out):
        """step_host_f32 for the listed vertices only: float32 arrays [n, 3] in the order of `vertices` (uint32)."""
        fp = C.POINTER(C.c_float)
        f = lambda a: None if a is None else a.ctypes.data_as(fp)
        assert vertices.dtype == np.uint32 and vertices.flags["C_CONTIGUOUS"]
        for a in (x_in, v_in, x_out, v_out):
            assert a is None or (a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.shape[0] == vertices.shape[0])
        self._ck(self._L.sbsb200_step_host_vertices_f32(
            self._h, body, vertices.shape[0], vertices.ctypes.data_as(_u32p), f(x_in), f(v_in), dt, substeps, iterations,
            DETECT_PER_SUBSTEP if detect_every_substep else DETECT_PER_FRAME, f(x_out), f(v_out)))

    def count_non_finite(self):
        return self._ck(self._L.sbsb200_count_non_finite(self._h))

    def contact_count(self):
        return self._ck(self._L.sbsb200_get_contacts(self._h, 0, None, None, None, None, None))

    def step(self, dt, substeps, iterations, detect_every_substep=False):
        self._ck(self._L.sbsb200_step(self._h, dt, substeps, iterations,
                                      DETECT_PER_SUBSTEP if detect_every_substep else DETECT_PER_FRAME))

    def step_host(self, body, x_in, v_in, dt, substeps, iterations, detect_every_substep, x_out, v_out):
        """x_in/v_in/x_out/v_out: C-contiguous float64 host arrays [nV, 3] (v_in may be None)."""
        self._ck(self._L.sbsb200_step_host(self._h, body, _d(x_in), None if v_in is None else _d(v_in), dt,
                                           substeps, iterations,
                                           DETECT_PER_SUBSTEP if detect_every_substep else DETECT_PER_FRAME,
                                           None if x_out is None else _d(x_out), None if v_out is None else _d(v_out)))

    # ---- one scene decomposed over several GPUs ----------------------------------------------
    def set_partition(self, rank, world):
        self._ck(self._L.sbsb200_set_partition(self._h, rank, world))

    def mailbox_handle(self):
        buf = C.create_string_buffer(64)
        self._ck(self._L.sbsb200_get_mailbox_handle(self._h, buf))
        return buf.raw

    def connect_peers(self, handles):
        """handles: list of `world` 64-byte CUDA IPC handles (own entry ignored)."""
        blob = b"".join(handles)
        self._ck(self._L.sbsb200_connect_peers(self._h, blob, len(handles)))

    def connect_peer_context(self, peer_rank, peer):
        self._ck(self._L.sbsb200_connect_peer_context(self._h, peer_rank, peer._h))

    def vertex_ranks(self, body):
        out = np.empty(self._nv[body], np.int32)
        self._ck(self._L.sbsb200_get_vertex_ranks(self._h, body, out.ctypes.data_as(_i32p), len(out)))
        return out

    def debug_trace(self):
        """[regions, steps, 8] clock stamps (see sbsb200_debug_read_trace); empty when tracing is off."""
        n = self._L.sbsb200_debug_read_trace(self._h, None, 0)
        if n <= 0:
            return np.zeros((0, 0, 8), np.int64)
        buf = np.zeros(n, np.int64)
        self._L.sbsb200_debug_read_trace(self._h, buf.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return buf

    def synchronize(self):
        self._ck(self._L.sbsb200_synchronize(self._h))

    def contacts(self):
        n = self._ck(self._L.sbsb200_get_contacts(self._h, 0, None, None, None, None, None))
        body = np.empty(max(n, 1), np.int32)
        vert = np.empty(max(n, 1), np.uint32)
        sdf = np.empty(max(n, 1), np.int32)
        pt = np.empty((max(n, 1), 3))
        nr = np.empty((max(n, 1), 3))
        self._ck(self._L.sbsb200_get_contacts(self._h, n, body.ctypes.data_as(_i32p), vert.ctypes.data_as(_u32p),
                                              sdf.ctypes.data_as(_i32p), _d(pt), _d(nr)))
        return body[:n], vert[:n], sdf[:n], pt[:n], nr[:n]
