"""ctypes binding of the CPU oracle (oracle/xpbd_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_fp = C.POINTER(C.c_float)


def build(force=False):
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "xpbd_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(os.path.join(_HERE, "xpbd_oracle.h"))):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-std=c11", "-fPIC", "-shared",
                           "-ffp-contract=off", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_collision_compliance.argtypes = [C.c_void_p, C.c_double]
        L.orc_add_tet_body.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _u32p, C.c_double,
                                       C.c_double, C.c_double, C.c_double]
        L.orc_add_distance_constraints.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _u32p,
                                                   C.c_double, C.c_double]
        L.orc_add_sdf_plane.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.orc_add_sdf_sphere.argtypes = [C.c_void_p, _dp, C.c_double, _dp]
        L.orc_add_sdf_box.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.orc_set_body_collideable.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_constraint_count.argtypes = [C.c_void_p]
        L.orc_set_constraint_order.argtypes = [C.c_void_p, _u32p, C.c_int]
        L.orc_remove_constraint.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_upload.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_download.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_set_mass.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_step.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_get_contacts.argtypes = [C.c_void_p, C.c_int, _i32p, _u32p, _i32p, _dp, _dp]
        L.orc_get_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.orc_bar_model.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _i32p]
        L.orc_svd3.argtypes = [_dp, _dp, _dp, _dp]
        L.orc_green_rest_state.argtypes = [_dp, _dp, _dp]
        L.orc_green_project.argtypes = [_dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_double, _dp, _dp]
        L.orc_green_project.restype = C.c_int
        L.orc_boundary_surface.argtypes = [C.c_int, C.c_int, _u32p, _u32p, _u32p, _i32p]
        L.orc_grid_node_count.argtypes = [_u32p]
        L.orc_grid_node_count.restype = C.c_int64
        L.orc_grid_node_position.argtypes = [_dp, _dp, _u32p, C.c_int64, _dp]
        L.orc_grid_shape.argtypes = [_dp, _dp, _dp]
        L.orc_grid_interpolate.argtypes = [_dp, _dp, _u32p, _dp, _dp, _dp]
        L.orc_grid_interpolate.restype = C.c_double
        L.orc_mesh_signed_distance.argtypes = [C.c_int, _dp, C.c_int, _u32p, C.c_int, _dp, _dp]
        L.orc_mesh_sdf_domain.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_bake_mesh_sdf.argtypes = [C.c_int, _dp, C.c_int, _u32p, _dp, _u32p, _dp, _dp, C.c_int64]
        L.orc_bake_mesh_sdf.restype = C.c_int64
        L.orc_add_sdf_grid.argtypes = [C.c_void_p, _dp, _dp, _u32p, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def bar_model(W, H, D):
    """get_simple_bar_model -> (positions float32 [V,3], tets int32 [T,4])."""
    pos = np.empty((W * H * D, 3), np.float32)
    idx = np.empty((5 * (W - 1) * (H - 1) * (D - 1), 4), np.int32)
    lib().orc_bar_model(W, H, D, pos.ctypes.data_as(_fp), idx.ctypes.data_as(_i32p))
    return pos, idx


def svd3(F):
    F = _f64(F).reshape(9)
    U = np.empty(9)
    s = np.empty(3)
    V = np.empty(9)
    lib().orc_svd3(_d(F), _d(U), _d(s), _d(V))
    return U.reshape(3, 3), s, V.reshape(3, 3)


def green_rest_state(x0):
    x0 = _f64(x0).reshape(12)
    DmInv = np.empty(9)
    V0 = np.empty(1)
    lib().orc_green_rest_state(_d(x0), _d(DmInv), _d(V0))
    return DmInv.reshape(3, 3), float(V0[0])


def green_project(xi, xn, w, DmInv, V0, young, poisson, alpha, beta, dt, lagrange=0.0):
    xi = _f64(xi).reshape(12).copy()
    xn = _f64(xn).reshape(12)
    w = _f64(w).reshape(4)
    DmInv = _f64(DmInv).reshape(9)
    lam = np.array([lagrange], np.float64)
    diag = np.zeros(8)
    ran = lib().orc_green_project(_d(xi), _d(xn), _d(w), _d(DmInv), V0, young, poisson, alpha, beta,
                                  dt, _d(lam), _d(diag))
    return xi.reshape(4, 3), float(lam[0]), bool(ran), diag


def boundary_surface(nV, tets):
    tets = _u32(tets).reshape(-1, 4)
    nT = tets.shape[0]
    s2t = np.empty(max(nV, 1), np.uint32)
    tris = np.empty((max(4 * nT, 1), 3), np.uint32)
    ntri = C.c_int32(0)
    nvs = lib().orc_boundary_surface(nV, nT, tets.ctypes.data_as(_u32p), s2t.ctypes.data_as(_u32p),
                                     tris.ctypes.data_as(_u32p), C.byref(ntri))
    return s2t[:nvs].copy(), tris[:ntri.value].copy()


def grid_node_count(res):
    return int(lib().orc_grid_node_count(_u32(res).ctypes.data_as(_u32p)))


def grid_node_positions(dmin, dmax, res):
    """Positions of all nodes of a CubicLagrangeDiscreteGrid, in node order."""
    dmin, dmax, res = _f64(dmin), _f64(dmax), _u32(res)
    n = grid_node_count(res)
    out = np.empty((n, 3))
    x = np.empty(3)
    f = lib().orc_grid_node_position
    for l in range(n):
        f(_d(dmin), _d(dmax), res.ctypes.data_as(_u32p), l, _d(x))
        out[l] = x
    return out


def grid_shape(xi):
    N = np.empty(32)
    dN = np.empty((32, 3))
    lib().orc_grid_shape(_d(_f64(xi)), _d(N), _d(dN))
    return N, dN


def grid_interpolate(dmin, dmax, res, nodes, points):
    """CubicLagrangeDiscreteGrid::interpolate at every point: (values, gradients)."""
    dmin, dmax, res, nodes = _f64(dmin), _f64(dmax), _u32(res), _f64(nodes)
    pts = _f64(points).reshape(-1, 3)
    val = np.empty(len(pts))
    grad = np.empty((len(pts), 3))
    g = np.empty(3)
    f = lib().orc_grid_interpolate
    for i, p in enumerate(pts):
        val[i] = f(_d(dmin), _d(dmax), res.ctypes.data_as(_u32p), _d(nodes), _d(np.ascontiguousarray(p)), _d(g))
        grad[i] = g
    return val, grad


def mesh_signed_distance(x, faces, points):
    x, faces, pts = _f64(x).reshape(-1, 3), _u32(faces).reshape(-1, 3), _f64(points).reshape(-1, 3)
    out = np.empty(len(pts))
    lib().orc_mesh_signed_distance(len(x), _d(x), len(faces), faces.ctypes.data_as(_u32p), len(pts), _d(pts), _d(out))
    return out


def mesh_sdf_domain(x, domain):
    x = _f64(x).reshape(-1, 3)
    out = np.empty(6)
    lib().orc_mesh_sdf_domain(len(x), _d(x), _d(_f64(domain).reshape(6)), _d(out))
    return out


def bake_mesh_sdf(x, faces, domain, res):
    """environment_body_t(sim, id, geometry, domain, resolution): (extended domain[6], node values)."""
    x, faces, res = _f64(x).reshape(-1, 3), _u32(faces).reshape(-1, 3), _u32(res)
    n = grid_node_count(res)
    nodes = np.empty(n)
    dom = np.empty(6)
    lib().orc_bake_mesh_sdf(len(x), _d(x), len(faces), faces.ctypes.data_as(_u32p), _d(_f64(domain).reshape(6)),
                            res.ctypes.data_as(_u32p), _d(dom), _d(nodes), n)
    return dom, nodes


class World:
    """The reference's simulation_t + timestep_t restated (see xpbd_oracle.h)."""

    def __init__(self):
        self._h = C.c_void_p(lib().orc_create())
        self._nv = {}

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_collision_compliance(self, a):
        lib().orc_set_collision_compliance(self._h, a)

    def add_tet_body(self, x0, tets, mass=None, young=1e6, poisson=0.3, alpha=1e-4, beta=0.0):
        x0 = _f64(x0).reshape(-1, 3)
        tets = _u32(tets).reshape(-1, 4)
        m = None if mass is None else _f64(mass)
        b = lib().orc_add_tet_body(self._h, x0.shape[0], _d(x0), None if m is None else _d(m),
                                   tets.shape[0], tets.ctypes.data_as(_u32p), young, poisson, alpha, beta)
        self._nv[b] = x0.shape[0]
        return b

    def add_distance_constraints(self, b1, b2, pairs, alpha=1e-4, beta=0.0):
        pairs = _u32(pairs).reshape(-1, 2)
        rc = lib().orc_add_distance_constraints(self._h, b1, b2, pairs.shape[0],
                                                pairs.ctypes.data_as(_u32p), alpha, beta)
        if rc:
            raise RuntimeError("orc_add_distance_constraints failed")

    def add_sdf_plane(self, normal, point, volume):
        return lib().orc_add_sdf_plane(self._h, _d(_f64(normal)), _d(_f64(point)), _d(_f64(volume).reshape(6)))

    def add_sdf_sphere(self, centre, radius, volume):
        return lib().orc_add_sdf_sphere(self._h, _d(_f64(centre)), radius, _d(_f64(volume).reshape(6)))

    def add_sdf_box(self, bmin, bmax, volume):
        return lib().orc_add_sdf_box(self._h, _d(_f64(bmin)), _d(_f64(bmax)), _d(_f64(volume).reshape(6)))

    def add_sdf_grid(self, dmin, dmax, res, nodes, volume=None):
        vol = _d(_f64(volume).reshape(6)) if volume is not None else None
        return lib().orc_add_sdf_grid(self._h, _d(_f64(dmin)), _d(_f64(dmax)), _u32(res).ctypes.data_as(_u32p),
                                      _d(_f64(nodes)), vol)

    def add_sdf_mesh(self, x, faces, domain, res=None):
        """environment_body_t(sim, id, geometry, domain, resolution) (environment_body.cpp:12-78)."""
        res = (10, 10, 10) if res is None else res
        dom, nodes = bake_mesh_sdf(x, faces, domain, res)
        return self.add_sdf_grid(dom[:3], dom[3:], res, nodes, dom)

    def set_body_collideable(self, body, flag):
        if lib().orc_set_body_collideable(self._h, body, 1 if flag else 0):
            raise RuntimeError("bad body")

    def constraint_count(self):
        return lib().orc_constraint_count(self._h)

    def remove_constraint(self, index):
        """simulation_t::remove_constraint(index) (simulation.cpp:34-39): swapped with the last one and dropped."""
        if lib().orc_remove_constraint(self._h, int(index)):
            raise RuntimeError("orc_remove_constraint failed")

    def set_constraint_order(self, order):
        order = _u32(order)
        rc = lib().orc_set_constraint_order(self._h, order.ctypes.data_as(_u32p), order.shape[0])
        if rc:
            raise RuntimeError("orc_set_constraint_order failed: %d" % rc)

    def upload(self, body, x, v=None):
        x = _f64(x)
        vv = None if v is None else _f64(v)
        rc = lib().orc_upload(self._h, body, _d(x), None if vv is None else _d(vv))
        if rc:
            raise RuntimeError("orc_upload failed")

    def download(self, body):
        n = self._nv[body]
        x = np.empty((n, 3))
        v = np.empty((n, 3))
        rc = lib().orc_download(self._h, body, _d(x), _d(v))
        if rc:
            raise RuntimeError("orc_download failed")
        return x, v

    def set_mass(self, body, vertex, mass):
        if lib().orc_set_mass(self._h, body, vertex, mass):
            raise RuntimeError("orc_set_mass failed")

    def step(self, dt, substeps, iterations, detect_every_substep=False):
        rc = lib().orc_step(self._h, dt, substeps, iterations, 1 if detect_every_substep else 0)
        if rc:
            raise RuntimeError("orc_step failed")

    def contacts(self):
        n = lib().orc_get_contacts(self._h, 0, None, None, None, None, None)
        body = np.empty(max(n, 1), np.int32)
        vert = np.empty(max(n, 1), np.uint32)
        sdf = np.empty(max(n, 1), np.int32)
        pt = np.empty((max(n, 1), 3))
        nr = np.empty((max(n, 1), 3))
        lib().orc_get_contacts(self._h, n, body.ctypes.data_as(_i32p), vert.ctypes.data_as(_u32p),
                               sdf.ctypes.data_as(_i32p), _d(pt), _d(nr))
        return body[:n], vert[:n], sdf[:n], pt[:n], nr[:n]

    def counters(self):
        a = C.c_uint64(0)
        b = C.c_uint64(0)
        lib().orc_get_counters(self._h, C.byref(a), C.byref(b))
        return a.value, b.value
