"""ctypes binding of oracle/_ref/libsbsref.so: the reference's OWN physics sources
(/root/reference/src/physics/**, unmodified) compiled against oracle/ref_shim by
oracle/build_ref.sh.  TEST INFRASTRUCTURE ONLY — same World interface as oracle.oracle.World.

The library is built in the authoring container (where /root/reference is mounted) and travels
to the GPU box as a prebuilt, git-ignored file; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsbsref.so")

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)


def available():
    return os.path.exists(LIB_PATH)


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("ref_driver.cpp", "build_ref.sh", "ref_shim/shim_eigen.h",
                                             "ref_shim/shim_discregrid.h")]
    if not force and available() and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    subprocess.check_call(["sh", os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ref_create.restype = vp
        L.ref_destroy.argtypes = [vp]
        L.ref_set_collision_compliance.argtypes = [vp, C.c_double]
        L.ref_add_tet_body.argtypes = [vp, C.c_int, _dp, _dp, C.c_int, _u32p, C.c_double, C.c_double, C.c_double,
                                       C.c_double]
        L.ref_add_distance_constraints.argtypes = [vp, C.c_int, C.c_int, C.c_int, _u32p, C.c_double, C.c_double]
        L.ref_add_sdf_plane.argtypes = [vp, _dp, _dp, _dp]
        L.ref_add_sdf_sphere.argtypes = [vp, _dp, C.c_double, _dp]
        L.ref_add_sdf_box.argtypes = [vp, _dp, _dp, _dp]
        L.ref_add_sdf_mesh.argtypes = [vp, C.c_int, _dp, C.c_int, _u32p, _dp, _u32p]
        L.ref_add_sdf_grid.argtypes = [vp, _dp, _dp, _u32p, _dp, _dp]
        L.ref_sdf_evaluate.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, _dp]
        L.ref_get_volume.argtypes = [vp, C.c_int, _dp]
        L.ref_set_body_collideable.argtypes = [vp, C.c_int, C.c_int]
        L.ref_constraint_count.argtypes = [vp]
        L.ref_set_constraint_order.argtypes = [vp, _u32p, C.c_int]
        L.ref_remove_constraint.argtypes = [vp, C.c_uint32]
        L.ref_upload.argtypes = [vp, C.c_int, _dp, _dp]
        L.ref_download.argtypes = [vp, C.c_int, _dp, _dp]
        L.ref_set_mass.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        L.ref_step.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_int]
        L.ref_get_contacts.argtypes = [vp, C.c_int, _i32p, _u32p, _i32p, _dp, _dp]
        L.ref_get_surface_map.argtypes = [vp, C.c_int, _u32p, C.c_int]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    return a.ctypes.data_as(_dp)


class World:
    def __init__(self):
        self._h = C.c_void_p(lib().ref_create())
        self._nv = {}

    def close(self):
        if self._h:
            lib().ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_collision_compliance(self, a):
        lib().ref_set_collision_compliance(self._h, a)

    def add_tet_body(self, x0, tets, mass=None, young=1e6, poisson=0.3, alpha=1e-4, beta=0.0):
        x0 = _f64(x0).reshape(-1, 3)
        tets = np.ascontiguousarray(tets, dtype=np.uint32).reshape(-1, 4)
        m = None if mass is None else _f64(mass)
        b = lib().ref_add_tet_body(self._h, x0.shape[0], _d(x0), None if m is None else _d(m), tets.shape[0],
                                   tets.ctypes.data_as(_u32p), young, poisson, alpha, beta)
        self._nv[b] = x0.shape[0]
        return b

    def add_distance_constraints(self, b1, b2, pairs, alpha=1e-4, beta=0.0):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        lib().ref_add_distance_constraints(self._h, b1, b2, pairs.shape[0], pairs.ctypes.data_as(_u32p), alpha, beta)

    def add_sdf_plane(self, normal, point, volume):
        return lib().ref_add_sdf_plane(self._h, _d(_f64(normal)), _d(_f64(point)), _d(_f64(volume).reshape(6)))

    def add_sdf_sphere(self, centre, radius, volume):
        return lib().ref_add_sdf_sphere(self._h, _d(_f64(centre)), radius, _d(_f64(volume).reshape(6)))

    def add_sdf_box(self, bmin, bmax, volume):
        return lib().ref_add_sdf_box(self._h, _d(_f64(bmin)), _d(_f64(bmax)), _d(_f64(volume).reshape(6)))

    def add_sdf_mesh(self, x, faces, domain, res=None):
        """environment_body_t(sim, id, geometry, domain, resolution): the reference's own bake."""
        res = (10, 10, 10) if res is None else res
        x = _f64(x).reshape(-1, 3)
        faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        res = np.ascontiguousarray(res, dtype=np.uint32)
        return lib().ref_add_sdf_mesh(self._h, len(x), _d(x), len(faces), faces.ctypes.data_as(_u32p),
                                      _d(_f64(domain).reshape(6)), res.ctypes.data_as(_u32p))

    def add_sdf_grid(self, dmin, dmax, res, nodes, volume=None):
        res = np.ascontiguousarray(res, dtype=np.uint32)
        vol = _d(_f64(volume).reshape(6)) if volume is not None else None
        return lib().ref_add_sdf_grid(self._h, _d(_f64(dmin)), _d(_f64(dmax)), res.ctypes.data_as(_u32p),
                                      _d(_f64(nodes)), vol)

    def sdf_evaluate(self, body, points):
        pts = _f64(points).reshape(-1, 3)
        sd = np.empty(len(pts))
        g = np.empty((len(pts), 3))
        if lib().ref_sdf_evaluate(self._h, body, len(pts), _d(pts), _d(sd), _d(g)):
            raise RuntimeError("not an environment body")
        return sd, g

    def volume(self, body):
        out = np.empty(6)
        lib().ref_get_volume(self._h, body, _d(out))
        return out

    def set_body_collideable(self, body, flag):
        if lib().ref_set_body_collideable(self._h, body, 1 if flag else 0):
            raise RuntimeError("bad body")

    def constraint_count(self):
        return lib().ref_constraint_count(self._h)

    def remove_constraint(self, index):
        """simulation_t::remove_constraint(index): the reference swaps it with the last one and drops it."""
        if lib().ref_remove_constraint(self._h, int(index)):
            raise RuntimeError("ref_remove_constraint failed")

    def set_constraint_order(self, order):
        order = np.ascontiguousarray(order, dtype=np.uint32)
        rc = lib().ref_set_constraint_order(self._h, order.ctypes.data_as(_u32p), order.shape[0])
        if rc:
            raise RuntimeError("ref_set_constraint_order failed: %d" % rc)

    def upload(self, body, x, v=None):
        x = _f64(x)
        vv = None if v is None else _f64(v)
        lib().ref_upload(self._h, body, _d(x), None if vv is None else _d(vv))

    def download(self, body):
        n = self._nv[body]
        x = np.empty((n, 3))
        v = np.empty((n, 3))
        lib().ref_download(self._h, body, _d(x), _d(v))
        return x, v

    def set_mass(self, body, vertex, mass):
        lib().ref_set_mass(self._h, body, vertex, mass)

    def step(self, dt, substeps, iterations, detect_every_substep=False):
        lib().ref_step(self._h, dt, substeps, iterations, 1 if detect_every_substep else 0)

    def contacts(self):
        n = lib().ref_get_contacts(self._h, 0, None, None, None, None, None)
        body = np.empty(max(n, 1), np.int32)
        vert = np.empty(max(n, 1), np.uint32)
        sdf = np.empty(max(n, 1), np.int32)
        pt = np.empty((max(n, 1), 3))
        nr = np.empty((max(n, 1), 3))
        lib().ref_get_contacts(self._h, n, body.ctypes.data_as(_i32p), vert.ctypes.data_as(_u32p),
                               sdf.ctypes.data_as(_i32p), _d(pt), _d(nr))
        return body[:n], vert[:n], sdf[:n], pt[:n], nr[:n]

    def surface_map(self, body):
        n = lib().ref_get_surface_map(self._h, body, None, 0)
        m = np.empty(max(n, 1), np.uint32)
        lib().ref_get_surface_map(self._h, body, m.ctypes.data_as(_u32p), n)
        return m[:n]
