"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding over oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The shipped engine (moveit_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)
c_lp = C.POINTER(C.c_long)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("df_oracle_capi.cpp", "df_oracle.hpp")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_field_create.restype = C.c_void_p
        L.orc_field_create.argtypes = [C.c_double] * 8 + [C.c_int]
        L.orc_field_destroy.argtypes = [C.c_void_p]
        L.orc_field_dims.argtypes = [C.c_void_p, c_ip]
        for n in ("orc_field_add_points", "orc_field_remove_points"):
            getattr(L, n).argtypes = [C.c_void_p, c_dp, C.c_long]
        L.orc_field_update_points.argtypes = [C.c_void_p, c_dp, C.c_long, c_dp, C.c_long]
        L.orc_field_reset.argtypes = [C.c_void_p]
        L.orc_field_get_d2.argtypes = [C.c_void_p, c_i32p, C.c_int]
        L.orc_field_distance_world.restype = C.c_double
        L.orc_field_distance_world.argtypes = [C.c_void_p] + [C.c_double] * 3
        L.orc_field_distance_cell.restype = C.c_double
        L.orc_field_distance_cell.argtypes = [C.c_void_p] + [C.c_int] * 3
        L.orc_field_distance_gradient.argtypes = [C.c_void_p] + [C.c_double] * 3 + [c_dp]
        L.orc_field_nearest_cell.restype = C.c_int
        L.orc_field_nearest_cell.argtypes = [C.c_void_p] + [C.c_int] * 3 + [c_dp, c_ip]
        L.orc_field_world_to_grid.restype = C.c_int
        L.orc_field_world_to_grid.argtypes = [C.c_void_p] + [C.c_double] * 3 + [c_ip]
        L.orc_field_grid_to_world.argtypes = [C.c_void_p] + [C.c_int] * 3 + [c_dp]
        L.orc_field_uninitialized_distance.restype = C.c_double
        L.orc_field_uninitialized_distance.argtypes = [C.c_void_p]
        L.orc_field_pack_occupancy.restype = C.c_long
        L.orc_field_pack_occupancy.argtypes = [C.c_void_p, c_u8p, C.c_long]
        L.orc_field_unpack_occupancy.argtypes = [C.c_void_p, c_u8p, C.c_long]
        L.orc_field_write_stream.restype = C.c_long
        L.orc_field_write_stream.argtypes = [C.c_void_p, c_u8p, C.c_long]
        L.orc_field_from_stream.restype = C.c_void_p
        L.orc_field_from_stream.argtypes = [c_u8p, C.c_long, C.c_double, C.c_int]
        L.orc_field_geometry.argtypes = [C.c_void_p, c_dp]
        L.orc_field_add_voxels.argtypes = [C.c_void_p, c_i32p, C.c_long]
        L.orc_field_make_exact.argtypes = [C.c_void_p]
        for n in ("orc_edt_exact", "orc_edt_brute"):
            getattr(L, n).argtypes = [c_u8p, C.c_int, C.c_int, C.c_int, C.c_int, c_i32p]
        L.orc_shape_points.restype = C.c_long
        L.orc_shape_points.argtypes = [C.c_int, c_dp, c_dp, C.c_double, C.c_double, C.c_double, c_dp, C.c_long]
        L.orc_octree_points.restype = C.c_long
        L.orc_octree_points.argtypes = [C.c_double, c_dp, C.c_int, c_dp, C.c_long]
        L.orc_body_decomposition.argtypes = [C.c_int, c_ip, c_dp, c_dp, C.c_double, C.c_double, c_dp, C.c_long, c_lp,
                                             c_dp, C.c_long, c_lp, c_dp]
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_dp, c_ip, c_ip, c_ip, c_dp, c_dp, C.c_int]
        L.orc_model_destroy.argtypes = [C.c_void_p]
        L.orc_model_set_link_geometry.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_dp, C.c_long, c_dp]
        L.orc_model_fk.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_env_create.restype = C.c_void_p
        L.orc_env_create.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_env_destroy.argtypes = [C.c_void_p]
        L.orc_env_world_field.restype = C.c_void_p
        L.orc_env_world_field.argtypes = [C.c_void_p]
        L.orc_env_self_field.restype = C.c_void_p
        L.orc_env_self_field.argtypes = [C.c_void_p]
        L.orc_env_set_group.argtypes = [C.c_void_p, c_ip, C.c_int, c_ip, c_dp]
        L.orc_env_build_link_fields.argtypes = [C.c_void_p]
        L.orc_env_group_sphere_count.restype = C.c_int
        L.orc_env_group_sphere_count.argtypes = [C.c_void_p]
        L.orc_env_group_tables.argtypes = [C.c_void_p, c_u8p, c_u8p]
        L.orc_env_sphere_centers.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_env_check_batch.restype = C.c_double
        L.orc_env_check_batch.argtypes = [C.c_void_p, c_dp, C.c_long, C.c_int, c_u8p, C.c_int]
        L.orc_env_gradients_batch.restype = C.c_double
        L.orc_env_gradients_batch.argtypes = [C.c_void_p, c_dp, C.c_long, C.c_int, c_dp, c_dp, c_i32p, c_dp, c_dp]
        L.orc_env_check_contacts.restype = C.c_long
        L.orc_env_check_contacts.argtypes = [C.c_void_p, c_dp, C.c_long, C.c_long, c_dp, C.c_long, c_ip]
        L.orc_hardware_concurrency.restype = C.c_int
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_ip)


SHAPE_SPHERE, SHAPE_CYLINDER, SHAPE_BOX = 1, 2, 4


def pose12(xyz=(0, 0, 0), R=None):
    T = np.zeros((3, 4))
    T[:, :3] = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    T[:, 3] = xyz
    return T


class Field:
    """PropagationDistanceField restatement (corner origin, like the reference constructor)."""

    def __init__(self, size, res, origin, max_dist, signed=False, _borrowed=None):
        self._own = _borrowed is None
        if self._own:
            self.h = lib().orc_field_create(size[0], size[1], size[2], res, origin[0], origin[1], origin[2], max_dist,
                                            int(signed))
        else:
            self.h = _borrowed
        d = (C.c_int * 4)()
        lib().orc_field_dims(self.h, d)
        self.dims = (d[0], d[1], d[2])
        self.max_d2 = d[3]

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib().orc_field_destroy(self.h)
            self.h = None

    def add_points(self, pts):
        a, p = _d(np.reshape(pts, (-1, 3)))
        lib().orc_field_add_points(self.h, p, len(a))

    def remove_points(self, pts):
        a, p = _d(np.reshape(pts, (-1, 3)))
        lib().orc_field_remove_points(self.h, p, len(a))

    def update_points(self, old, new):
        a, pa = _d(np.reshape(old, (-1, 3)))
        b, pb = _d(np.reshape(new, (-1, 3)))
        lib().orc_field_update_points(self.h, pa, len(a), pb, len(b))

    def add_voxels(self, ijk):
        a = np.ascontiguousarray(np.reshape(ijk, (-1, 3)), dtype=np.int32)
        lib().orc_field_add_voxels(self.h, a.ctypes.data_as(c_i32p), len(a))

    def reset(self):
        lib().orc_field_reset(self.h)

    def make_exact(self):
        """Swap the propagated distances for the exact clamped EDT of the same zero set (second oracle, F4)."""
        lib().orc_field_make_exact(self.h)

    def d2(self, negative=False):
        out = np.empty(self.dims, dtype=np.int32)
        lib().orc_field_get_d2(self.h, out.ctypes.data_as(c_i32p), int(negative))
        return out

    def distance_world(self, x, y, z):
        return lib().orc_field_distance_world(self.h, x, y, z)

    def distance_cell(self, x, y, z):
        return lib().orc_field_distance_cell(self.h, x, y, z)

    def distance_gradient(self, x, y, z):
        out = (C.c_double * 5)()
        lib().orc_field_distance_gradient(self.h, x, y, z, out)
        return out[0], np.array([out[1], out[2], out[3]]), bool(out[4])

    def nearest_cell(self, x, y, z):
        dist = C.c_double()
        pos = (C.c_int * 3)()
        r = lib().orc_field_nearest_cell(self.h, x, y, z, C.byref(dist), pos)
        return r, dist.value, (pos[0], pos[1], pos[2])

    def world_to_grid(self, x, y, z):
        o = (C.c_int * 3)()
        v = lib().orc_field_world_to_grid(self.h, x, y, z, o)
        return bool(v), (o[0], o[1], o[2])

    def grid_to_world(self, x, y, z):
        o = (C.c_double * 3)()
        lib().orc_field_grid_to_world(self.h, x, y, z, o)
        return (o[0], o[1], o[2])

    def uninitialized_distance(self):
        return lib().orc_field_uninitialized_distance(self.h)

    def pack_occupancy(self):
        n = lib().orc_field_pack_occupancy(self.h, None, 0)
        out = np.empty(n, dtype=np.uint8)
        lib().orc_field_pack_occupancy(self.h, out.ctypes.data_as(c_u8p), n)
        return out

    def write_stream(self):
        """writeToStream (propagation_distance_field.cpp:636-673) -> bytes."""
        n = lib().orc_field_write_stream(self.h, None, 0)
        assert n > 0
        out = np.empty(n, dtype=np.uint8)
        lib().orc_field_write_stream(self.h, out.ctypes.data_as(c_u8p), n)
        return out.tobytes()

    @classmethod
    def from_stream(cls, data, max_dist, signed=False):
        """PropagationDistanceField(std::istream&, max_distance, propagate_negative); None when readFromStream fails."""
        b = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        h = lib().orc_field_from_stream(b.ctypes.data_as(c_u8p), len(b), max_dist, int(signed))
        if not h:
            return None
        f = cls.__new__(cls)
        f._own = True
        f.h = h
        d = (C.c_int * 4)()
        lib().orc_field_dims(h, d)
        f.dims = (d[0], d[1], d[2])
        f.max_d2 = d[3]
        return f

    def geometry(self):
        """(resolution, size xyz, origin xyz)"""
        g = np.empty(7)
        lib().orc_field_geometry(self.h, g.ctypes.data_as(c_dp))
        return g[0], tuple(g[1:4]), tuple(g[4:7])

    def unpack_occupancy(self, b):
        b = np.ascontiguousarray(b, dtype=np.uint8)
        lib().orc_field_unpack_occupancy(self.h, b.ctypes.data_as(c_u8p), len(b))


def edt_exact(occ, max_d2, brute=False):
    occ = np.ascontiguousarray(occ, dtype=np.uint8)
    out = np.empty(occ.shape, dtype=np.int32)
    f = lib().orc_edt_brute if brute else lib().orc_edt_exact
    f(occ.ctypes.data_as(c_u8p), occ.shape[0], occ.shape[1], occ.shape[2], int(max_d2), out.ctypes.data_as(c_i32p))
    return out


def shape_points(shape_type, dims, pose, res, padding=0.0, scale=1.0):
    dims3 = np.zeros(3)
    dims3[:len(dims)] = dims
    _, pd = _d(dims3)
    pose_a, pp = _d(pose)
    n = lib().orc_shape_points(shape_type, pd, pp, scale, padding, res, None, 0)
    out = np.empty((n, 3))
    lib().orc_shape_points(shape_type, pd, pp, scale, padding, res, out.ctypes.data_as(c_dp), n)
    return out


def octree_points(res, leaves):
    a, p = _d(np.reshape(leaves, (-1, 4)))
    n = lib().orc_octree_points(res, p, len(a), None, 0)
    out = np.empty((n, 3))
    lib().orc_octree_points(res, p, len(a), out.ctypes.data_as(c_dp), n)
    return out


def body_decomposition(shapes, res, padding):
    """shapes: list of (type, dims, pose3x4).  Returns spheres[ns,4], points[np,3], bsphere[4]."""
    types, tp = _i([s[0] for s in shapes])
    dims = np.zeros((len(shapes), 3))
    for k, s in enumerate(shapes):
        dims[k, :len(s[1])] = s[1]
    _, dp = _d(dims)
    poses, pp = _d(np.stack([np.asarray(s[2], dtype=np.float64) for s in shapes]))
    ns, npt = C.c_long(), C.c_long()
    bs = np.zeros(4)
    lib().orc_body_decomposition(len(shapes), tp, dp, pp, res, padding, None, 0, C.byref(ns), None, 0, C.byref(npt),
                                 bs.ctypes.data_as(c_dp))
    sph = np.empty((ns.value, 4))
    pts = np.empty((npt.value, 3))
    lib().orc_body_decomposition(len(shapes), tp, dp, pp, res, padding, sph.ctypes.data_as(c_dp), ns.value,
                                 C.byref(ns), pts.ctypes.data_as(c_dp), npt.value, C.byref(npt),
                                 bs.ctypes.data_as(c_dp))
    return sph, pts, bs


class Model:
    """desc: dict with parent, joint_type, axis, origin (L,3,4), origin_is_identity, var_index, mimic_var,
    mimic_mult, mimic_offset, nvar and per-link geometry lists (spheres, points, bsphere, has_geometry)."""

    def __init__(self, desc):
        L = len(desc["parent"])
        self.L = L
        self.nvar = int(desc["nvar"])
        _, p_parent = _i(desc["parent"])
        _, p_jt = _i(desc["joint_type"])
        _, p_axis = _d(desc["axis"])
        _, p_origin = _d(desc["origin"])
        _, p_oid = _i(desc["origin_is_identity"])
        _, p_var = _i(desc["var_index"])
        _, p_mv = _i(desc["mimic_var"])
        _, p_mm = _d(desc["mimic_mult"])
        _, p_mo = _d(desc["mimic_offset"])
        self.h = lib().orc_model_create(L, p_parent, p_jt, p_axis, p_origin, p_oid, p_var, p_mv, p_mm, p_mo, self.nvar)
        for i in range(L):
            sph, ps = _d(np.reshape(desc["spheres"][i], (-1, 4)))
            pts, pp = _d(np.reshape(desc["points"][i], (-1, 3)))
            _, pb = _d(desc["bsphere"][i])
            lib().orc_model_set_link_geometry(self.h, i, int(desc["has_geometry"][i]), ps, len(sph), pp, len(pts), pb)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_model_destroy(self.h)
            self.h = None

    def fk(self, q):
        _, pq = _d(q)
        out = np.empty((self.L, 3, 4))
        lib().orc_model_fk(self.h, pq, out.ctypes.data_as(c_dp))
        return out


class Env:
    """CollisionEnvDistanceField restatement.  origin is the CENTRE of the field (reference semantics)."""

    def __init__(self, model, size, origin, res=0.02, tol=0.0, max_prop=0.25, signed=False):
        self.model = model
        _, ps = _d(size)
        _, po = _d(origin)
        self.h = lib().orc_env_create(model.h, ps, po, int(signed), res, tol, max_prop)
        self.world = Field(None, None, None, None, _borrowed=lib().orc_env_world_field(self.h))
        self.n_group = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_env_destroy(self.h)
            self.h = None

    def set_group(self, group_links, acm_entry=None, self_state=None):
        g, pg = _i(group_links)
        pa = None
        if acm_entry is not None:
            a, pa = _i(np.reshape(acm_entry, (-1,)))
        pq = None
        if self_state is not None:
            q, pq = _d(self_state)
        lib().orc_env_set_group(self.h, pg, len(g), pa, pq)
        self.n_group = len(g)

    def self_field(self):
        h = lib().orc_env_self_field(self.h)
        return Field(None, None, None, None, _borrowed=h) if h else None

    def build_link_fields(self):
        lib().orc_env_build_link_fields(self.h)

    def sphere_count(self):
        return lib().orc_env_group_sphere_count(self.h)

    def group_tables(self):
        n = self.n_group
        s = np.zeros(n, dtype=np.uint8)
        m = np.zeros((n, n), dtype=np.uint8)
        lib().orc_env_group_tables(self.h, s.ctypes.data_as(c_u8p), m.ctypes.data_as(c_u8p))
        return s, m

    def sphere_centers(self, q):
        _, pq = _d(q)
        out = np.empty((self.sphere_count(), 3))
        lib().orc_env_sphere_centers(self.h, pq, out.ctypes.data_as(c_dp))
        return out

    def check_batch(self, q, mode=3, n_threads=1, return_time=False):
        q, n = _states(q, self.model.nvar)
        pq = q.ctypes.data_as(c_dp)
        out = np.empty(n, dtype=np.uint8)
        t = lib().orc_env_check_batch(self.h, pq, n, mode, out.ctypes.data_as(c_u8p), n_threads)
        return (out, t) if return_time else out

    def check_contacts(self, q, max_contacts=1, max_contacts_per_pair=1):
        """checkCollision with req.contacts == true for ONE state -> (collision, [(link1, body2, sphere, pos[3])]);
        link indices are positions in the group's link list, body2 -1 = "self", -2 = "environment"."""
        _, pq = _d(q)
        cap = max(int(max_contacts), 1) + 8
        out = np.zeros((cap, 6))
        col = C.c_int(0)
        n = lib().orc_env_check_contacts(self.h, pq, int(max_contacts), int(max_contacts_per_pair),
                                         out.ctypes.data_as(c_dp), cap, C.byref(col))
        assert n <= cap
        return bool(col.value), [(int(r[0]), int(r[1]), int(r[2]), r[3:6].copy()) for r in out[:n]]

    def gradients_batch(self, q, flags=7):
        q, N = _states(q, self.model.nvar)
        pq = q.ctypes.data_as(c_dp)
        S = self.sphere_count()
        dist = np.empty((N, S))
        grad = np.empty((N, S, 3))
        typ = np.empty((N, S), dtype=np.int32)
        cen = np.empty((N, S, 3))
        clo = np.empty((N, self.n_group))
        t = lib().orc_env_gradients_batch(self.h, pq, N, flags, dist.ctypes.data_as(c_dp), grad.ctypes.data_as(c_dp),
                                          typ.ctypes.data_as(c_i32p), cen.ctypes.data_as(c_dp),
                                          clo.ctypes.data_as(c_dp))
        return dict(dist=dist, grad=grad, type=typ, centers=cen, closest=clo, seconds=t)


def _states(q, nvar):
    q = np.asarray(q, dtype=np.float64)
    if nvar == 0:
        n = q.shape[0] if q.ndim >= 1 else 1
        return np.zeros((n, 0)), n
    q = np.ascontiguousarray(q.reshape(-1, nvar))
    return q, len(q)


def hardware_concurrency():
    return lib().orc_hardware_concurrency()
