"""Python harness over the C ABI: thin object wrappers plus the host bookkeeping the reference does on the CPU
(ACM -> per-link tables, group link lists, model assembly).  Used by tests and bench.py; the C++ host mirror of
CollisionEnv lives in moveit_b200/host/.
"""
import ctypes as C

import numpy as np

from . import binding as B
from .robot_description import RobotDescription

DEFAULT_RESOLUTION = 0.02            # collision_env_distance_field.h:52
DEFAULT_COLLISION_TOLERANCE = 0.0    # :53
DEFAULT_MAX_PROPAGATION = 0.25       # :54
DEFAULT_WORLD_PADDING = 0.01         # BodyDecomposition(shape, resolution, padding = 0.01), types.h:227


def _dims3(dims):
    d = np.zeros(3)
    d[:len(dims)] = dims
    return d


def body_decomposition(shapes, resolution, padding):
    """BodyDecomposition via the engine (GPU rasteriser).  shapes: [(type, dims, pose3x4)].
    Returns spheres[ns,4], points[np,3], bsphere[4]."""
    lib = B.load()
    n = len(shapes)
    types = np.array([s[0] for s in shapes], dtype=np.int32)
    dims = np.array([_dims3(s[1]) for s in shapes], dtype=np.float64).reshape(n, 3)
    poses = np.array([np.asarray(s[2], dtype=np.float64) for s in shapes]).reshape(n, 12)
    ns, npt = C.c_int64(0), C.c_int64(0)
    bs = np.zeros(4)
    args = (n, B.ptr(types, C.c_int32), B.ptr(dims, C.c_double), B.ptr(poses, C.c_double), resolution, padding)
    B.check(lib.b200df_body_decomposition(*args, None, 0, C.byref(ns), None, 0, C.byref(npt), B.ptr(bs, C.c_double)))
    sph = np.zeros((max(ns.value, 1), 4))
    pts = np.zeros((max(npt.value, 1), 3))
    B.check(lib.b200df_body_decomposition(*args, B.ptr(sph, C.c_double), ns.value, C.byref(ns),
                                          B.ptr(pts, C.c_double), npt.value, C.byref(npt), B.ptr(bs, C.c_double)))
    return sph[:ns.value], pts[:npt.value], bs


def assemble_model(desc: RobotDescription, decompose, resolution=DEFAULT_RESOLUTION, link_padding=0.0,
                   use_sphere_overrides=True, padding_by_link=None):
    """CollisionEnvDistanceField::addLinkBodyDecompositions (collision_env_distance_field.cpp:1065-1093): one
    BodyDecomposition per link with geometry, spheres replaced by the override table when one is given.
    `decompose(shapes, resolution, padding)` is the BodyDecomposition provider (engine or oracle)."""
    kin = desc.kinematics_arrays()
    L = desc.n_links
    spheres, points, bsphere, has_geom = [], [], [], []
    for i in range(L):
        if desc.shapes[i]:
            pad = (padding_by_link or {}).get(desc.link_names[i], link_padding)
            sph, pts, bs = decompose(desc.shapes[i], resolution, pad)
            name = desc.link_names[i]
            if use_sphere_overrides and name in desc.spheres_override:
                sph = np.array(desc.spheres_override[name], dtype=np.float64).reshape(-1, 4)
            spheres.append(np.asarray(sph, dtype=np.float64).reshape(-1, 4))
            points.append(np.asarray(pts, dtype=np.float64).reshape(-1, 3))
            bsphere.append(np.asarray(bs, dtype=np.float64))
            has_geom.append(1)
        else:
            spheres.append(np.zeros((0, 4)))
            points.append(np.zeros((0, 3)))
            bsphere.append(np.zeros(4))
            has_geom.append(0)
    kin.update(spheres=spheres, points=points, bsphere=bsphere, has_geometry=np.array(has_geom, dtype=np.int32))
    return kin


def resolve_acm(model_desc, group_links, acm_entry):
    """generateDistanceFieldCacheEntry (collision_env_distance_field.cpp:758-848) without attached bodies:
    self_collision_enabled_[n] and intra_group_collision_enabled_[n][n] from getEntry()==ALWAYS (Q12).
    acm_entry: [L,L] int (-1 none, 0 NEVER, 1 ALWAYS) or None for the "acm is null" branch."""
    n = len(group_links)
    self_en = np.ones(n, dtype=np.uint8)
    intra = np.ones((n, n), dtype=np.uint8)
    for i, li in enumerate(group_links):
        if not model_desc["has_geometry"][li]:
            self_en[i] = 0
            intra[i, :] = 0
            continue
        if acm_entry is None:
            continue
        if acm_entry[li, li] == 1:
            self_en[i] = 0
        for j in range(i + 1, n):
            if acm_entry[li, group_links[j]] == 1:
                intra[i, j] = 0
    return self_en, intra


def link_distance_fields(model_desc, group_links, acm_entry, resolution=DEFAULT_RESOLUTION,
                         max_distance=DEFAULT_MAX_PROPAGATION, signed=False):
    """GroupStateRepresentation::link_distance_fields_ (getGroupStateRepresentation,
    collision_env_distance_field.cpp:1187-1210) for the group's links with geometry, plus the ACM gate of
    getSelfProximityGradients (:388-404: an entry that exists and is not NEVER switches link j's field off for link
    i).  Returns (fields, enabled[G,G]) or (None, None) when the ACM is null / empty (the reference skips the term)."""
    geo = [li for li in group_links if model_desc["has_geometry"][li]]
    if acm_entry is None or not (np.asarray(acm_entry) >= 0).any():
        return None, None
    fields = []
    for li in geo:
        bs = np.asarray(model_desc["bsphere"][li], dtype=np.float64)
        diameter = 2 * bs[3]
        origin = bs[:3] - 0.5 * np.array([diameter, diameter, diameter])
        f = Field((diameter, diameter, diameter), resolution, origin, max_distance, signed)
        pts = np.asarray(model_desc["points"][li], dtype=np.float64).reshape(-1, 3)
        if len(pts):
            f.add_points(pts)
        fields.append(f)
    G = len(geo)
    enabled = np.zeros((G, G), dtype=np.uint8)
    for a, la in enumerate(geo):
        for b, lb in enumerate(geo):
            if a != b and acm_entry[la, lb] in (-1, 0):
                enabled[a, b] = 1
    return fields, enabled


def _model_desc_struct(model_desc, keep):
    """model dict -> B.ModelDesc (ctypes); `keep` receives the arrays the struct points into."""
    L = len(model_desc["parent"])

    def arr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a

    sph_off = np.zeros(L + 1, dtype=np.int32)
    pt_off = np.zeros(L + 1, dtype=np.int64)
    for i in range(L):
        sph_off[i + 1] = sph_off[i] + len(model_desc["spheres"][i])
        pt_off[i + 1] = pt_off[i] + len(model_desc["points"][i])
    sph = arr(np.concatenate([np.reshape(s, (-1, 4)) for s in model_desc["spheres"]]) if L else np.zeros((0, 4)),
              np.float64)
    pts = arr(np.concatenate([np.reshape(p, (-1, 3)) for p in model_desc["points"]]) if L else np.zeros((0, 3)),
              np.float64)
    d = B.ModelDesc()
    d.n_links = L
    d.n_vars = int(model_desc["nvar"])
    d.parent = B.ptr(arr(model_desc["parent"], np.int32), C.c_int32)
    d.joint_type = B.ptr(arr(model_desc["joint_type"], np.int32), C.c_int32)
    d.axis = B.ptr(arr(model_desc["axis"], np.float64), C.c_double)
    d.origin = B.ptr(arr(model_desc["origin"], np.float64), C.c_double)
    d.origin_is_identity = B.ptr(arr(model_desc["origin_is_identity"], np.int32), C.c_int32)
    d.var_index = B.ptr(arr(model_desc["var_index"], np.int32), C.c_int32)
    d.mimic_var = B.ptr(arr(model_desc["mimic_var"], np.int32), C.c_int32)
    d.mimic_mult = B.ptr(arr(model_desc["mimic_mult"], np.float64), C.c_double)
    d.mimic_offset = B.ptr(arr(model_desc["mimic_offset"], np.float64), C.c_double)
    d.has_geometry = B.ptr(arr(model_desc["has_geometry"], np.int32), C.c_int32)
    d.sphere_offset = B.ptr(arr(sph_off, np.int32), C.c_int32)
    d.spheres = B.ptr(sph, C.c_double)
    d.bsphere = B.ptr(arr(np.stack(model_desc["bsphere"]), np.float64), C.c_double)
    d.point_offset = B.ptr(arr(pt_off, np.int64), C.c_int64)
    d.points = B.ptr(pts, C.c_double)
    return d


def _group_desc_struct(group_links, self_enabled, intra_enabled, resolution, tolerance, max_propagation, keep):
    gl = np.ascontiguousarray(group_links, dtype=np.int32)
    se = np.ascontiguousarray(self_enabled, dtype=np.uint8)
    ie = np.ascontiguousarray(intra_enabled, dtype=np.uint8)
    keep.extend([gl, se, ie])
    d = B.GroupDesc()
    d.n_group_links = len(gl)
    d.group_links = B.ptr(gl, C.c_int32)
    d.self_enabled = B.ptr(se, C.c_uint8)
    d.intra_enabled = B.ptr(ie, C.c_uint8)
    d.resolution = resolution
    d.collision_tolerance = tolerance
    d.max_propagation_distance = max_propagation
    return d


class MultiGroup:
    """b200df_multi: one (model, group) replicated on several devices of THIS process; batches are sharded over them
    by host threads inside the library (the single-process form of sharding.py's torchrun ranks)."""

    def __init__(self, model_desc, group_links, self_enabled, intra_enabled, devices, resolution=DEFAULT_RESOLUTION,
                 tolerance=DEFAULT_COLLISION_TOLERANCE, max_propagation=DEFAULT_MAX_PROPAGATION):
        keep = []
        md = _model_desc_struct(model_desc, keep)
        gd = _group_desc_struct(group_links, self_enabled, intra_enabled, resolution, tolerance, max_propagation, keep)
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        B.check(B.load().b200df_multi_create(C.byref(md), C.byref(gd), B.ptr(dev, C.c_int32), len(dev), C.byref(h)))
        self.h = h
        self.devices = [int(d) for d in dev]
        self.n_vars = int(model_desc["nvar"])

    def __del__(self):
        if getattr(self, "h", None):
            B.load().b200df_multi_destroy(self.h)
            self.h = None

    def set_field(self, which, field):
        """which: 0 world, 1 self.  Rebuilds `field` if needed and copies it to every replica (peer copies)."""
        B.check(B.load().b200df_multi_set_field(self.h, which, field.h if field is not None else None))

    def replica_d2(self, which, slot, dims):
        f = C.c_void_p()
        B.check(B.load().b200df_multi_field(self.h, which, slot, C.byref(f)))
        out = np.empty(dims, dtype=np.int32)
        B.check(B.load().b200df_field_download_d2(f, B.ptr(out, C.c_int32), 0))
        return out

    def check(self, q, flags=B.CHECK_ROBOT, precision=B.PRECISION_EXACT):
        q, dt = _q_array(q, self.n_vars)
        out = np.empty(len(q), dtype=np.uint8)
        B.check(B.load().b200df_multi_check_batch(self.h, q.ctypes.data_as(C.c_void_p), dt, len(q), flags, precision,
                                                  B.ptr(out, C.c_uint8)))
        return out

    def check_device(self, q_ptrs, q_dtype, counts, verdict_ptrs, flags=B.CHECK_ROBOT, precision=B.PRECISION_EXACT,
                     streams=None):
        n = len(self.devices)
        qa = (C.c_void_p * n)(*[C.c_void_p(p) for p in q_ptrs])
        va = (C.c_void_p * n)(*[C.c_void_p(p) for p in verdict_ptrs])
        sa = (C.c_void_p * n)(*[C.c_void_p(s) if s else None for s in (streams or [None] * n)])
        cnt = np.ascontiguousarray(counts, dtype=np.int64)
        B.check(B.load().b200df_multi_check_batch_device(self.h, qa, q_dtype, B.ptr(cnt, C.c_int64), flags, precision,
                                                         va, sa))


class Model:
    def __init__(self, model_desc):
        self.desc = model_desc
        L = len(model_desc["parent"])
        self.n_links = L
        self.n_vars = int(model_desc["nvar"])
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a

        sph_off = np.zeros(L + 1, dtype=np.int32)
        pt_off = np.zeros(L + 1, dtype=np.int64)
        for i in range(L):
            sph_off[i + 1] = sph_off[i] + len(model_desc["spheres"][i])
            pt_off[i + 1] = pt_off[i] + len(model_desc["points"][i])
        sph = arr(np.concatenate([np.reshape(s, (-1, 4)) for s in model_desc["spheres"]]) if L else np.zeros((0, 4)),
                  np.float64)
        pts = arr(np.concatenate([np.reshape(p, (-1, 3)) for p in model_desc["points"]]) if L else np.zeros((0, 3)),
                  np.float64)
        d = B.ModelDesc()
        d.n_links = L
        d.n_vars = self.n_vars
        d.parent = B.ptr(arr(model_desc["parent"], np.int32), C.c_int32)
        d.joint_type = B.ptr(arr(model_desc["joint_type"], np.int32), C.c_int32)
        d.axis = B.ptr(arr(model_desc["axis"], np.float64), C.c_double)
        d.origin = B.ptr(arr(model_desc["origin"], np.float64), C.c_double)
        d.origin_is_identity = B.ptr(arr(model_desc["origin_is_identity"], np.int32), C.c_int32)
        d.var_index = B.ptr(arr(model_desc["var_index"], np.int32), C.c_int32)
        d.mimic_var = B.ptr(arr(model_desc["mimic_var"], np.int32), C.c_int32)
        d.mimic_mult = B.ptr(arr(model_desc["mimic_mult"], np.float64), C.c_double)
        d.mimic_offset = B.ptr(arr(model_desc["mimic_offset"], np.float64), C.c_double)
        d.has_geometry = B.ptr(arr(model_desc["has_geometry"], np.int32), C.c_int32)
        d.sphere_offset = B.ptr(arr(sph_off, np.int32), C.c_int32)
        d.spheres = B.ptr(sph, C.c_double)
        d.bsphere = B.ptr(arr(np.stack(model_desc["bsphere"]), np.float64), C.c_double)
        d.point_offset = B.ptr(arr(pt_off, np.int64), C.c_int64)
        d.points = B.ptr(pts, C.c_double)
        h = C.c_void_p()
        B.check(B.load().b200df_model_create(C.byref(d), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            B.load().b200df_model_destroy(self.h)
            self.h = None

    def fk(self, q):
        q, dt = _q_array(q, self.n_vars)
        out = np.empty((len(q), self.n_links, 3, 4))
        B.check(B.load().b200df_model_fk(self.h, q.ctypes.data_as(C.c_void_p), dt, len(q), B.ptr(out, C.c_double)))
        return out


def _q_array(q, nvar):
    q = np.asarray(q)
    if nvar == 0:  # a robot without variables: one row per state, no columns
        n = q.shape[0] if q.ndim >= 1 else 1
        return np.zeros((n, 0), dtype=np.float64), B.Q_F64
    if q.dtype == np.float32:
        q = np.ascontiguousarray(q.reshape(-1, nvar))
        return q, B.Q_F32
    q = np.ascontiguousarray(q.reshape(-1, nvar), dtype=np.float64)
    return q, B.Q_F64


class Group:
    def __init__(self, model: Model, group_links, self_enabled, intra_enabled, resolution=DEFAULT_RESOLUTION,
                 tolerance=DEFAULT_COLLISION_TOLERANCE, max_propagation=DEFAULT_MAX_PROPAGATION):
        self.model = model
        gl = np.ascontiguousarray(group_links, dtype=np.int32)
        se = np.ascontiguousarray(self_enabled, dtype=np.uint8)
        ie = np.ascontiguousarray(intra_enabled, dtype=np.uint8)
        d = B.GroupDesc()
        d.n_group_links = len(gl)
        d.group_links = B.ptr(gl, C.c_int32)
        d.self_enabled = B.ptr(se, C.c_uint8)
        d.intra_enabled = B.ptr(ie, C.c_uint8)
        d.resolution = resolution
        d.collision_tolerance = tolerance
        d.max_propagation_distance = max_propagation
        h = C.c_void_p()
        B.check(B.load().b200df_group_create(model.h, C.byref(d), C.byref(h)))
        self.h = h
        n = C.c_int32(0)
        B.check(B.load().b200df_group_sphere_count(self.h, C.byref(n)))
        self.n_spheres = n.value
        self.group_links = gl

    def __del__(self):
        if getattr(self, "h", None):
            B.load().b200df_group_destroy(self.h)
            self.h = None

    def check(self, q, world=None, self_field=None, flags=B.CHECK_ROBOT, precision=B.PRECISION_EXACT, out=None):
        """out: optional preallocated uint8 [n] (a caller's reused verdict buffer)"""
        q, dt = _q_array(q, self.model.n_vars)
        if out is None:
            out = np.empty(len(q), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.shape == (len(q),) and out.flags.c_contiguous
        B.check(B.load().b200df_check_batch(self.h, world.h if world else None, self_field.h if self_field else None,
                                            q.ctypes.data_as(C.c_void_p), dt, len(q), flags, precision,
                                            B.ptr(out, C.c_uint8)))
        return out

    def check_device(self, q_ptr, q_dtype, n, verdict_ptr, world=None, self_field=None, flags=B.CHECK_ROBOT,
                     precision=B.PRECISION_EXACT, stream=None):
        B.check(B.load().b200df_check_batch_device(self.h, world.h if world else None,
                                                   self_field.h if self_field else None, C.c_void_p(q_ptr), q_dtype, n,
                                                   flags, precision, C.c_void_p(verdict_ptr),
                                                   C.c_void_p(stream) if stream else None))

    def gradients(self, q, world=None, self_field=None, flags=B.CHECK_ROBOT | B.CHECK_INTRA, want_centers=True,
                  out=None, out_dtype=np.float64):
        """out: optional dict of preallocated C-contiguous arrays (dist, grad, type, centers), e.g. views of pinned
        host memory, to keep the result copies off the pageable path.  out_dtype=np.float32 takes the
        single-precision-output entry point (b200df_gradients_batch_f32)."""
        q, dt = _q_array(q, self.model.n_vars)
        N, S = len(q), self.n_spheres
        out = out or {}
        ot = np.dtype(out_dtype)
        assert ot in (np.dtype(np.float64), np.dtype(np.float32))
        dist = out.get("dist") if out.get("dist") is not None else np.empty((N, S), dtype=ot)
        grad = out.get("grad") if out.get("grad") is not None else np.empty((N, S, 3), dtype=ot)
        typ = out.get("type") if out.get("type") is not None else np.empty((N, S), dtype=np.uint8)
        cen = (out.get("centers") if out.get("centers") is not None else np.empty((N, S, 3), dtype=ot)) \
            if want_centers else None
        assert dist.shape == (N, S) and grad.shape == (N, S, 3) and typ.shape == (N, S)
        assert dist.dtype == ot and grad.dtype == ot and (cen is None or cen.dtype == ot)
        ct = C.c_double if ot == np.dtype(np.float64) else C.c_float
        fn = B.load().b200df_gradients_batch if ot == np.dtype(np.float64) else B.load().b200df_gradients_batch_f32
        B.check(fn(self.h, world.h if world else None, self_field.h if self_field else None,
                   q.ctypes.data_as(C.c_void_p), dt, N, flags, B.ptr(dist, ct), B.ptr(grad, ct), B.ptr(typ, C.c_uint8),
                   B.ptr(cen, ct) if want_centers else None))
        return dict(dist=dist, grad=grad, type=typ, centers=cen)

    def check_edges(self, qa, qb, segments, world=None, self_field=None, flags=B.CHECK_ROBOT, precision=B.PRECISION_EXACT,
                    var_is_angular=None):
        """M motions qa[i] -> qb[i] cut into `segments` pieces; returns first_invalid[M] (0 = valid).
        var_is_angular[v]: 0 linear, 1 shortest-arc angle (continuous revolute, planar theta), 2 / 3 the first / the
        other three quaternion variables of a floating joint (slerp)."""
        qa = np.ascontiguousarray(qa, dtype=np.float64).reshape(-1, self.model.n_vars)
        qb = np.ascontiguousarray(qb, dtype=np.float64).reshape(-1, self.model.n_vars)
        assert qa.shape == qb.shape
        out = np.zeros(len(qa), dtype=np.int32)
        ang = None
        if var_is_angular is not None:
            ang = np.ascontiguousarray(var_is_angular, dtype=np.uint8)
            assert len(ang) == self.model.n_vars
        B.check(B.load().b200df_check_edges(self.h, world.h if world else None, self_field.h if self_field else None,
                                            B.ptr(qa, C.c_double), B.ptr(qb, C.c_double), len(qa), int(segments),
                                            B.ptr(ang, C.c_uint8) if ang is not None else None, flags, precision,
                                            B.ptr(out, C.c_int32)))
        return out

    def set_link_fields(self, fields, enabled):
        """fields: one engine.Field (or None) per group link WITH geometry, in group order; enabled: [G,G] uint8."""
        n = C.c_int32(0)
        B.check(B.load().b200df_group_link_count(self.h, C.byref(n)))
        if fields is None:
            B.check(B.load().b200df_group_set_link_fields(self.h, None, None))
            self._link_fields = None
            return
        assert len(fields) == n.value
        arr = (C.c_void_p * max(n.value, 1))(*[f.h if f is not None else None for f in fields])
        en = np.ascontiguousarray(enabled, dtype=np.uint8).reshape(n.value, n.value)
        B.check(B.load().b200df_group_set_link_fields(self.h, arr, B.ptr(en, C.c_uint8)))
        self._link_fields = list(fields)  # keep them alive

    def distance(self, q, world):
        q, dt = _q_array(q, self.model.n_vars)
        out = np.empty(len(q))
        B.check(B.load().b200df_distance_batch(self.h, world.h, q.ctypes.data_as(C.c_void_p), dt, len(q),
                                               B.ptr(out, C.c_double)))
        return out

    def sphere_centers(self, q):
        q, dt = _q_array(q, self.model.n_vars)
        out = np.empty((len(q), self.n_spheres, 3))
        B.check(B.load().b200df_sphere_centers(self.h, q.ctypes.data_as(C.c_void_p), dt, len(q), B.ptr(out, C.c_double)))
        return out


class Field:
    """PropagationDistanceField replacement.  origin = grid CORNER (as the reference constructor)."""

    def __init__(self, size, resolution, origin, max_distance, signed=False):
        d = B.FieldDesc()
        d.size = (C.c_double * 3)(*[float(s) for s in size])
        d.origin = (C.c_double * 3)(*[float(o) for o in origin])
        d.resolution = resolution
        d.max_distance = max_distance
        d.propagate_negative = int(signed)
        h = C.c_void_p()
        B.check(B.load().b200df_field_create(C.byref(d), C.byref(h)))
        self.h = h
        dims = (C.c_int32 * 4)()
        B.check(B.load().b200df_field_dims(self.h, dims))
        self.dims = (dims[0], dims[1], dims[2])
        self.max_d2 = dims[3]
        self.resolution = resolution
        self.size = tuple(float(v) for v in size)
        self.origin = tuple(float(v) for v in origin)

    @classmethod
    def centered(cls, size, resolution, center, max_distance, signed=False):
        """CollisionEnvDistanceField's convention: origin argument is the field CENTRE (Q14)."""
        corner = [center[i] - 0.5 * size[i] for i in range(3)]
        return cls(size, resolution, corner, max_distance, signed)

    def __del__(self):
        if getattr(self, "h", None):
            B.load().b200df_field_destroy(self.h)
            self.h = None

    def reset(self):
        B.check(B.load().b200df_field_reset(self.h))

    def add_points(self, pts):
        p = B.as_f64(np.reshape(pts, (-1, 3)))
        B.check(B.load().b200df_field_add_points(self.h, B.ptr(p, C.c_double), len(p)))

    def add_points_device(self, ptr, n, stream=None):
        """pts already on the GPU: fp64 [n][3] device pointer."""
        B.check(B.load().b200df_field_add_points_device(self.h, C.c_void_p(ptr), n, C.c_void_p(stream) if stream else None))

    def remove_points(self, pts):
        p = B.as_f64(np.reshape(pts, (-1, 3)))
        B.check(B.load().b200df_field_remove_points(self.h, B.ptr(p, C.c_double), len(p)))

    def update_points(self, old, new):
        a = B.as_f64(np.reshape(old, (-1, 3)))
        b = B.as_f64(np.reshape(new, (-1, 3)))
        B.check(B.load().b200df_field_update_points(self.h, B.ptr(a, C.c_double), len(a), B.ptr(b, C.c_double), len(b)))

    def add_voxels(self, ijk):
        a = np.ascontiguousarray(np.reshape(ijk, (-1, 3)), dtype=np.int32)
        B.check(B.load().b200df_field_add_voxels(self.h, B.ptr(a, C.c_int32), len(a)))

    def add_shape(self, shape_type, dims, pose, scale=1.0, padding=0.0):
        d = _dims3(dims)
        p = B.as_f64(pose).reshape(12)
        B.check(B.load().b200df_field_add_shape(self.h, shape_type, B.ptr(d, C.c_double), B.ptr(p, C.c_double), scale,
                                                padding))

    def remove_shape(self, shape_type, dims, pose, scale=1.0, padding=0.0):
        d = _dims3(dims)
        p = B.as_f64(pose).reshape(12)
        B.check(B.load().b200df_field_remove_shape(self.h, shape_type, B.ptr(d, C.c_double), B.ptr(p, C.c_double),
                                                   scale, padding))

    def add_posed_body(self, shape_type, dims, pose, padding=DEFAULT_WORLD_PADDING):
        d = _dims3(dims)
        p = B.as_f64(pose).reshape(12)
        B.check(B.load().b200df_field_add_posed_body(self.h, shape_type, B.ptr(d, C.c_double), B.ptr(p, C.c_double),
                                                     padding))

    def add_octree_leaves(self, leaves):
        a = B.as_f64(np.reshape(leaves, (-1, 4)))
        B.check(B.load().b200df_field_add_octree_leaves(self.h, B.ptr(a, C.c_double), len(a)))

    def remove_octree_leaves(self, leaves):
        a = B.as_f64(np.reshape(leaves, (-1, 4)))
        B.check(B.load().b200df_field_remove_octree_leaves(self.h, B.ptr(a, C.c_double), len(a)))

    def rebuild(self):
        ms = C.c_float(0)
        B.check(B.load().b200df_field_rebuild(self.h, C.byref(ms)))
        return ms.value

    def d2(self, negative=False):
        out = np.empty(self.dims, dtype=np.int32)
        B.check(B.load().b200df_field_download_d2(self.h, B.ptr(out, C.c_int32), int(negative)))
        return out

    def upload_d2(self, d2, neg_d2=None):
        """Reference-field mode: serve caller-supplied squared distances (e.g. the reference's own propagation
        result) until the next occupancy edit."""
        a = np.ascontiguousarray(d2, dtype=np.int32).reshape(self.dims)
        b = None if neg_d2 is None else np.ascontiguousarray(neg_d2, dtype=np.int32).reshape(self.dims)
        B.check(B.load().b200df_field_upload_d2(self.h, B.ptr(a, C.c_int32), B.ptr(b, C.c_int32) if b is not None else None))

    def values_uploaded(self):
        n = C.c_int32(0)
        B.check(B.load().b200df_field_values_uploaded(self.h, C.byref(n)))
        return bool(n.value)

    def nearest_cells(self, ijk):
        a = np.ascontiguousarray(np.reshape(ijk, (-1, 3)), dtype=np.int32)
        n = len(a)
        dist, pos, found = np.empty(n), np.empty((n, 3), dtype=np.int32), np.empty(n, dtype=np.uint8)
        B.check(B.load().b200df_field_nearest_cells(self.h, B.ptr(a, C.c_int32), n, B.ptr(dist, C.c_double),
                                                    B.ptr(pos, C.c_int32), B.ptr(found, C.c_uint8)))
        return dist, pos, found.astype(bool)

    def occupancy(self):
        out = np.empty(self.dims, dtype=np.uint8)
        B.check(B.load().b200df_field_download_occupancy(self.h, B.ptr(out, C.c_uint8)))
        return out

    def distance_gradient(self, pts):
        p = B.as_f64(np.reshape(pts, (-1, 3)))
        n = len(p)
        dist = np.empty(n)
        grad = np.empty((n, 3))
        inb = np.empty(n, dtype=np.uint8)
        B.check(B.load().b200df_field_distance_gradient(self.h, B.ptr(p, C.c_double), n, B.ptr(dist, C.c_double),
                                                        B.ptr(grad, C.c_double), B.ptr(inb, C.c_uint8)))
        return dist, grad, inb.astype(bool)

    def pack_occupancy(self):
        n = C.c_int64(0)
        B.check(B.load().b200df_field_pack_occupancy(self.h, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint8)
        B.check(B.load().b200df_field_pack_occupancy(self.h, B.ptr(out, C.c_uint8), n.value, C.byref(n)))
        return out

    def unpack_occupancy(self, payload):
        b = np.ascontiguousarray(payload, dtype=np.uint8)
        B.check(B.load().b200df_field_unpack_occupancy(self.h, B.ptr(b, C.c_uint8), len(b)))

    def write_stream(self, fileobj):
        """PropagationDistanceField::writeToStream (propagation_distance_field.cpp:636-673)."""
        from . import field_stream
        return field_stream.write_stream(fileobj, self.resolution, self.size, self.origin, self.pack_occupancy().tobytes())

    @classmethod
    def from_stream(cls, fileobj, max_distance, signed=False):
        """PropagationDistanceField(std::istream&, max_distance, propagate_negative) (:64-74, 675-761); None when the
        stream does not hold a field."""
        from . import field_stream
        got = field_stream.read_stream(fileobj)
        if got is None:
            return None
        h, payload = got
        f = cls((h["size_x"], h["size_y"], h["size_z"]), h["resolution"], (h["origin_x"], h["origin_y"], h["origin_z"]),
                max_distance, signed)
        if len(payload) < field_stream.payload_size(f.dims):
            return None
        f.unpack_occupancy(np.frombuffer(payload, dtype=np.uint8)[:field_stream.payload_size(f.dims)])
        return f

    def device_arrays(self):
        d2, neg = C.c_void_p(), C.c_void_p()
        nb, es = C.c_int64(0), C.c_int32(0)
        B.check(B.load().b200df_field_device_arrays(self.h, C.byref(d2), C.byref(neg), C.byref(nb), C.byref(es)))
        return d2.value, neg.value, nb.value, es.value

    def mark_built(self):
        B.check(B.load().b200df_field_mark_built(self.h))

    def copy_from(self, other):
        B.check(B.load().b200df_field_copy_from(self.h, other.h))

    def export_device(self, occ_ptr=None, d2_ptr=None, neg_ptr=None, stream=None):
        B.check(B.load().b200df_field_export_device(self.h, occ_ptr, d2_ptr, neg_ptr, stream))

    def import_device(self, occ_ptr=None, d2_ptr=None, neg_ptr=None, stream=None):
        B.check(B.load().b200df_field_import_device(self.h, occ_ptr, d2_ptr, neg_ptr, stream))
