"""ctypes binding of the C ABI in include/b200df.h (libb200df.so, built in-tree by moveit_b200/build.py).

This is plumbing for tests, bench.py and the Python harness; the product is the library.  There is no fallback:
a missing library or a missing CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200DF_LIB: a differently built copy of the same library (kernel tuning experiments, tools/build_variant.py)
LIB_PATH = os.environ.get("B200DF_LIB") or os.path.join(_HERE, "libb200df.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
CHECK_ROBOT, CHECK_SELF, CHECK_INTRA, CHECK_LINK_FIELDS = 1, 2, 4, 8
Q_F32, Q_F64 = 0, 1
PRECISION_EXACT, PRECISION_FAST, PRECISION_DEBUG_SWEEP, PRECISION_DEBUG_SCREEN = 0, 1, 2, 3
ROUTE_EXACT, ROUTE_FAST, ROUTE_SMALL = 0, 1, 2
NOFAST_TABLES, NOFAST_SIGNED, NOFAST_BATCH, NOFAST_SMEM = 1, 2, 4, 8
TYPE_NONE, TYPE_SELF, TYPE_INTRA, TYPE_ENVIRONMENT = 0, 1, 2, 3

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_fp = C.POINTER(C.c_float)


class B200dfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200df error %d: %s" % (code, msg))
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("n_vars", C.c_int32), ("parent", c_i32p), ("joint_type", c_i32p),
                ("axis", c_dp), ("origin", c_dp), ("origin_is_identity", c_i32p), ("var_index", c_i32p),
                ("mimic_var", c_i32p), ("mimic_mult", c_dp), ("mimic_offset", c_dp), ("has_geometry", c_i32p),
                ("sphere_offset", c_i32p), ("spheres", c_dp), ("bsphere", c_dp), ("point_offset", c_i64p),
                ("points", c_dp)]


class GroupDesc(C.Structure):
    _fields_ = [("n_group_links", C.c_int32), ("group_links", c_i32p), ("self_enabled", c_u8p),
                ("intra_enabled", c_u8p), ("resolution", C.c_double), ("collision_tolerance", C.c_double),
                ("max_propagation_distance", C.c_double)]


class FieldDesc(C.Structure):
    _fields_ = [("size", C.c_double * 3), ("origin", C.c_double * 3), ("resolution", C.c_double),
                ("max_distance", C.c_double), ("propagate_negative", C.c_int32)]


# every symbol include/b200df.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_VPP = C.POINTER(C.c_void_p)
SYMBOLS = {
    "b200df_last_error": (C.c_char_p, []),
    "b200df_version": (C.c_int, []),
    "b200df_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200df_set_device": (C.c_int, [C.c_int]),
    "b200df_model_create": (C.c_int, [C.POINTER(ModelDesc), _VPP]),
    "b200df_model_destroy": (C.c_int, [_VP]),
    "b200df_model_fk": (C.c_int, [_VP, _VP, C.c_int, C.c_int64, c_dp]),
    "b200df_body_decomposition": (C.c_int, [C.c_int32, c_i32p, c_dp, c_dp, C.c_double, C.c_double, c_dp, C.c_int64,
                                            c_i64p, c_dp, C.c_int64, c_i64p, c_dp]),
    "b200df_group_create": (C.c_int, [_VP, C.POINTER(GroupDesc), _VPP]),
    "b200df_group_destroy": (C.c_int, [_VP]),
    "b200df_group_sphere_count": (C.c_int, [_VP, c_i32p]),
    "b200df_group_link_count": (C.c_int, [_VP, c_i32p]),
    "b200df_group_set_link_fields": (C.c_int, [_VP, _VPP, c_u8p]),
    "b200df_field_create": (C.c_int, [C.POINTER(FieldDesc), _VPP]),
    "b200df_field_destroy": (C.c_int, [_VP]),
    "b200df_field_dims": (C.c_int, [_VP, c_i32p]),
    "b200df_field_get_desc": (C.c_int, [_VP, C.POINTER(FieldDesc)]),
    "b200df_field_reset": (C.c_int, [_VP]),
    "b200df_field_add_points": (C.c_int, [_VP, c_dp, C.c_int64]),
    "b200df_field_add_points_device": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "b200df_field_remove_points": (C.c_int, [_VP, c_dp, C.c_int64]),
    "b200df_field_update_points": (C.c_int, [_VP, c_dp, C.c_int64, c_dp, C.c_int64]),
    "b200df_field_add_voxels": (C.c_int, [_VP, c_i32p, C.c_int64]),
    "b200df_field_add_shape": (C.c_int, [_VP, C.c_int32, c_dp, c_dp, C.c_double, C.c_double]),
    "b200df_field_remove_shape": (C.c_int, [_VP, C.c_int32, c_dp, c_dp, C.c_double, C.c_double]),
    "b200df_field_add_posed_body": (C.c_int, [_VP, C.c_int32, c_dp, c_dp, C.c_double]),
    "b200df_field_remove_posed_body": (C.c_int, [_VP, C.c_int32, c_dp, c_dp, C.c_double]),
    "b200df_field_add_octree_leaves": (C.c_int, [_VP, c_dp, C.c_int64]),
    "b200df_field_remove_octree_leaves": (C.c_int, [_VP, c_dp, C.c_int64]),
    "b200df_field_rebuild": (C.c_int, [_VP, c_fp]),
    "b200df_field_download_d2": (C.c_int, [_VP, c_i32p, C.c_int]),
    "b200df_field_download_occupancy": (C.c_int, [_VP, c_u8p]),
    "b200df_field_distance_gradient": (C.c_int, [_VP, c_dp, C.c_int64, c_dp, c_dp, c_u8p]),
    "b200df_field_pack_occupancy": (C.c_int, [_VP, c_u8p, C.c_int64, c_i64p]),
    "b200df_field_unpack_occupancy": (C.c_int, [_VP, c_u8p, C.c_int64]),
    "b200df_field_device_arrays": (C.c_int, [_VP, _VPP, _VPP, c_i64p, c_i32p]),
    "b200df_field_mark_built": (C.c_int, [_VP]),
    "b200df_field_copy_from": (C.c_int, [_VP, _VP]),
    "b200df_field_export_device": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "b200df_field_import_device": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "b200df_field_upload_d2": (C.c_int, [_VP, c_i32p, c_i32p]),
    "b200df_field_values_uploaded": (C.c_int, [_VP, c_i32p]),
    "b200df_field_nearest_cells": (C.c_int, [_VP, c_i32p, C.c_int64, c_dp, c_i32p, c_u8p]),
    "b200df_check_batch": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int64, C.c_int, C.c_int, c_u8p]),
    "b200df_check_batch_device": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int64, C.c_int, C.c_int, _VP, _VP]),
    "b200df_check_route": (C.c_int, [_VP, _VP, _VP, C.c_int, C.c_int64, C.c_int, C.c_int, c_i32p, c_i32p]),
    "b200df_check_edges": (C.c_int, [_VP, _VP, _VP, c_dp, c_dp, C.c_int64, C.c_int32, c_u8p, C.c_int, C.c_int, c_i32p]),
    "b200df_gradients_batch": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int64, C.c_int, c_dp, c_dp, c_u8p, c_dp]),
    "b200df_gradients_batch_device": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int64, C.c_int, _VP, _VP, _VP, _VP, _VP]),
    "b200df_gradients_batch_f32": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int64, C.c_int, c_fp, c_fp, c_u8p, c_fp]),
    "b200df_gradients_batch_device_f32": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int64, C.c_int, _VP, _VP, _VP, _VP,
                                                    _VP]),
    "b200df_distance_batch": (C.c_int, [_VP, _VP, _VP, C.c_int, C.c_int64, c_dp]),
    "b200df_sphere_centers": (C.c_int, [_VP, _VP, C.c_int, C.c_int64, c_dp]),
    "b200df_multi_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(GroupDesc), c_i32p, C.c_int32, _VPP]),
    "b200df_multi_destroy": (C.c_int, [_VP]),
    "b200df_multi_device_count": (C.c_int, [_VP, c_i32p]),
    "b200df_multi_set_field": (C.c_int, [_VP, C.c_int, _VP]),
    "b200df_multi_field": (C.c_int, [_VP, C.c_int, C.c_int32, _VPP]),
    "b200df_multi_check_batch": (C.c_int, [_VP, _VP, C.c_int, C.c_int64, C.c_int, C.c_int, c_u8p]),
    "b200df_multi_check_batch_device": (C.c_int, [_VP, _VPP, C.c_int, c_i64p, C.c_int, C.c_int, _VPP, _VPP]),
    "b200df_host_alloc": (C.c_int, [C.c_size_t, _VPP]),
    "b200df_host_free": (C.c_int, [_VP]),
    "b200df_launch_count": (C.c_int64, []),
}

_lib = None


def load():
    """Loads libb200df.so (building is the job of __graft_entry__.build / moveit_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libb200df.so is not built: run `python -m moveit_b200.build` (needs nvcc); "
                              "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        raise B200dfError(rc, load().b200df_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    rc = load().b200df_device_count(C.byref(n))
    return n.value if rc == OK else 0


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))
