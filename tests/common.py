"""Builds the same workload twice: once in the CPU oracle (checker) and once in the engine (thing under test)."""
import numpy as np

from moveit_b200 import scenarios
from moveit_b200.engine import DEFAULT_WORLD_PADDING, assemble_model, resolve_acm


def oracle_world_points(oracle, objects, resolution, padding=DEFAULT_WORLD_PADDING):
    """World-object route of CollisionEnvDistanceField::updateDistanceObject (cpp:1749-1798): BodyDecomposition in the
    body frame (default padding 0.01), then PosedBodyPointDecomposition(bd, pose)."""
    out = []
    for shape_type, dims, pose in objects:
        _, pts, _ = oracle.body_decomposition([(shape_type, dims, oracle.pose12())], resolution, padding)
        pose = np.asarray(pose)
        R, t = pose[:, :3], pose[:, 3]
        # Isometry * Vector3: t + ((r0*x + r1*y) + r2*z), same order as the oracle's apply()
        w = t[None, :] + ((R[None, :, 0] * pts[:, 0:1] + R[None, :, 1] * pts[:, 1:2]) + R[None, :, 2] * pts[:, 2:3])
        out.append(w)
    return np.concatenate(out) if out else np.zeros((0, 3))


class OracleSetup:
    """exact_field=True: the oracle's fields hold the EXACT clamped EDT of the reference's obstacle voxels — what the
    engine must match bit for bit.  exact_field=False: the reference's own (inexact, F4) propagation result, used for
    the mismatch census."""

    def __init__(self, oracle, desc, field, objects=(), group="", signed=False, self_state=None, use_acm=True,
                 sphere_overrides=True, tol=0.0, exact_field=True, padding_by_link=None, acm_override=None,
                 group_links=None):
        self.desc = desc
        self.model_desc = assemble_model(desc, oracle.body_decomposition, field["resolution"],
                                         use_sphere_overrides=sphere_overrides, padding_by_link=padding_by_link)
        self.model = oracle.Model(self.model_desc)
        self.env = oracle.Env(self.model, field["size"], field["center"], res=field["resolution"], tol=tol,
                              max_prop=field["max_distance"], signed=signed)
        self.group_links = group_links if group_links is not None else desc.group_links(group)
        self.acm = (acm_override if acm_override is not None else desc.acm_entry_matrix()) if use_acm else None
        self.env.set_group(self.group_links, self.acm, self_state)
        self.world_points = oracle_world_points(oracle, objects, field["resolution"])
        if len(self.world_points):
            self.env.world.add_points(self.world_points)
        if exact_field:
            self.env.world.make_exact()
            sf = self.env.self_field()
            if sf is not None:
                sf.make_exact()


class EngineSetup:
    def __init__(self, desc, field, objects=(), group="", signed=False, self_points=None, use_acm=True,
                 sphere_overrides=True, model_desc=None, tol=0.0):
        from moveit_b200 import engine
        self.desc = desc
        self.model_desc = model_desc or assemble_model(desc, engine.body_decomposition, field["resolution"],
                                                       use_sphere_overrides=sphere_overrides)
        self.model = engine.Model(self.model_desc)
        self.group_links = desc.group_links(group)
        acm = desc.acm_entry_matrix() if use_acm else None
        se, ie = resolve_acm(self.model_desc, self.group_links, acm)
        self.group = engine.Group(self.model, self.group_links, se, ie, field["resolution"], tol,
                                  field["max_distance"])
        self.world = engine.Field.centered(field["size"], field["resolution"], field["center"], field["max_distance"],
                                           signed)
        for shape_type, dims, pose in objects:
            self.world.add_posed_body(shape_type, dims, pose)
        self.self_field = None
        if self_points is not None:
            self.self_field = engine.Field.centered(field["size"], field["resolution"], field["center"],
                                                    field["max_distance"], signed)
            self.self_field.add_points(self_points)


def self_field_points(model_desc, desc, group_links, fk_transforms):
    """generateDistanceFieldCacheEntry :922-970: points of every link with geometry outside the group, posed."""
    in_group = set(group_links)
    pts = []
    for li in range(desc.n_links):
        if not model_desc["has_geometry"][li] or li in in_group:
            continue
        T = fk_transforms[li]
        p = model_desc["points"][li]
        R, t = T[:, :3], T[:, 3]
        pts.append(t[None, :] + ((R[None, :, 0] * p[:, 0:1] + R[None, :, 1] * p[:, 1:2]) + R[None, :, 2] * p[:, 2:3]))
    return np.concatenate(pts) if pts else np.zeros((0, 3))


FIELD_SMALL = dict(size=(1.0, 1.0, 1.0), resolution=0.1, center=(0.5, 0.5, 0.5), max_distance=0.3)
__all__ = ["OracleSetup", "EngineSetup", "oracle_world_points", "self_field_points", "scenarios", "FIELD_SMALL"]
