"""Builds the same workload twice: once in the CPU oracle (checker) and once in the engine (thing under test)."""
import numpy as np

from moveit_b200 import robot_description as rd
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
def random_robot(seed, n_links, floating_root=False, pair_off=0.5, spread=(0.12, 0.3), radius=(0.012, 0.03),
                  p_identity=0.25, max_spheres=9, dense_tail=0, p_geometry=0.85):
    """A random kinematic tree: every joint type, oblique axes, identity and general origins, branches, links without
    geometry, 1-9 spheres per link (batch-of-4 boundaries), mimic joints, a random ACM."""
    import math
    rng = np.random.default_rng(seed)
    r = rd.RobotDescription("fuzz%d" % seed)
    one_var = []  # (joint name, type) of earlier one-variable joints, candidates for mimicking
    for i in range(n_links):
        name, jname = "l%d" % i, "j%d" % i
        parent = None if i == 0 else "l%d" % (i - 1 if rng.random() < 0.6 else rng.integers(0, i))
        u = rng.random()
        if i == 0:
            jt = rd.JOINT_FLOATING if floating_root else (rd.JOINT_PLANAR if u < 0.5 else rd.JOINT_FIXED)
        else:
            jt = rd.JOINT_FIXED if u < 0.2 else rd.JOINT_REVOLUTE if u < 0.75 else rd.JOINT_PRISMATIC if u < 0.93 \
                else rd.JOINT_PLANAR
        axis = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, -1), (0, -1, 0),
                tuple(rng.normal(size=3))][rng.integers(0, 6)]
        ident = rng.random() < p_identity
        xyz = (0, 0, 0) if ident else tuple(rng.uniform(spread[0], spread[1], 3) * rng.choice([-1.0, 1.0], 3))
        rpy = (0, 0, 0) if ident else tuple(rng.uniform(-math.pi, math.pi, 3))
        limits, mimic = None, None
        if jt == rd.JOINT_REVOLUTE:
            limits = (-2.5, 2.5)
        elif jt == rd.JOINT_PRISMATIC:
            limits = (-0.2, 0.3)
        elif jt == rd.JOINT_PLANAR:
            limits = [(-0.4, 0.4), (-0.4, 0.4), (-math.pi, math.pi)]
        if jt in (rd.JOINT_REVOLUTE, rd.JOINT_PRISMATIC):
            same = [j for j, t in one_var if t == jt]
            if same and rng.random() < 0.12:
                mimic = (same[rng.integers(0, len(same))], float(rng.uniform(-1.5, 1.5)), float(rng.uniform(-0.1, 0.1)))
            else:
                one_var.append((jname, jt))
        r.add_link(name, parent, jname, jt, xyz, rpy, axis=axis, limits=limits, mimic=mimic)
        if rng.random() < p_geometry:
            r.add_shape(name, rd.SHAPE_BOX, (0.1, 0.08, 0.06))
            k = int(rng.integers(1, max_spheres + 1))
            r.set_spheres(name, [(rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05),
                                  rng.uniform(radius[0], radius[1])) for _ in range(k)])
    geom = [r.link_names[i] for i in range(n_links) if r.shapes[i]]
    for i in range(n_links):  # parent-child pairs never collide; some random extra pairs are allowed to as well
        if r.parent[i] >= 0:
            r.disabled_pairs.append((r.link_names[r.parent[i]], r.link_names[i]))
    for a in range(len(geom)):
        for b in range(a + 1, len(geom)):
            if b < len(geom) - dense_tail and rng.random() < pair_off:  # the last dense_tail links keep every pair
                r.disabled_pairs.append((geom[a], geom[b]))
    return r


FUZZ_CASES = [  # seed, links, floating root, generator settings (tuned so that each check has both outcomes)
    (1, 12, False, dict(pair_off=0.6, p_identity=0.1)),
    (2, 25, False, dict(pair_off=0.8, p_identity=0.1)),
    (3, 40, False, dict(pair_off=0.8)),
    (4, 40, False, dict(pair_off=0.0, spread=(0.35, 0.7), radius=(0.01, 0.022), p_identity=0.0)),  # > 32 partners
    (5, 18, True, dict(pair_off=0.7, p_identity=0.1)),
    (6, 33, True, dict(pair_off=0.8)),
    # fits FAST's tables (<= 128 spheres, <= 512 link pairs) AND has links with more than 32 partners
    (7, 40, False, dict(pair_off=0.85, max_spheres=3, dense_tail=3, spread=(0.25, 0.5), radius=(0.02, 0.04),
                        p_geometry=0.97)),
]


__all__ = ["OracleSetup", "EngineSetup", "oracle_world_points", "self_field_points", "scenarios", "FIELD_SMALL",
           "random_robot", "FUZZ_CASES"]
