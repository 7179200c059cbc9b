"""Kinematic + collision-geometry description of a robot, in the flat form the C-ABI (include/b200df.h) takes.

This is host-side input data, not an algorithm: it plays the role of moveit::core::RobotModel for the
distance-field path (link order = DFS pre-order with parent index < child index, robot_model.cpp:783-838;
one variable vector per state, robot_state.h).  URDF/SRDF parsing is out of scope.

Panda constants: public franka_description / moveit_resources_panda_description values;
they are NOT in /root/reference.  Collision geometry of the real package is meshes, which the distance-field
path cannot use directly either way: spheres come from the reference's own ``link_body_decompositions``
override (collision_env_distance_field.cpp:1085-1088) and link points / bounding spheres from primitive shapes.
"""
import math

import numpy as np

JOINT_FIXED, JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_PLANAR, JOINT_FLOATING = 0, 1, 2, 3, 4
SHAPE_SPHERE, SHAPE_CYLINDER, SHAPE_BOX = 1, 2, 4  # shapes::ShapeType values used by the C-ABI
_NVARS = {JOINT_FIXED: 0, JOINT_REVOLUTE: 1, JOINT_PRISMATIC: 1, JOINT_PLANAR: 3, JOINT_FLOATING: 7}


def rpy_xyz_to_pose(xyz=(0.0, 0.0, 0.0), rpy=(0.0, 0.0, 0.0)):
    """URDF origin -> 3x4 pose the way urdfdom + MoveIt do it: rpy -> quaternion (urdf::Rotation::setFromRPY),
    quaternion -> matrix (Eigen::Quaterniond::toRotationMatrix)."""
    phi, the, psi = rpy[0] / 2.0, rpy[1] / 2.0, rpy[2] / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    y = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    T = np.zeros((3, 4))
    T[0, 0] = 1.0 - (tyy + tzz)
    T[0, 1] = txy - twz
    T[0, 2] = txz + twy
    T[1, 0] = txy + twz
    T[1, 1] = 1.0 - (txx + tzz)
    T[1, 2] = tyz - twx
    T[2, 0] = txz - twy
    T[2, 1] = tyz + twx
    T[2, 2] = 1.0 - (txx + tyy)
    T[:, 3] = xyz
    return T


def quat_xyz_to_pose(xyz, quat_xyzw):
    x, y, z, w = quat_xyzw
    n = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    T = np.zeros((3, 4))
    T[0, :3] = (1.0 - (tyy + tzz), txy - twz, txz + twy)
    T[1, :3] = (txy + twz, 1.0 - (txx + tzz), tyz - twx)
    T[2, :3] = (txz - twy, tyz + twx, 1.0 - (txx + tyy))
    T[:, 3] = xyz
    return T


class RobotDescription:
    """Ordered list of links, each with its parent joint.  Links must be added parent-before-child."""

    def __init__(self, name):
        self.name = name
        self.link_names = []
        self.joint_names = []
        self.parent = []
        self.joint_type = []
        self.axis = []
        self.origin = []
        self.var_index = []
        self.mimic = []  # (source joint name, multiplier, offset) or None
        self.limits = []  # per variable (lo, hi)
        self.variable_names = []
        self.shapes = []  # per link: list of (type, dims, pose3x4)
        self.spheres_override = {}  # link name -> [(x, y, z, r)]
        self.disabled_pairs = []  # SRDF disable_collisions pairs (link names)
        self.groups = {}  # group name -> list of joint names

    def add_link(self, link, parent_link, joint, joint_type, xyz=(0, 0, 0), rpy=(0, 0, 0), axis=(0, 0, 1),
                 limits=None, mimic=None, origin_pose=None):
        if parent_link is None:
            p = -1
        else:
            p = self.link_names.index(parent_link)
        self.link_names.append(link)
        self.joint_names.append(joint)
        self.parent.append(p)
        self.joint_type.append(joint_type)
        self.axis.append(tuple(float(a) for a in axis))
        self.origin.append(np.array(origin_pose, dtype=np.float64) if origin_pose is not None else rpy_xyz_to_pose(xyz, rpy))
        nv = _NVARS[joint_type]
        if nv:
            self.var_index.append(len(self.variable_names))
            if nv == 1:
                self.variable_names.append(joint)
                self.limits.append(limits if limits is not None else (-math.pi, math.pi))
            elif joint_type == JOINT_PLANAR:
                self.variable_names += [joint + "/x", joint + "/y", joint + "/theta"]
                self.limits += list(limits) if limits is not None else [(-1.0, 1.0), (-1.0, 1.0), (-math.pi, math.pi)]
            else:
                self.variable_names += [joint + s for s in ("/trans_x", "/trans_y", "/trans_z", "/rot_x", "/rot_y",
                                                            "/rot_z", "/rot_w")]
                self.limits += [(-1.0, 1.0)] * 7
        else:
            self.var_index.append(-1)
        self.mimic.append(mimic)
        self.shapes.append([])
        return self

    def add_shape(self, link, shape_type, dims, xyz=(0, 0, 0), rpy=(0, 0, 0)):
        self.shapes[self.link_names.index(link)].append((shape_type, tuple(float(d) for d in dims),
                                                         rpy_xyz_to_pose(xyz, rpy)))
        return self

    def set_spheres(self, link, spheres):
        self.spheres_override[link] = [tuple(float(v) for v in s) for s in spheres]
        return self

    @property
    def n_links(self):
        return len(self.link_names)

    @property
    def nvar(self):
        return len(self.variable_names)

    def kinematics_arrays(self):
        """Flat arrays for b200df_model_create / the oracle."""
        L = self.n_links
        origin = np.stack(self.origin) if L else np.zeros((0, 3, 4))
        ident = np.array([int(np.array_equal(o, np.hstack([np.eye(3), np.zeros((3, 1))]))) for o in self.origin],
                         dtype=np.int32)
        mimic_var = np.full(L, -1, dtype=np.int32)
        mimic_mult = np.ones(L)
        mimic_off = np.zeros(L)
        for i, m in enumerate(self.mimic):
            if m is not None:
                src = self.joint_names.index(m[0])
                mimic_var[i] = self.var_index[src]
                mimic_mult[i] = m[1]
                mimic_off[i] = m[2]
        return dict(parent=np.array(self.parent, dtype=np.int32), joint_type=np.array(self.joint_type, dtype=np.int32),
                    axis=np.array(self.axis, dtype=np.float64).reshape(L, 3), origin=origin, origin_is_identity=ident,
                    var_index=np.array(self.var_index, dtype=np.int32), mimic_var=mimic_var, mimic_mult=mimic_mult,
                    mimic_offset=mimic_off, nvar=self.nvar)

    def links_with_geometry(self):
        return [i for i in range(self.n_links) if self.shapes[i]]

    def descendant_links_of_joints(self, joint_names):
        """JointModelGroup::getUpdatedLinkModelNames (joint_model_group.cpp:223-247): every link moved by the
        group's joints, in model order."""
        moved = set()
        for i in range(self.n_links):
            if self.joint_names[i] in joint_names or (self.parent[i] >= 0 and self.parent[i] in moved):
                moved.add(i)
        return sorted(moved)

    def group_links(self, group_name):
        """Link indices a CollisionRequest.group_name selects ("" = all links,
        collision_env_distance_field.cpp:736-739)."""
        if group_name == "":
            return list(range(self.n_links))
        return self.descendant_links_of_joints(set(self.groups[group_name]))

    def acm_entry_matrix(self, extra_disabled=()):
        """AllowedCollisionMatrix as the distance-field env reads it (getEntry semantics, Q12): [L,L] int32 with
        -1 = no entry, 0 = NEVER, 1 = ALWAYS.  Mirrors AllowedCollisionMatrix(names, false) followed by
        setEntry(a, b, true) for each SRDF disable_collisions pair (planning_scene.cpp default ACM)."""
        L = self.n_links
        acm = np.zeros((L, L), dtype=np.int32)
        for a, b in list(self.disabled_pairs) + list(extra_disabled):
            i, j = self.link_names.index(a), self.link_names.index(b)
            acm[i, j] = acm[j, i] = 1
        return acm

    def sample_states(self, n, seed, dtype=np.float64):
        """Uniform samples inside the variable limits from the portable SplitMix64 stream."""
        from .rng import uniform01
        u = uniform01(seed, n * self.nvar).reshape(n, self.nvar)
        lo = np.array([l[0] for l in self.limits])
        hi = np.array([l[1] for l in self.limits])
        q = lo + (hi - lo) * u
        for i, m in enumerate(self.mimic):  # RobotState::updateMimicJoint, robot_state.h:1854-1863
            if m is not None:
                src = self.var_index[self.joint_names.index(m[0])]
                q[:, self.var_index[i]] = m[1] * q[:, src] + m[2]
        return np.ascontiguousarray(q.astype(dtype))


def panda():
    """Franka Emika Panda with hand."""
    hp = math.pi / 2.0
    r = RobotDescription("panda")
    r.add_link("panda_link0", None, "virtual_joint", JOINT_FIXED)
    r.add_link("panda_link1", "panda_link0", "panda_joint1", JOINT_REVOLUTE, (0, 0, 0.333), (0, 0, 0),
               limits=(-2.8973, 2.8973))
    r.add_link("panda_link2", "panda_link1", "panda_joint2", JOINT_REVOLUTE, (0, 0, 0), (-hp, 0, 0),
               limits=(-1.7628, 1.7628))
    r.add_link("panda_link3", "panda_link2", "panda_joint3", JOINT_REVOLUTE, (0, -0.316, 0), (hp, 0, 0),
               limits=(-2.8973, 2.8973))
    r.add_link("panda_link4", "panda_link3", "panda_joint4", JOINT_REVOLUTE, (0.0825, 0, 0), (hp, 0, 0),
               limits=(-3.0718, -0.0698))
    r.add_link("panda_link5", "panda_link4", "panda_joint5", JOINT_REVOLUTE, (-0.0825, 0.384, 0), (-hp, 0, 0),
               limits=(-2.8973, 2.8973))
    r.add_link("panda_link6", "panda_link5", "panda_joint6", JOINT_REVOLUTE, (0, 0, 0), (hp, 0, 0),
               limits=(-0.0175, 3.7525))
    r.add_link("panda_link7", "panda_link6", "panda_joint7", JOINT_REVOLUTE, (0.088, 0, 0), (hp, 0, 0),
               limits=(-2.8973, 2.8973))
    r.add_link("panda_link8", "panda_link7", "panda_joint8", JOINT_FIXED, (0, 0, 0.107))
    r.add_link("panda_hand", "panda_link8", "panda_hand_joint", JOINT_FIXED, (0, 0, 0), (0, 0, -math.pi / 4.0))
    r.add_link("panda_leftfinger", "panda_hand", "panda_finger_joint1", JOINT_PRISMATIC, (0, 0, 0.0584),
               axis=(0, 1, 0), limits=(0.0, 0.04))
    r.add_link("panda_rightfinger", "panda_hand", "panda_finger_joint2", JOINT_PRISMATIC, (0, 0, 0.0584),
               axis=(0, -1, 0), limits=(0.0, 0.04), mimic=("panda_finger_joint1", 1.0, 0.0))

    # primitive collision shapes (link frame): used for link points + bounding spheres
    r.add_shape("panda_link0", SHAPE_CYLINDER, (0.10, 0.14), (-0.04, 0.0, 0.07))
    r.add_shape("panda_link1", SHAPE_CYLINDER, (0.06, 0.26), (0.0, -0.03, -0.08))
    r.add_shape("panda_link2", SHAPE_CYLINDER, (0.06, 0.26), (0.0, -0.08, 0.03), (hp, 0, 0))
    r.add_shape("panda_link3", SHAPE_CYLINDER, (0.06, 0.22), (0.04, 0.02, -0.05))
    r.add_shape("panda_link4", SHAPE_CYLINDER, (0.06, 0.22), (-0.04, 0.05, 0.02), (hp, 0, 0))
    r.add_shape("panda_link5", SHAPE_CYLINDER, (0.06, 0.36), (0.0, 0.04, -0.11))
    r.add_shape("panda_link6", SHAPE_CYLINDER, (0.06, 0.20), (0.04, 0.01, 0.0), (0, hp, 0))
    r.add_shape("panda_link7", SHAPE_CYLINDER, (0.05, 0.14), (0.0, 0.0, 0.05))
    r.add_shape("panda_hand", SHAPE_BOX, (0.05, 0.20, 0.09), (0.0, 0.0, 0.035))
    r.add_shape("panda_leftfinger", SHAPE_BOX, (0.02, 0.02, 0.05), (0.0, 0.008, 0.03))
    r.add_shape("panda_rightfinger", SHAPE_BOX, (0.02, 0.02, 0.05), (0.0, -0.008, 0.03))

    # sphere decomposition (link_body_decompositions override)
    r.set_spheres("panda_link0", [(0.0, 0.0, 0.05, 0.10), (-0.09, 0.0, 0.06, 0.08)])
    r.set_spheres("panda_link1", [(0.0, -0.08, 0.0, 0.06), (0.0, -0.03, 0.0, 0.06), (0.0, 0.0, -0.12, 0.06),
                                  (0.0, 0.0, -0.17, 0.06)])
    r.set_spheres("panda_link2", [(0.0, 0.0, 0.03, 0.06), (0.0, 0.0, 0.08, 0.06), (0.0, -0.12, 0.0, 0.06),
                                  (0.0, -0.17, 0.0, 0.06)])
    r.set_spheres("panda_link3", [(0.0, 0.0, -0.06, 0.05), (0.0, 0.0, -0.10, 0.06), (0.08, 0.06, 0.0, 0.055),
                                  (0.08, 0.02, 0.0, 0.055)])
    r.set_spheres("panda_link4", [(0.0, 0.0, 0.02, 0.055), (0.0, 0.0, 0.06, 0.055), (-0.08, 0.095, 0.0, 0.06),
                                  (-0.08, 0.06, 0.0, 0.055)])
    r.set_spheres("panda_link5", [(0.0, 0.055, 0.0, 0.06), (0.0, 0.075, 0.0, 0.06), (0.0, 0.0, -0.22, 0.06),
                                  (0.0, 0.05, -0.18, 0.05), (0.01, 0.08, -0.14, 0.025), (0.01, 0.085, -0.11, 0.025),
                                  (0.01, 0.09, -0.08, 0.025), (0.01, 0.095, -0.05, 0.025),
                                  (-0.01, 0.08, -0.14, 0.025), (-0.01, 0.085, -0.11, 0.025),
                                  (-0.01, 0.09, -0.08, 0.025), (-0.01, 0.095, -0.05, 0.025)])
    r.set_spheres("panda_link6", [(0.0, 0.0, 0.0, 0.06), (0.08, 0.03, 0.0, 0.06), (0.08, -0.01, 0.0, 0.06)])
    r.set_spheres("panda_link7", [(0.0, 0.0, 0.07, 0.05), (0.02, 0.04, 0.08, 0.025), (0.04, 0.02, 0.08, 0.025),
                                  (0.04, 0.06, 0.085, 0.02), (0.06, 0.04, 0.085, 0.02)])
    r.set_spheres("panda_hand", [(0.0, -0.075, 0.01, 0.028), (0.0, -0.045, 0.01, 0.028), (0.0, -0.015, 0.01, 0.028),
                                 (0.0, 0.015, 0.01, 0.028), (0.0, 0.045, 0.01, 0.028), (0.0, 0.075, 0.01, 0.028),
                                 (0.0, -0.075, 0.03, 0.026), (0.0, -0.045, 0.03, 0.026), (0.0, -0.015, 0.03, 0.026),
                                 (0.0, 0.015, 0.03, 0.026), (0.0, 0.045, 0.03, 0.026), (0.0, 0.075, 0.03, 0.026),
                                 (0.0, -0.075, 0.05, 0.024), (0.0, -0.045, 0.05, 0.024), (0.0, -0.015, 0.05, 0.024),
                                 (0.0, 0.015, 0.05, 0.024), (0.0, 0.045, 0.05, 0.024), (0.0, 0.075, 0.05, 0.024)])
    r.set_spheres("panda_leftfinger", [(0.0, 0.01, 0.043, 0.011), (0.0, 0.02, 0.015, 0.011)])
    r.set_spheres("panda_rightfinger", [(0.0, -0.01, 0.043, 0.011), (0.0, -0.02, 0.015, 0.011)])

    # SRDF disable_collisions pairs (panda_moveit_config style: adjacent links and never-colliding pairs)
    P = "panda_"
    dis = [("link0", "link1"), ("link0", "link2"), ("link0", "link3"), ("link0", "link4"), ("link1", "link2"),
           ("link1", "link3"), ("link1", "link4"), ("link2", "link3"), ("link2", "link4"), ("link2", "link6"),
           ("link3", "link4"), ("link3", "link5"), ("link3", "link6"), ("link3", "link7"), ("link4", "link5"),
           ("link4", "link6"), ("link4", "link7"), ("link4", "link8"), ("link5", "link6"), ("link5", "link7"),
           ("link6", "link7"), ("link6", "link8"), ("link7", "link8"), ("link3", "hand"), ("link4", "hand"),
           ("link5", "hand"), ("link6", "hand"), ("link7", "hand"), ("link8", "hand"), ("hand", "leftfinger"),
           ("hand", "rightfinger"), ("leftfinger", "rightfinger"), ("link3", "leftfinger"), ("link4", "leftfinger"),
           ("link6", "leftfinger"), ("link7", "leftfinger"), ("link8", "leftfinger"), ("link3", "rightfinger"),
           ("link4", "rightfinger"), ("link6", "rightfinger"), ("link7", "rightfinger"), ("link8", "rightfinger")]
    r.disabled_pairs = [(P + a, P + b) for a, b in dis]
    r.groups = {
        "panda_arm": ["panda_joint%d" % i for i in range(1, 8)],
        "hand": ["panda_finger_joint1", "panda_finger_joint2"],
        "panda_arm_hand": ["panda_joint%d" % i for i in range(1, 8)] + ["panda_finger_joint1", "panda_finger_joint2"],
    }
    return r


def one_robot():
    """The inline URDF of robot_state_test.cpp:237-421 (OneRobot fixture), kinematics only."""
    r = RobotDescription("one_robot")
    r.add_link("base_link", None, "base_joint", JOINT_PLANAR)
    r.add_link("link_a", "base_link", "joint_a", JOINT_REVOLUTE, (0, 0, 0), (0, 0, 0), axis=(0, 0, 1))
    r.add_link("link_b", "link_a", "joint_b", JOINT_FIXED, (0.0, 0.5, 0.0), (0.0, -0.42, 0.0))
    r.add_link("link_c", "link_b", "joint_c", JOINT_PRISMATIC, (0.0, -0.1, 0.0), (0.0, 0.42, 0.0), axis=(1, 0, 0),
               limits=(0.0, 0.09))
    r.add_link("link_d", "link_c", "mim_f", JOINT_PRISMATIC, (0.1, 0.1, 0.0), (0, 0, 0), axis=(1, 0, 0),
               limits=(0.0, 0.19), mimic=("joint_f", 1.5, 0.1))
    r.add_link("link_e", "link_d", "joint_f", JOINT_PRISMATIC, (0.1, 0.1, 0.0), (0, 0, 0), axis=(1, 0, 0),
               limits=(0.0, 0.19))
    r.add_link("link/with/slash", "base_link", "joint_link_with_slash", JOINT_FIXED)
    return r


def mobile_arm():
    """Small synthetic mobile manipulator that exercises every joint code path of the kernels at once: planar base,
    revolute joints about x, y, -z and an oblique axis, a prismatic lift, a revolute mimic joint (x1.5 + 0.1), a
    branch (two grippers) and non-trivial rpy origins.  Fits the constant-bank tables of the FAST kernel."""
    hp = math.pi / 2.0
    r = RobotDescription("mobile_arm")
    r.add_link("base", None, "world_joint", JOINT_PLANAR, limits=[(-0.4, 0.4), (-0.4, 0.4), (-math.pi, math.pi)])
    r.add_shape("base", SHAPE_BOX, (0.5, 0.4, 0.2), (0, 0, 0.1))
    r.add_link("lift", "base", "lift_joint", JOINT_PRISMATIC, (0.1, 0, 0.2), axis=(0, 0, 1), limits=(0.0, 0.4))
    r.add_shape("lift", SHAPE_CYLINDER, (0.07, 0.3), (0, 0, 0.15))
    r.add_link("shoulder", "lift", "shoulder_joint", JOINT_REVOLUTE, (0, 0, 0.32), (0.1, -0.2, 0.3), axis=(0, 0, -1),
               limits=(-2.5, 2.5))
    r.add_shape("shoulder", SHAPE_CYLINDER, (0.06, 0.16), (0, 0, 0.06))
    r.add_link("upper", "shoulder", "upper_joint", JOINT_REVOLUTE, (0, 0, 0.12), (0, 0, 0), axis=(0, 1, 0),
               limits=(-1.8, 1.8))
    r.add_shape("upper", SHAPE_CYLINDER, (0.05, 0.3), (0.15, 0, 0), (0, hp, 0))
    r.add_link("fore", "upper", "fore_joint", JOINT_REVOLUTE, (0.3, 0, 0), (hp, 0, 0), axis=(1, 0, 0),
               limits=(-3.0, 3.0))
    r.add_shape("fore", SHAPE_BOX, (0.25, 0.08, 0.08), (0.12, 0, 0))
    r.add_link("wrist", "fore", "wrist_joint", JOINT_REVOLUTE, (0.25, 0, 0), (0, 0.3, -0.4), axis=(0.3, 0.5, 0.8),
               limits=(-2.0, 2.0))
    r.add_shape("wrist", SHAPE_SPHERE, (0.05,), (0.03, 0, 0))
    r.add_link("palm", "wrist", "palm_joint", JOINT_FIXED, (0.06, 0, 0), (0, 0, 0.5))
    r.add_shape("palm", SHAPE_BOX, (0.04, 0.12, 0.04))
    r.add_link("finger_l", "palm", "finger_l_joint", JOINT_REVOLUTE, (0.03, 0.05, 0), axis=(0, 0, 1), limits=(-0.5, 0.5))
    r.add_shape("finger_l", SHAPE_BOX, (0.08, 0.015, 0.03), (0.04, 0, 0))
    r.add_link("finger_r", "palm", "finger_r_joint", JOINT_REVOLUTE, (0.03, -0.05, 0), axis=(0, 0, 1),
               limits=(-0.65, 0.85), mimic=("finger_l_joint", 1.5, 0.1))
    r.add_shape("finger_r", SHAPE_BOX, (0.08, 0.015, 0.03), (0.04, 0, 0))
    r.set_spheres("base", [(-0.15, -0.1, 0.1, 0.14), (0.15, -0.1, 0.1, 0.14), (-0.15, 0.1, 0.1, 0.14),
                           (0.15, 0.1, 0.1, 0.14)])
    r.set_spheres("lift", [(0, 0, 0.05, 0.075), (0, 0, 0.15, 0.075), (0, 0, 0.25, 0.075)])
    r.set_spheres("shoulder", [(0, 0, 0.03, 0.065), (0, 0, 0.1, 0.065)])
    r.set_spheres("upper", [(0.05, 0, 0, 0.055), (0.13, 0, 0, 0.055), (0.21, 0, 0, 0.055), (0.28, 0, 0, 0.055)])
    r.set_spheres("fore", [(0.04, 0, 0, 0.05), (0.12, 0, 0, 0.05), (0.2, 0, 0, 0.05)])
    r.set_spheres("wrist", [(0.03, 0, 0, 0.05)])
    r.set_spheres("palm", [(0, -0.04, 0, 0.025), (0, 0, 0, 0.025), (0, 0.04, 0, 0.025)])
    r.set_spheres("finger_l", [(0.02, 0, 0, 0.015), (0.05, 0, 0, 0.015), (0.075, 0, 0, 0.012)])
    r.set_spheres("finger_r", [(0.02, 0, 0, 0.015), (0.05, 0, 0, 0.015), (0.075, 0, 0, 0.012)])
    P = [("base", "lift"), ("lift", "shoulder"), ("shoulder", "upper"), ("upper", "fore"), ("fore", "wrist"),
         ("wrist", "palm"), ("palm", "finger_l"), ("palm", "finger_r"), ("fore", "palm"), ("base", "shoulder"),
         ("wrist", "finger_l"), ("wrist", "finger_r")]
    r.disabled_pairs = list(P)
    r.groups = {"arm": ["shoulder_joint", "upper_joint", "fore_joint", "wrist_joint", "finger_l_joint",
                        "finger_r_joint"]}
    return r


def pr2_like():
    """PR2-like dual-arm tree (the real PR2 URDF lives in moveit_resources, F3):
    planar base, prismatic torso, head pan/tilt, two 7-DoF arms with 4-finger-link grippers driven by mimic joints."""
    hp = math.pi / 2.0
    r = RobotDescription("pr2_like")
    r.add_link("base_footprint", None, "world_joint", JOINT_PLANAR,
               limits=[(-0.5, 0.5), (-0.5, 0.5), (-math.pi, math.pi)])
    r.add_link("base_link", "base_footprint", "base_footprint_joint", JOINT_FIXED, (0, 0, 0.051))
    r.add_shape("base_link", SHAPE_BOX, (0.65, 0.65, 0.25), (0, 0, 0.125))
    r.add_link("torso_lift_link", "base_link", "torso_lift_joint", JOINT_PRISMATIC, (-0.05, 0, 0.74), axis=(0, 0, 1),
               limits=(0.0, 0.31))
    r.add_shape("torso_lift_link", SHAPE_BOX, (0.3, 0.4, 0.7), (-0.1, 0, -0.3))
    r.add_link("head_pan_link", "torso_lift_link", "head_pan_joint", JOINT_REVOLUTE, (-0.017, 0, 0.381),
               limits=(-2.8, 2.8))
    r.add_shape("head_pan_link", SHAPE_CYLINDER, (0.08, 0.1), (0, 0, 0.03))
    r.add_link("head_tilt_link", "head_pan_link", "head_tilt_joint", JOINT_REVOLUTE, (0.068, 0, 0), axis=(0, 1, 0),
               limits=(-0.37, 1.29))
    r.add_shape("head_tilt_link", SHAPE_BOX, (0.2, 0.3, 0.15), (0.03, 0, 0.08))
    for side, sy in (("r", -1.0), ("l", 1.0)):
        p = side + "_"
        r.add_link(p + "shoulder_pan_link", "torso_lift_link", p + "shoulder_pan_joint", JOINT_REVOLUTE,
                   (0.0, sy * 0.188, 0.0), limits=(-2.1, 0.56) if side == "r" else (-0.56, 2.1))
        r.add_shape(p + "shoulder_pan_link", SHAPE_CYLINDER, (0.12, 0.3), (0.0, 0.0, -0.12))
        r.add_link(p + "shoulder_lift_link", p + "shoulder_pan_link", p + "shoulder_lift_joint", JOINT_REVOLUTE,
                   (0.1, 0, 0), axis=(0, 1, 0), limits=(-0.35, 1.29))
        r.add_shape(p + "shoulder_lift_link", SHAPE_CYLINDER, (0.08, 0.16), (0, 0, 0), (hp, 0, 0))
        r.add_link(p + "upper_arm_roll_link", p + "shoulder_lift_link", p + "upper_arm_roll_joint", JOINT_REVOLUTE,
                   (0, 0, 0), axis=(1, 0, 0), limits=(-3.75, 0.65) if side == "r" else (-0.65, 3.75))
        r.add_link(p + "upper_arm_link", p + "upper_arm_roll_link", p + "upper_arm_joint", JOINT_FIXED)
        r.add_shape(p + "upper_arm_link", SHAPE_CYLINDER, (0.07, 0.36), (0.21, 0, 0), (0, hp, 0))
        r.add_link(p + "elbow_flex_link", p + "upper_arm_link", p + "elbow_flex_joint", JOINT_REVOLUTE, (0.4, 0, 0),
                   axis=(0, 1, 0), limits=(-2.12, -0.15))
        r.add_shape(p + "elbow_flex_link", SHAPE_CYLINDER, (0.06, 0.14), (0, 0, 0), (hp, 0, 0))
        r.add_link(p + "forearm_roll_link", p + "elbow_flex_link", p + "forearm_roll_joint", JOINT_REVOLUTE,
                   (0, 0, 0), axis=(1, 0, 0), limits=(-math.pi, math.pi))
        r.add_link(p + "forearm_link", p + "forearm_roll_link", p + "forearm_joint", JOINT_FIXED)
        r.add_shape(p + "forearm_link", SHAPE_CYLINDER, (0.055, 0.3), (0.17, 0, 0), (0, hp, 0))
        r.add_link(p + "wrist_flex_link", p + "forearm_link", p + "wrist_flex_joint", JOINT_REVOLUTE, (0.321, 0, 0),
                   axis=(0, 1, 0), limits=(-2.0, -0.1))
        r.add_shape(p + "wrist_flex_link", SHAPE_CYLINDER, (0.045, 0.1), (0, 0, 0), (hp, 0, 0))
        r.add_link(p + "wrist_roll_link", p + "wrist_flex_link", p + "wrist_roll_joint", JOINT_REVOLUTE, (0, 0, 0),
                   axis=(1, 0, 0), limits=(-math.pi, math.pi))
        r.add_link(p + "gripper_palm_link", p + "wrist_roll_link", p + "gripper_palm_joint", JOINT_FIXED)
        r.add_shape(p + "gripper_palm_link", SHAPE_BOX, (0.1, 0.1, 0.06), (0.06, 0, 0))
        r.add_link(p + "gripper_l_finger_link", p + "gripper_palm_link", p + "gripper_l_finger_joint", JOINT_REVOLUTE,
                   (0.07691, 0.01, 0), limits=(0.0, 0.548))
        r.add_shape(p + "gripper_l_finger_link", SHAPE_BOX, (0.1, 0.025, 0.03), (0.045, 0.012, 0))
        r.add_link(p + "gripper_l_finger_tip_link", p + "gripper_l_finger_link", p + "gripper_l_finger_tip_joint",
                   JOINT_REVOLUTE, (0.09137, 0.00495, 0), axis=(0, 0, -1), limits=(0.0, 0.548),
                   mimic=(p + "gripper_l_finger_joint", 1.0, 0.0))
        r.add_shape(p + "gripper_l_finger_tip_link", SHAPE_BOX, (0.04, 0.015, 0.025), (0.02, -0.005, 0))
        r.add_link(p + "gripper_r_finger_link", p + "gripper_palm_link", p + "gripper_r_finger_joint", JOINT_REVOLUTE,
                   (0.07691, -0.01, 0), axis=(0, 0, -1), limits=(0.0, 0.548),
                   mimic=(p + "gripper_l_finger_joint", 1.0, 0.0))
        r.add_shape(p + "gripper_r_finger_link", SHAPE_BOX, (0.1, 0.025, 0.03), (0.045, -0.012, 0))
        r.add_link(p + "gripper_r_finger_tip_link", p + "gripper_r_finger_link", p + "gripper_r_finger_tip_joint",
                   JOINT_REVOLUTE, (0.09137, -0.00495, 0), limits=(0.0, 0.548),
                   mimic=(p + "gripper_l_finger_joint", 1.0, 0.0))
        r.add_shape(p + "gripper_r_finger_tip_link", SHAPE_BOX, (0.04, 0.015, 0.025), (0.02, 0.005, 0))
    # ACM = all pairs enabled except parent-child (test_collision_distance_field.cpp:65,196 style)
    r.disabled_pairs = []
    for i in range(r.n_links):
        # nearest ancestor with geometry counts as "adjacent" too (roll links carry no geometry)
        a = r.parent[i]
        while a >= 0:
            r.disabled_pairs.append((r.link_names[a], r.link_names[i]))
            if r.shapes[a]:
                break
            a = r.parent[a]
    # ... and the pairs that collide in every random state (grandparent / grandchild cylinders overlap by construction):
    # what the Setup Assistant's "always in collision" rule finds (moveit_setup_assistant/src/tools/
    # compute_default_collisions.cpp:474-547, >= 95 % of the samples; tools/pr2_default_collisions.py reproduces the
    # list with the oracle's sphere decomposition) and the real PR2's SRDF disables likewise
    for side in ("r", "l"):
        p = side + "_"
        r.disabled_pairs += [("torso_lift_link", p + "shoulder_lift_link"), (p + "shoulder_pan_link", p + "upper_arm_link"),
                             (p + "upper_arm_link", p + "forearm_link"), (p + "forearm_link", p + "gripper_palm_link")]
    r.groups = {
        "right_arm": ["r_shoulder_pan_joint", "r_shoulder_lift_joint", "r_upper_arm_roll_joint", "r_elbow_flex_joint",
                      "r_forearm_roll_joint", "r_wrist_flex_joint", "r_wrist_roll_joint"],
        "left_arm": ["l_shoulder_pan_joint", "l_shoulder_lift_joint", "l_upper_arm_roll_joint", "l_elbow_flex_joint",
                     "l_forearm_roll_joint", "l_wrist_flex_joint", "l_wrist_roll_joint"],
    }
    r.groups["arms"] = r.groups["right_arm"] + r.groups["left_arm"]
    return r
