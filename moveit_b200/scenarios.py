"""Synthetic workloads of BASELINE.json's configs, generated from the portable
SplitMix64 stream so every implementation sees identical inputs.  Pure data generation: no oracle, no kernels."""
import math

import numpy as np

from .rng import uniform01
from .robot_description import SHAPE_BOX, SHAPE_CYLINDER, quat_xyz_to_pose

SEED_SCENE = 0xB2000001
SEED_STATES = 0xB2000002
SEED_CLOUD = 0xB2000003
SEED_TRAJ = 0xB2000005

# Field A: 2 m cube @ 2 cm centred at (0,0,0.5); Field B: 4 m cube @ 1 cm centred at the origin
FIELD_A = dict(size=(2.0, 2.0, 2.0), resolution=0.02, center=(0.0, 0.0, 0.5), max_distance=0.25)
FIELD_B = dict(size=(4.0, 4.0, 4.0), resolution=0.01, center=(0.0, 0.0, 0.0), max_distance=0.25)
FIELD_DEFAULT = dict(size=(3.0, 3.0, 4.0), resolution=0.02, center=(0.0, 0.0, 0.0), max_distance=0.25)


def random_scene(seed=SEED_SCENE, n_boxes=8, n_cylinders=4):
    """8 boxes + 4 cylinders, random pose (normalised 4xU(-1,1) quaternion like
    compare_collision_speed_checking_fcl_bullet.cpp:114-139).  Returns [(type, dims, pose3x4)]."""
    n = n_boxes + n_cylinders
    u = uniform01(seed, n * 10).reshape(n, 10)
    objs = []
    for k in range(n):
        r = u[k]
        xyz = (-0.8 + 1.6 * r[0], -0.8 + 1.6 * r[1], 1.0 * r[2])
        quat = [2.0 * r[3] - 1.0, 2.0 * r[4] - 1.0, 2.0 * r[5] - 1.0, 2.0 * r[6] - 1.0]
        if sum(c * c for c in quat) < 1e-12:
            quat = [0.0, 0.0, 0.0, 1.0]
        pose = quat_xyz_to_pose(xyz, quat)
        if k < n_boxes:
            dims = (0.05 + 0.15 * r[7], 0.05 + 0.15 * r[8], 0.05 + 0.15 * r[9])
            objs.append((SHAPE_BOX, dims, pose))
        else:
            dims = (0.03 + 0.07 * r[7], 0.1 + 0.4 * r[8])
            objs.append((SHAPE_CYLINDER, dims, pose))
    return objs


def field_corner(field):
    return [field["center"][i] - 0.5 * field["size"][i] for i in range(3)]


def synthetic_cloud(n_points, field=FIELD_B, seed=SEED_CLOUD, leaf=0.01):
    """Config 3: points on 6 planes / 10 boxes / 5 spheres, snapped to octomap leaf centres
    ((key - 32768 + 0.5) * leaf).  Returns float64 [n,3]."""
    half = [0.5 * s for s in field["size"]]
    c = field["center"]
    u = uniform01(seed, n_points * 4).reshape(n_points, 4)
    g = uniform01(seed ^ 0x5555, 15 * 7).reshape(15, 7)
    pts = np.empty((n_points, 3))
    kind = (u[:, 0] * 21).astype(np.int64)  # 0-5 planes, 6-15 boxes, 16-20 spheres
    a, b, w = u[:, 1], u[:, 2], u[:, 3]
    for k in range(21):
        m = kind == k
        cnt = int(m.sum())
        if cnt == 0:
            continue
        if k < 6:  # walls slightly inside the field
            axis, side = k // 2, (k % 2) * 2 - 1
            p = np.empty((cnt, 3))
            others = [d for d in range(3) if d != axis]
            p[:, axis] = c[axis] + side * 0.9 * half[axis]
            p[:, others[0]] = c[others[0]] + (2 * a[m] - 1) * 0.9 * half[others[0]]
            p[:, others[1]] = c[others[1]] + (2 * b[m] - 1) * 0.9 * half[others[1]]
        elif k < 16:  # box surfaces
            gi = g[k - 6]
            ctr = np.array([c[d] + (2 * gi[d] - 1) * 0.6 * half[d] for d in range(3)])
            ext = 0.1 + 0.3 * gi[3:6]
            face = (w[m] * 6).astype(np.int64)
            p = np.empty((cnt, 3))
            uv = np.stack([2 * a[m] - 1, 2 * b[m] - 1], axis=1)
            for f in range(6):
                fm = face == f
                axis, side = f // 2, (f % 2) * 2 - 1
                others = [d for d in range(3) if d != axis]
                p[fm, axis] = ctr[axis] + side * ext[axis]
                p[fm, others[0]] = ctr[others[0]] + uv[fm, 0] * ext[others[0]]
                p[fm, others[1]] = ctr[others[1]] + uv[fm, 1] * ext[others[1]]
        else:  # sphere surfaces
            gi = g[k - 6]
            ctr = np.array([c[d] + (2 * gi[d] - 1) * 0.6 * half[d] for d in range(3)])
            rad = 0.15 + 0.35 * gi[3]
            zz = 2 * a[m] - 1
            ph = 2 * math.pi * b[m]
            rr = np.sqrt(np.maximum(0.0, 1 - zz * zz))
            p = np.stack([ctr[0] + rad * rr * np.cos(ph), ctr[1] + rad * rr * np.sin(ph), ctr[2] + rad * zz], axis=1)
        pts[m] = p
    return (np.floor(pts / leaf) + 0.5) * leaf


def chomp_trajectories(desc, n_traj, n_waypoints, seed=SEED_TRAJ, sigma=0.05):
    """Config 5: straight-line interpolation between random start/goal + Gaussian noise (Box-Muller on the
    SplitMix64 stream), clipped to the limits.  Returns float64 [n_traj, n_waypoints, nvar]."""
    nv = desc.nvar
    q0 = desc.sample_states(n_traj, seed)
    q1 = desc.sample_states(n_traj, seed + 1)
    t = np.linspace(0.0, 1.0, n_waypoints).reshape(1, n_waypoints, 1)
    traj = q0[:, None, :] * (1 - t) + q1[:, None, :] * t
    u = uniform01(seed + 2, n_traj * n_waypoints * nv * 2).reshape(2, n_traj, n_waypoints, nv)
    noise = np.sqrt(-2.0 * np.log(1.0 - u[0])) * np.cos(2 * math.pi * u[1]) * sigma
    traj = traj + noise
    lo = np.array([l[0] for l in desc.limits])
    hi = np.array([l[1] for l in desc.limits])
    traj = np.clip(traj, lo, hi)
    for i, m in enumerate(desc.mimic):
        if m is not None:
            src = desc.var_index[desc.joint_names.index(m[0])]
            traj[..., desc.var_index[i]] = m[1] * traj[..., src] + m[2]
    return np.ascontiguousarray(traj)
