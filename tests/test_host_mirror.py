"""The C++ host mirror (moveit_b200/host/collision_env_b200.hpp: CollisionEnvB200 + CollisionDetectorAllocatorB200,
the reference-shaped plugin body) driven end to end by host_driver and compared with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from moveit_b200 import robot_description as rd
from moveit_b200 import scenario_io, scenarios

from common import OracleSetup, oracle_world_points, self_field_points

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_driver():
    import runpy
    mod = runpy.run_path(os.path.join(ROOT, "moveit_b200", "host", "build_host.py"))
    return mod["build"]()


def test_host_driver_compiles_against_the_c_abi():
    """g++ -std=c++17 -Wall -Wextra over the header-only mirror; links libb200df.so only (no torch, no oracle)."""
    exe = _build_driver()
    assert os.access(exe, os.X_OK)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libb200df.so" in out and "liboracle" not in out and "torch" not in out


def test_scenario_writer_round_trips_doubles(tmp_path):
    desc = rd.panda()
    p = tmp_path / "s.txt"
    scenario_io.write_scenario(str(p), desc, scenarios.FIELD_A, scenarios.random_scene())
    toks = open(p).read().split()
    assert toks[0] == "robot" and int(toks[2]) == desc.n_links
    i = toks.index("obj0")
    assert float(toks[i + 2]) == scenarios.random_scene()[0][1][0]


def _run(tmp_path, desc, field, objs, q, **kw):
    exe = _build_driver()
    scen, states, prefix = str(tmp_path / "scen.txt"), str(tmp_path / "q.f64"), str(tmp_path / "out")
    scenario_io.write_scenario(scen, desc, field, objs, **kw)
    np.ascontiguousarray(q, dtype=np.float64).tofile(states)
    r = subprocess.run([exe, scen, states, prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "B200DistanceField" in r.stdout
    print(r.stdout.strip())
    return prefix


@pytest.mark.gpu
def test_panda_arm_group_through_the_cpp_env(gpu, oracle, tmp_path):
    """Planning group panda_arm with the SRDF ACM: batched and single-state verdicts, contacts, gradients and a
    world mutation through the observer, all against the oracle."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=41)
    fa = scenarios.FIELD_A
    n = 4000
    q = desc.sample_states(n, 2024)
    q0 = np.zeros(desc.nvar)
    prefix = _run(tmp_path, desc, fa, objs, q, group="panda_arm", n_single=64, n_contacts=200, max_contacts=10,
                  max_contacts_per_pair=2, n_grad=50, mutate=True)
    o = OracleSetup(oracle, desc, fa, objs, group="panda_arm", self_state=q0)
    robot = np.fromfile(prefix + ".robot.u8", dtype=np.uint8)
    self_ = np.fromfile(prefix + ".self.u8", dtype=np.uint8)
    both = np.fromfile(prefix + ".all.u8", dtype=np.uint8)
    assert np.array_equal(robot, o.env.check_batch(q, mode=1, n_threads=4))
    assert np.array_equal(self_, o.env.check_batch(q, mode=2, n_threads=4))
    assert np.array_equal(both, o.env.check_batch(q, mode=4, n_threads=4))
    assert 0 < robot.sum() < n and self_.sum() > 0  # link1 always touches link0's voxels in the self field
    single = np.fromfile(prefix + ".single.u8", dtype=np.uint8)
    assert np.array_equal(single & 1, robot[:64]) and np.array_equal((single >> 1) & 1, self_[:64])
    assert np.array_equal((single >> 2) & 1, both[:64])

    # contacts: same pairs, same order inside a pair, same positions
    names = [desc.link_names[i] for i in o.group_links]
    got, cur = {}, None
    for line in open(prefix + ".contacts.txt"):
        t = line.split()
        if t[0] == "state":
            cur = int(t[1])
            got[cur] = dict(collision=int(t[2]), count=int(t[3]), contacts=[])
        else:
            got[cur]["contacts"].append((t[1], t[2], np.array([float(t[3]), float(t[4]), float(t[5])])))
    n_with = 0
    for i in range(200):
        col, recs = o.env.check_contacts(q[i], 10, 2)
        want = {}
        for (l1, b2, _, pos) in recs:
            key = (names[l1], "self" if b2 == -1 else "environment" if b2 == -2 else names[b2])
            want.setdefault(key, []).append(pos)
        assert got[i]["collision"] == int(col) and got[i]["count"] == len(recs)
        flat = [(k[0], k[1], p) for k in sorted(want) for p in want[k]]  # std::map iteration order
        assert len(flat) == len(got[i]["contacts"])
        for (a, b, p), (ga, gb, gp) in zip(flat, got[i]["contacts"]):
            assert (a, b) == (ga, gb) and np.max(np.abs(p - gp)) < 1e-12  # CUDA vs glibc sincos: <= 2 ulp
        n_with += bool(recs)
    assert n_with > 20

    # gradients: per-link PosedDistanceFields + self field + intra + environment, like getCollisionGradients
    g = np.fromfile(prefix + ".grad.f64").reshape(50, -1, 8)
    o.env.build_link_fields()
    ref = o.env.gradients_batch(q[:50], flags=7)
    assert g.shape[1] == ref["dist"].shape[1]
    assert np.max(np.abs(g[:, :, 5:8] - ref["centers"])) < 1e-12
    assert np.array_equal(g[:, :, 4].astype(np.int32), ref["type"])
    fin = ref["dist"] < 1e300
    assert np.array_equal(fin, g[:, :, 0] < 1e300)
    assert np.max(np.abs(g[:, :, 0][fin] - ref["dist"][fin])) < 1e-5
    assert np.max(np.abs(g[:, :, 1:4] - ref["grad"])) < 1e-5

    # batched motion validation (8 segments between consecutive states) and isPathValid's invalid_index
    first = np.fromfile(prefix + ".motion.i32", dtype=np.int32)
    qa, qb = q[0::2], q[1::2]
    K = 8
    inter = np.stack([qa + (qb - qa) * (float(j) / float(K)) for j in range(1, K + 1)], axis=1)
    ref = o.env.check_batch(inter.reshape(-1, desc.nvar), mode=4, n_threads=8).reshape(len(qa), K)
    assert np.array_equal(first, np.where(ref.any(axis=1), ref.argmax(axis=1) + 1, 0))
    path = np.fromfile(prefix + ".path.i32", dtype=np.int32)
    assert np.array_equal(path[:-1], np.nonzero(both[:500])[0]) and path[-1] == int(not both[:500].any())

    # world mutation: object 0 moved by +0.1 m in x, object 1 destroyed (removePointsFromField + addPointsToField)
    pts = [oracle_world_points(oracle, [ob], fa["resolution"]) for ob in objs]
    moved = (objs[0][0], objs[0][1], np.array(objs[0][2]) + np.array([[0, 0, 0, 0.1]] * 3) * np.array([[1], [0], [0]]))
    o.env.world.remove_points(pts[0])
    o.env.world.add_points(oracle_world_points(oracle, [moved], fa["resolution"]))
    o.env.world.remove_points(pts[1])
    o.env.world.make_exact()
    robot2 = np.fromfile(prefix + ".robot2.u8", dtype=np.uint8)
    assert np.array_equal(robot2, o.env.check_batch(q, mode=1, n_threads=4))
    assert not np.array_equal(robot2, robot)


@pytest.mark.gpu
def test_whole_robot_no_acm_and_derived_spheres(gpu, oracle, tmp_path):
    """group "" (all links), null ACM branch, spheres from determineCollisionSpheres, collision tolerance."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=43)
    q = desc.sample_states(1500, 5)
    prefix = _run(tmp_path, desc, scenarios.FIELD_A, objs, q, group="", use_acm=False, sphere_overrides=False,
                  tol=0.01, n_single=8)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, use_acm=False, sphere_overrides=False, tol=0.01,
                    self_state=np.zeros(desc.nvar))
    assert np.array_equal(np.fromfile(prefix + ".robot.u8", dtype=np.uint8), o.env.check_batch(q, mode=1, n_threads=4))
    assert np.array_equal(np.fromfile(prefix + ".all.u8", dtype=np.uint8), o.env.check_batch(q, mode=4, n_threads=4))


@pytest.mark.gpu
def test_attached_body_through_the_cpp_env(gpu, oracle, tmp_path):
    """A box attached to panda_hand (RobotState::attachBody).  In the reference an attached shape is posed by
    link_transform * attach_trans and decomposed with the default 0.01 padding, i.e. it behaves like one more fixed
    link; the ACM entry and the touch links of a body are only honoured for the link it is attached to
    (generateDistanceFieldCacheEntry :806-834).  The oracle is given that equivalent model."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=41)
    fa = scenarios.FIELD_A
    q = desc.sample_states(3000, 777)
    pose = rd.rpy_xyz_to_pose((0.0, 0.0, 0.22), (0.1, 0.2, 0.3))
    box = (rd.SHAPE_BOX, (0.06, 0.10, 0.08), rd.rpy_xyz_to_pose())
    attached = [("grasped_box", "panda_hand", [(box[0], box[1], pose)], ["panda_hand", "panda_leftfinger"])]
    prefix = _run(tmp_path, desc, fa, objs, q, group="", n_single=32, n_contacts=100, max_contacts=20,
                  max_contacts_per_pair=3, attached=attached, extra_acm_always=[("grasped_box", "panda_link7")])
    # the equivalent model for the oracle: the box as a fixed child link of panda_hand
    aug = rd.panda()
    aug.add_link("grasped_box", "panda_hand", "grasped_box/attach0", rd.JOINT_FIXED, origin_pose=pose)
    aug.add_shape("grasped_box", box[0], box[1])
    L = aug.n_links
    acm = np.zeros((L, L), dtype=np.int32)
    acm[:L - 1, :L - 1] = desc.acm_entry_matrix()
    b, hand = L - 1, aug.link_names.index("panda_hand")
    acm[b, :] = acm[:, b] = -1          # other links: the reference never looks at their entry / touch status
    acm[b, b] = 0                       # AllowedCollisionMatrix(names, false) has no entry for the body: not ALWAYS
    acm[hand, b] = acm[b, hand] = 1     # touch link == the attach link
    o = OracleSetup(oracle, aug, fa, objs, self_state=np.zeros(desc.nvar), padding_by_link={"grasped_box": 0.01},
                    acm_override=acm, group_links=list(range(L)))
    robot = np.fromfile(prefix + ".robot.u8", dtype=np.uint8)
    self_ = np.fromfile(prefix + ".self.u8", dtype=np.uint8)
    both = np.fromfile(prefix + ".all.u8", dtype=np.uint8)
    assert np.array_equal(robot, o.env.check_batch(q, mode=1, n_threads=4))
    assert np.array_equal(self_, o.env.check_batch(q, mode=2, n_threads=4))
    assert np.array_equal(both, o.env.check_batch(q, mode=4, n_threads=4))
    # the box matters: the same states without it give different verdicts
    plain = OracleSetup(oracle, desc, fa, objs, self_state=np.zeros(desc.nvar))
    assert (plain.env.check_batch(q, mode=1, n_threads=4) != robot).any()
    assert (plain.env.check_batch(q, mode=2, n_threads=4) != self_).any()
    single = np.fromfile(prefix + ".single.u8", dtype=np.uint8)
    assert np.array_equal((single >> 2) & 1, both[:32])
    # contacts name the body and mark it ROBOT_ATTACHED (BodyTypes::ROBOT_ATTACHED == 1)
    seen = False
    for line in open(prefix + ".contacts.txt"):
        t = line.split()
        if t[0] == "contact" and t[1] == "grasped_box":
            assert int(t[6]) == 1
            seen = True
        if t[0] == "contact" and t[2] == "grasped_box":
            assert int(t[7]) == 1
            seen = True
    assert seen


@pytest.mark.gpu
def test_distance_field_class_and_stream_format(gpu, oracle, tmp_path):
    """PropagationDistanceFieldB200 (the C++ mirror of the reference's DistanceField API) runs the reference's
    TestReadWrite (test_distance_field.cpp:882-928) and its moved-shape test; the streams it writes are read by the
    Python engine and by the oracle's restatement of readFromStream, and it reads a stream the Python engine wrote."""
    from moveit_b200 import engine
    exe = _build_driver()
    prefix = str(tmp_path / "fs")
    pts = [(0.1, 0.0, 0.0), (0.0, 0.1, 0.2), (0.4, 0.0, 0.0)]
    f = engine.Field((1.0, 1.0, 1.0), 0.1, (0.0, 0.0, 0.0), 0.3)
    f.add_points(pts)
    py_stream = str(tmp_path / "py.df")
    with open(py_stream, "wb") as fo:
        assert f.write_stream(fo)
    r = subprocess.run([exe, "--field-stream", prefix, py_stream], capture_output=True, text=True)
    assert r.returncode == 0 and "field stream test ok" in r.stdout, r.stderr + r.stdout
    # C++ read the Python-written stream
    assert np.array_equal(np.fromfile(prefix + ".foreign.d2.i32", dtype=np.int32).reshape(f.dims), f.d2())
    # the C++-written small stream: same bytes up to the zlib payload, same field through every reader
    small = open(prefix + ".small.df", "rb").read()
    assert small == open(py_stream, "rb").read()
    g = engine.Field.from_stream(open(prefix + ".small.df", "rb"), 0.25)
    want = np.fromfile(prefix + ".small.d2.i32", dtype=np.int32).reshape(g.dims)
    assert np.array_equal(g.d2(), want) and np.array_equal(want, f.d2())
    o = oracle.Field.from_stream(small, 0.25)
    assert o is not None and np.array_equal(o.d2(), want)  # three isolated points: propagation is exact here
    # the big one (150 x 150 x 200, a sphere of radius 0.5): Python engine and oracle read what C++ wrote
    big = open(prefix + ".big.df", "rb").read()
    gb = engine.Field.from_stream(open(prefix + ".big.df", "rb"), 0.25)
    assert gb.dims == (150, 150, 200)
    wb = np.fromfile(prefix + ".big.d2.i32", dtype=np.int32).reshape(gb.dims)
    assert np.array_equal(gb.d2(), wb) and (wb == 0).sum() > 10000
    ob = oracle.Field.from_stream(big, 0.25)
    assert ob is not None and np.array_equal(ob.d2() == 0, wb == 0)
    ob.make_exact()
    assert np.array_equal(ob.d2(), wb)
    assert engine.Field.from_stream(open(prefix + ".small.d2.i32", "rb"), 0.25) is None  # not a field stream


@pytest.mark.gpu
@pytest.mark.parametrize("case", [1, 5])  # a planar-root tree with mimic joints, a floating-root tree
def test_random_robots_through_the_cpp_env(gpu, oracle, tmp_path, case):
    """The C++ mirror builds its own kinematic model from RobotModel-shaped input (variable indexing, mimic joints by
    name, planar / floating variable blocks, ACM by link names): the fuzz robots of test_gpu_parity.py through it."""
    from common import FUZZ_CASES, random_robot
    seed, n_links, floating, kw = FUZZ_CASES[case]
    desc = random_robot(100 + seed, n_links, floating, **kw)
    fd = scenarios.FIELD_DEFAULT
    objs = scenarios.random_scene(seed=200 + seed, n_boxes=6, n_cylinders=3)
    q = desc.sample_states(3000, 300 + seed)
    prefix = _run(tmp_path, desc, fd, objs, q, group="", n_single=16, n_grad=20)
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    robot = np.fromfile(prefix + ".robot.u8", dtype=np.uint8)
    both = np.fromfile(prefix + ".all.u8", dtype=np.uint8)
    assert np.array_equal(robot, o.env.check_batch(q, mode=1, n_threads=8))
    assert np.array_equal(np.fromfile(prefix + ".self.u8", dtype=np.uint8), o.env.check_batch(q, mode=2, n_threads=8))
    assert np.array_equal(both, o.env.check_batch(q, mode=4, n_threads=8))
    assert 0 < both.sum() < len(q)
    single = np.fromfile(prefix + ".single.u8", dtype=np.uint8)
    assert np.array_equal(single & 1, robot[:16]) and np.array_equal((single >> 2) & 1, both[:16])
