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


def test_boundary_contract_without_a_device():
    """collision_env.h gives a CollisionEnv no error channel (void methods, no exceptions).  On this CPU-only box the
    env cannot initialise: it must construct anyway, never throw, record every failure and answer FAIL-CLOSED
    (res.collision = true, batch verdicts of ones) - through the plain env, the hybrid env and the allocator."""
    exe = _build_driver()
    r = subprocess.run([exe, "--contract"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout
    assert "contract test ok" in r.stdout and "errors recorded" in r.stdout
    assert "[b200df] ERROR" in r.stderr  # the ROS_ERROR stand-in: failures are logged, not swallowed


@pytest.mark.gpu
def test_boundary_contract_with_a_device(gpu):
    """The same contract on the GPU box, where the failure is provoked by a field the engine refuses (resolution 0)."""
    exe = _build_driver()
    r = subprocess.run([exe, "--contract"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout
    assert "contract test ok" in r.stdout


@pytest.mark.gpu
def test_chomp_consumer_loop_through_the_hybrid_env(gpu, oracle, tmp_path):
    """ChompOptimizer's use of the path (chomp_optimizer.cpp:81-106, 895-965) on CollisionEnvHybridB200: one
    getCollisionGradients with the scene ACM creates the GroupStateRepresentation, then 100 waypoints x 6 trajectories
    reuse it with acm = nullptr (the gsr's cache entry keeps the ACM and the per-link distance fields); the same
    waypoints through ONE batched call must give identical numbers (the driver checks that), and everything is compared
    with the oracle's getCollisionGradients."""
    exe = _build_driver()
    desc = rd.panda()
    objs = scenarios.random_scene(seed=41)
    fa = scenarios.FIELD_A
    n_traj, n_wp = 6, 100
    q = scenarios.chomp_trajectories(desc, n_traj, n_wp).reshape(-1, desc.nvar)
    scen, states, prefix = str(tmp_path / "scen.txt"), str(tmp_path / "q.f64"), str(tmp_path / "out")
    scenario_io.write_scenario(scen, desc, fa, objs, group="panda_arm")
    np.ascontiguousarray(q, dtype=np.float64).tofile(states)
    r = subprocess.run([exe, "--chomp", scen, states, prefix, str(n_wp)], capture_output=True, text=True)
    assert r.returncode == 0 and "chomp test ok" in r.stdout, r.stderr[-2000:] + r.stdout
    assert "link-field term 0" in r.stdout  # LINK_FIELDS_APPLIED
    print(r.stdout.strip())
    o = OracleSetup(oracle, desc, fa, objs, group="panda_arm", self_state=np.zeros(desc.nvar))
    o.env.build_link_fields()
    ref = o.env.gradients_batch(q, flags=7)
    g = np.fromfile(prefix + ".chomp.f64").reshape(len(q), -1, 8)
    assert g.shape[1] == ref["dist"].shape[1]
    assert np.array_equal(g[:, :, 4].astype(np.int32), ref["type"])
    assert np.max(np.abs(g[:, :, 5:8] - ref["centers"])) < 1e-12
    fin = ref["dist"] < 1e300
    assert np.array_equal(fin, g[:, :, 0] < 1e300)
    assert np.max(np.abs(g[:, :, 0][fin] - ref["dist"][fin])) <= 1e-5
    assert np.max(np.abs(g[:, :, 1:4] - ref["grad"])) <= 1e-5
    # CHOMP's per-point collision flag (distance - radius < radius) from the oracle's numbers
    radii = np.concatenate([np.reshape(o.model_desc["spheres"][li], (-1, 4))[:, 3] for li in o.group_links
                            if o.model_desc["has_geometry"][li]])
    assert radii.shape == (g.shape[1],)
    flag = np.fromfile(prefix + ".chomp_collision.u8", dtype=np.uint8)
    want = (ref["dist"] - radii[None, :] < radii[None, :]).any(axis=1)  # DBL_MAX distances never satisfy it
    assert flag.shape == (len(q),) and 0 < flag.sum() <= len(q)
    assert np.array_equal(flag.astype(bool), want)


@pytest.mark.gpu
def test_attached_body_without_an_acm_takes_part_in_no_test(gpu, oracle, tmp_path):
    """generateDistanceFieldCacheEntry only lists attached bodies inside `if (acm)` (.cpp:806-834): without an ACM a
    grasped object is tested against nothing - not the self field, not the other links, not the environment.  The
    verdicts must be those of the bare robot (round-1 advisor finding: the mirror used to report the box colliding
    with the hand it is attached to)."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=41)
    fa = scenarios.FIELD_A
    q = desc.sample_states(2000, 778)
    pose = rd.rpy_xyz_to_pose((0.0, 0.0, 0.22), (0.1, 0.2, 0.3))
    attached = [("grasped_box", "panda_hand", [(rd.SHAPE_BOX, (0.06, 0.10, 0.08), pose)], ["panda_hand"])]
    prefix = _run(tmp_path, desc, fa, objs, q, group="", use_acm=False, n_single=16, attached=attached)
    plain = OracleSetup(oracle, desc, fa, objs, use_acm=False, self_state=np.zeros(desc.nvar))
    robot = np.fromfile(prefix + ".robot.u8", dtype=np.uint8)
    both = np.fromfile(prefix + ".all.u8", dtype=np.uint8)
    assert np.array_equal(robot, plain.env.check_batch(q, mode=1, n_threads=4))
    assert np.array_equal(np.fromfile(prefix + ".self.u8", dtype=np.uint8), plain.env.check_batch(q, mode=2, n_threads=4))
    assert np.array_equal(both, plain.env.check_batch(q, mode=4, n_threads=4))
    assert 0 < robot.sum() < len(q)  # (without an ACM adjacent links always touch: `both` is all ones either way)
    single = np.fromfile(prefix + ".single.u8", dtype=np.uint8)
    assert np.array_equal(single & 1, robot[:16]) and np.array_equal((single >> 2) & 1, both[:16])


def _synthetic_octree(seed=3, res=0.02):
    """A small octomap-like tree over [0, 1.28)^3 in begin_tree() order (depth first, parents before children): inner
    nodes, occupied and free leaves at the voxel size, at 2x and at 4x the voxel size (pruned nodes)."""
    rng = np.random.default_rng(seed)
    nodes = []

    def visit(cx, cy, cz, size, depth):
        leaf_here = size <= res or (size <= 4 * res and rng.random() < 0.25) or rng.random() < 0.15 * depth - 0.25
        occupied = rng.random() < 0.3
        nodes.append((cx, cy, cz, size, int(occupied), int(leaf_here)))
        if leaf_here:
            return
        h = size / 4
        for dx in (-h, h):
            for dy in (-h, h):
                for dz in (-h, h):
                    if rng.random() < 0.55:  # octomap only allocates children it has seen
                        visit(cx + dx, cy + dy, cz + dz, size / 2, depth + 1)

    visit(0.64, 0.64, 0.64, 1.28, 0)
    return np.array(nodes)


@pytest.mark.gpu
def test_octomap_world_object_through_the_env(gpu, oracle, tmp_path):
    """The OCTREE branch of updateDistanceObject (.cpp:1773-1790): with the reference env's semantics every node
    begin_tree() visits contributes its centre (Q9, types.cpp:355-366); with setOctreeOccupiedLeavesOnly the occupied
    leaves go through the GPU form of getOcTreePoints (distance_field.cpp:254-279, big leaves become a lattice).
    Both against the oracle's field fed with the points the reference would add; removal empties the field again."""
    exe = _build_driver()
    nodes = _synthetic_octree()
    assert len(nodes) > 300 and (nodes[:, 5] == 0).sum() > 20 and ((nodes[:, 3] > 0.03) & (nodes[:, 5] == 1)).sum() > 5
    path, prefix = str(tmp_path / "nodes.txt"), str(tmp_path / "oct")
    with open(path, "w") as f:
        for x, y, z, size, occ, leaf in nodes:
            f.write("%r %r %r %r %d %d\n" % (float(x), float(y), float(z), float(size), int(occ), int(leaf)))
    r = subprocess.run([exe, "--octree", path, prefix], capture_output=True, text=True)
    assert r.returncode == 0 and "octree test ok" in r.stdout, r.stderr[-2000:] + r.stdout
    dims = (50, 50, 50)

    def field_with(points):
        o = oracle.Field((1.0, 1.0, 1.0), 0.02, (0.0, 0.0, 0.0), 0.25)
        assert o.dims == dims
        o.add_points(points)
        o.make_exact()
        return o.d2()

    load = lambda name: np.fromfile(prefix + name, dtype=np.int32).reshape(dims)
    q9 = field_with(nodes[:, :3])
    assert np.array_equal(load(".q9.d2.i32"), q9) and np.array_equal(load(".late.d2.i32"), q9)
    occ_leaves = nodes[(nodes[:, 4] == 1) & (nodes[:, 5] == 1)][:, :4]
    leaves = field_with(oracle.octree_points(0.02, [tuple(l) for l in occ_leaves]))
    assert np.array_equal(load(".leaves.d2.i32"), leaves)
    assert not np.array_equal(q9, leaves)
    empty = field_with(np.zeros((0, 3)))
    assert np.array_equal(load(".q9_removed.d2.i32"), empty) and np.array_equal(load(".leaves_removed.d2.i32"), empty)


_PARAMS = """\
# parameter namespace of the B200 distance-field detector (what `rosparam load` would put under the plugin's namespace)
b200_distance_field:
  size_x: 2.0
  size_y: 2.5   # metres
  size_z: 1.0
  origin: {x: 0.1, y: -0.2, z: 0.5}
  resolution: 0.025
  max_propogation_distance: 0.3
  collision_tolerance: 0.005
  use_signed_distance_field: true
  octree_occupied_leaves_only: yes
  devices: [0, 1, 2]
  link_spheres:
    - link: base
      spheres:
        - {x: 0.0, y: 0.0, z: 0.05, radius: 0.08}
        - {x: 0.0, y: 0.0, z: 0.15, radius: 0.06}
        - {x: 0.0, y: 0.0, radius: 0.01}
    - link: arm
      spheres: none
    - spheres: none
    - link: tool
      spheres:
        - x: 0.1
          y: 0.2
          z: 0.3
          radius: 0.04
"""


def _run_config(tmp_path):
    exe = _build_driver()
    path, prefix = str(tmp_path / "params.yaml"), str(tmp_path / "cfg")
    open(path, "w").write(_PARAMS)
    r = subprocess.run([exe, "--config", path, prefix, "b200_distance_field"], capture_output=True, text=True)
    assert r.returncode == 0 and "config test ok" in r.stdout, r.stderr[-2000:] + r.stdout
    out = {}
    for line in open(prefix + ".config.txt"):
        k, _, v = line.strip().partition(" ")
        out.setdefault(k, []).append(v)
    return out


def test_plugin_configuration_reader(tmp_path):
    """The plugin's parameter namespace (b200_config.hpp): field geometry with the reference's constructor defaults
    (collision_env_distance_field.h:48-54) and `link_spheres` in the format - and with the warnings - of
    loadLinkBodySphereDecompositions (collision_distance_field_ros_helpers.h:46-120)."""
    c = _run_config(tmp_path)
    assert [float(x) for x in c["size"][0].split()] == [2.0, 2.5, 1.0]
    assert [float(x) for x in c["origin"][0].split()] == [0.1, -0.2, 0.5]
    assert float(c["resolution"][0]) == 0.025 and float(c["max_propogation_distance"][0]) == 0.3
    assert float(c["collision_tolerance"][0]) == 0.005 and c["use_signed_distance_field"] == ["1"]
    assert c["octree_occupied_leaves_only"] == ["1"] and c["devices"] == ["0 1 2"]
    links = {l.split()[0]: l.split()[1:] for l in c["link"]}
    assert set(links) == {"base", "arm", "tool"}
    assert links["arm"] == ["0"]  # spheres: none -> a link without spheres
    assert links["base"][0] == "2" and [float(x) for x in links["base"][1:]] == [0, 0, 0.05, 0.08, 0, 0, 0.15, 0.06]
    assert [float(x) for x in links["tool"][1:]] == [0.1, 0.2, 0.3, 0.04]
    assert "All spheres must specify a value for z" in c["warning"] and "All link spheres must have link" in c["warning"]


@pytest.mark.gpu
def test_configured_allocator_builds_the_env_from_the_parameters(gpu, tmp_path):
    c = _run_config(tmp_path)
    assert c["field_cells"] == ["80 100 40"]  # int(size * (1 / resolution)), voxel_grid.h:387
    assert float(c["field_max_distance"][0]) == 0.3
    spheres = dict(l.split() for l in c["env_spheres"])
    assert spheres == {"base": "2", "arm": "0"}  # the overrides replaced the derived spheres (.cpp:1085-1088)
