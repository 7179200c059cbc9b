"""Pins the CPU oracle against every restatable known-answer test the reference holds for this path
(SURVEY.md section 8(c)).  Each test names the reference test it restates.  CPU only."""
import zlib

import numpy as np
import pytest

from moveit_b200 import robot_description as rd

WIDTH = HEIGHT = DEPTH = 1.0
RESOLUTION = 0.1
MAX_DIST = 0.3
POINT1 = (0.1, 0.0, 0.0)
POINT2 = (0.0, 0.1, 0.2)
POINT3 = (0.4, 0.0, 0.0)


def small_field(oracle, signed=False, max_dist=MAX_DIST):
    return oracle.Field((WIDTH, HEIGHT, DEPTH), RESOLUTION, (0.0, 0.0, 0.0), max_dist, signed)


def check_distance_field(f, points):
    """test_distance_field.cpp:281-327 check_distance_field: every zero cell is one of the points' cells."""
    cells = {f.world_to_grid(*p)[1] for p in points}
    d2 = f.d2()
    zero = {tuple(int(v) for v in idx) for idx in np.argwhere(d2 == 0)}
    assert zero <= cells
    return zero


# ---- test_voxel_grid.cpp:43-95 -------------------------------------------------------------------------------
def test_voxel_grid_dimensions(oracle):
    f = oracle.Field((0.02, 0.02, 0.02), 0.01, (0, 0, 0), 0.05)
    assert f.dims == (2, 2, 2)
    # z fastest: x*stride1 + y*stride2 + z (voxel_grid.h:426-429) — the d2 array reshapes as [x][y][z]
    f.add_voxels([(1, 0, 1)])
    assert f.d2()[1, 0, 1] == 0 and (f.d2() == 0).sum() == 1


# ---- test_distance_field.cpp:329-429 TestAddRemovePoints ---------------------------------------------------------
def test_add_remove_points(oracle):
    f = small_field(oracle)
    assert f.dims == (int(WIDTH / RESOLUTION + 0.5),) * 3
    assert abs(f.distance_world(1000.0, 1000.0, 1000.0) - MAX_DIST) < 1e-4
    dist, grad, inb = f.distance_gradient(1000.0, 1000.0, 1000.0)
    assert abs(dist - MAX_DIST) < 1e-4 and not inb

    f.add_points([POINT1, POINT2])
    f.update_points([POINT1], [POINT2, POINT3])
    zero = check_distance_field(f, [POINT2, POINT3])
    assert len(zero) == 2
    f.remove_points([POINT2])
    zero = check_distance_field(f, [POINT3])
    assert len(zero) == 1

    f.reset()
    f.add_points([POINT1])
    first = True
    nx, ny, nz = f.dims
    for z in range(1, nz - 1):
        for x in range(1, nx - 1):
            for y in range(1, ny - 1):
                dist = f.distance_cell(x, y, z)
                w = f.grid_to_world(x, y, z)
                dg, grad, inb = f.distance_gradient(*w)
                assert inb
                assert abs(dist - dg) < 1e-4
                if 0 < dist < MAX_DIST:
                    n = np.linalg.norm(grad)
                    comp = np.array(w) - grad / n * dist
                    first = False
                    assert np.all(np.abs(comp - np.array(POINT1)) <= RESOLUTION + 1e-12)
    assert not first


# ---- test_distance_field.cpp:431-499 TestSignedAddRemovePoints (block minus centre == fresh build) -------------
def _block_points(f):
    lw = f.grid_to_world(1, 1, 1)
    hw = f.grid_to_world(8, 8, 8)
    pts = []
    x = lw[0]
    while x <= hw[0]:
        y = lw[1]
        while y <= hw[1]:
            z = lw[2]
            while z <= hw[2]:
                pts.append((x, y, z))
                z += 0.1
            y += 0.1
        x += 0.1
    return pts


def test_signed_remove_center_equals_fresh_build(oracle):
    f = small_field(oracle, signed=True)
    pts = _block_points(f)
    f.reset()
    f.add_points(pts)
    center = f.grid_to_world(5, 5, 5)
    f.remove_points([center])
    g = small_field(oracle, signed=True)
    g.add_points([p for p in pts if p != center])
    # areDistanceFieldsDistancesEqual (:167-197): both d2 arrays exactly equal
    assert np.array_equal(f.d2(), g.d2())
    assert np.array_equal(f.d2(True), g.d2(True))


# ---- test_distance_field.cpp:500-598 signed sphere: negative interior, nearest cells, gradients ------------------
def test_signed_sphere_gradients(oracle):
    f = small_field(oracle, signed=True)
    pts = oracle.shape_points(oracle.SHAPE_SPHERE, (0.25,), oracle.pose12((0.5, 0.5, 0.5)), RESOLUTION)
    f.add_points(pts)
    neg = f.d2(True)
    pos = f.d2()
    assert neg[5, 5, 5] > 1
    nx, ny, nz = f.dims
    for z in range(1, nz - 1):
        for x in range(1, nx - 1):
            for y in range(1, ny - 1):
                dist = f.distance_cell(x, y, z)
                r, ncell_dist, ncell_pos = f.nearest_cell(x, y, z)
                assert ncell_dist == dist
                if r == 0:
                    if ncell_dist > 0:
                        assert ncell_dist >= f.uninitialized_distance()
                    elif ncell_dist < 0:
                        assert ncell_dist <= -f.uninitialized_distance()
                if neg[x, y, z] > 0:
                    assert dist < 0
                    w = f.grid_to_world(x, y, z)
                    dg, grad, inb = f.distance_gradient(*w)
                    assert inb and abs(dist - dg) < 1e-4
                    if r == 0:
                        continue
                    assert pos[ncell_pos] >= 1
                    g2 = float(grad @ grad)
                    if g2 < np.finfo(float).eps:
                        continue
                    comp = np.array(w) - grad / np.sqrt(g2) * dist
                    ok, cell = f.world_to_grid(*comp)
                    assert ok
                    assert pos[cell] >= 1


# ---- test_distance_field.cpp:601-649 TestShape -------------------------------------------------------------------
def test_shape_add_move(oracle):
    f = small_field(oracle, signed=True)
    p = oracle.pose12((0.5, 0.5, 0.5))
    q = oracle.pose12((0.7, 0.7, 0.7))
    pts1 = oracle.shape_points(oracle.SHAPE_SPHERE, (0.25,), p, RESOLUTION)
    f.add_points(pts1)  # addShapeToField = getShapePoints + addPointsToField (distance_field.cpp:221-226)
    zero = check_distance_field(f, pts1)
    assert len(zero) > 0
    assert np.all(f.d2(True)[f.d2() == 0] > 0)
    pts2 = oracle.shape_points(oracle.SHAPE_SPHERE, (0.25,), q, RESOLUTION)
    f.add_points(pts2)
    check_distance_field(f, np.vstack([pts2, pts1]))
    f.update_points(pts1, pts2)  # moveShapeInField (distance_field.cpp:289-308)
    check_distance_field(f, pts2)
    g = small_field(oracle, signed=True)
    g.add_points(pts2)
    assert np.array_equal(f.d2(), g.d2())
    assert np.array_equal(f.d2(True), g.d2(True))


# ---- test_distance_field.cpp:843-858 low-resolution octree leaves ------------------------------------------------
def test_octree_low_res_leaf_expansion(oracle):
    f = oracle.Field((3.0, 3.0, 4.0), 0.02, (0, 0, 0), 0.25)
    # octomap leaf centres of a 0.05 tree for the points (.5,.5,.5), (.7,.5,.5), (1.0,.5,.5):
    # key = floor(x/res), centre = (key + 0.5) * res
    def centre(v):
        return (np.floor(v / 0.05) + 0.5) * 0.05
    leaves = [(centre(0.5), centre(0.5), centre(0.5), 0.05), (centre(0.7), centre(0.5), centre(0.5), 0.05),
              (centre(1.0), centre(0.5), centre(0.5), 0.05)]
    pts = oracle.octree_points(0.02, leaves)
    f.add_points(pts)
    assert int((f.d2() == 0).sum()) == 3 * 4 * 4 * 4


# ---- test_distance_field.cpp:882-928 TestReadWrite ---------------------------------------------------------------
def test_read_write_round_trip(oracle):
    f = small_field(oracle)
    f.add_points([POINT1, POINT2, POINT3])
    payload = zlib.compress(f.pack_occupancy().tobytes())
    g = small_field(oracle, max_dist=0.25)  # small_df2(si, PERF_MAX_DIST, false): different max distance
    g.unpack_occupancy(np.frombuffer(zlib.decompress(payload), dtype=np.uint8))
    # areDistanceFieldsDistancesEqual(small_df, small_df2) holds in the reference: ceil(.3/.1)=ceil(.25/.1)=3
    assert np.array_equal(f.d2(), g.d2())

    big = oracle.Field((3.0, 3.0, 4.0), 0.02, (0, 0, 0), 0.25)
    pts = oracle.shape_points(oracle.SHAPE_SPHERE, (0.5,), oracle.pose12((0.5, 0.5, 0.5)), 0.02)
    big.add_points(pts)
    blob = big.pack_occupancy()
    assert len(blob) == 150 * 150 * 25
    big2 = oracle.Field((3.0, 3.0, 4.0), 0.02, (0, 0, 0), 0.25)
    big2.unpack_occupancy(blob)
    assert np.array_equal(big.d2(), big2.d2())
    big3 = oracle.Field((3.0, 3.0, 4.0), 0.02, (0, 0, 0), 0.25 + 0.02)
    big3.unpack_occupancy(blob)
    assert not np.array_equal(big.d2(), big3.d2())


def test_read_write_stream_format(oracle):
    """The same test through the complete on-disk format (writeToStream :636-673 / readFromStream :675-761): text
    header + one zlib stream, and the product's host-side framing (moveit_b200/field_stream.py) against it."""
    import io
    from moveit_b200 import field_stream
    f = small_field(oracle)
    f.add_points([POINT1, POINT2, POINT3])
    blob = f.write_stream()
    head = blob[:blob.index(b"origin_z")].decode()
    assert head == "resolution: 0.1\nsize_x: 1\nsize_y: 1\nsize_z: 1\norigin_x: 0\norigin_y: 0\n"
    g = oracle.Field.from_stream(blob, 0.25)  # small_df2(si, PERF_MAX_DIST, false)
    assert g is not None and g.dims == f.dims and np.array_equal(f.d2(), g.d2())
    # the product's framing reads what the oracle wrote ...
    hdr, payload = field_stream.read_stream(io.BytesIO(blob))
    assert hdr == dict(resolution=0.1, size_x=1.0, size_y=1.0, size_z=1.0, origin_x=0.0, origin_y=0.0, origin_z=0.0)
    assert payload == f.pack_occupancy().tobytes() and len(payload) == field_stream.payload_size(f.dims)
    # ... and writes what the oracle reads, byte for byte the same header
    out = io.BytesIO()
    field_stream.write_stream(out, 0.1, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), payload)
    assert out.getvalue()[:len(head)] == head.encode()
    h = oracle.Field.from_stream(out.getvalue(), 0.3)
    assert h is not None and np.array_equal(h.d2(), f.d2())
    # non-trivial numbers take the default six-digit formatting of operator<<(double)
    odd = oracle.Field((0.7, 0.35, 0.21), 0.07, (-0.123456789, 1e-7, 12345678.0), 0.2)
    odd.add_points([(0.0, 0.1, 12345678.1)])
    oblob = odd.write_stream()
    otext = oblob[:oblob.index(b"\n", oblob.index(b"origin_z")) + 1].decode()
    assert otext == ("resolution: 0.07\nsize_x: 0.7\nsize_y: 0.35\nsize_z: 0.21\norigin_x: -0.123457\n"
                     "origin_y: 1e-07\norigin_z: 1.23457e+07\n")
    out = io.BytesIO()
    field_stream.write_stream(out, 0.07, (0.7, 0.35, 0.21), (-0.123456789, 1e-7, 12345678.0), odd.pack_occupancy())
    assert out.getvalue()[:len(otext)] == otext.encode()
    # a stream that is not a field, an exhausted stream and a truncated payload are refused (returns false, no crash)
    assert oracle.Field.from_stream(b"", 0.25) is None and field_stream.read_stream(io.BytesIO(b"")) is None
    assert oracle.Field.from_stream(b"resolution 0.1\n", 0.25) is None
    assert field_stream.read_stream(io.BytesIO(b"resolution 0.1\n")) is None
    cut = blob[:len(blob) - 6]
    assert oracle.Field.from_stream(cut, 0.25) is None and field_stream.read_stream(io.BytesIO(cut)) is None


# ---- test_collision_distance_field.cpp:336-383 DistanceFieldCollisionDetectionPadding.Sphere ------------------
def _padding_sphere_env(oracle):
    r = rd.RobotDescription("test")
    r.add_link("base", None, "root_joint", rd.JOINT_FIXED)
    r.add_shape("base", rd.SHAPE_SPHERE, (0.04,))
    kin = r.kinematics_arrays()
    # BodyDecomposition(shapes, poses, resolution, getLinkPadding) with padding 0.04 (cpp:1080-1081)
    sph, pts, bs = oracle.body_decomposition(r.shapes[0], 0.02, 0.04)
    kin.update(spheres=[sph], points=[pts], bsphere=[bs], has_geometry=[1])
    model = oracle.Model(kin)
    env = oracle.Env(model, (0.1, 0.1, 0.2), (0.0, 0.0, 0.0), res=0.02, tol=0.0, max_prop=0.25)
    acm = np.zeros((1, 1), dtype=np.int32)  # AllowedCollisionMatrix(names) -> entries NEVER
    env.set_group([0], acm, np.zeros(0))
    return env, sph


def test_padding_sphere_known_answer(oracle):
    env, sph = _padding_sphere_env(oracle)
    assert sph.shape == (1, 4) and sph[0, 3] == 0.04 + 0.04
    assert env.world.dims == (5, 5, 10)
    box = (oracle.SHAPE_BOX, (0.02, 0.02, 0.02))

    def set_box(z):
        # World object route: getBodyDecompositionCacheEntry(shape, res) -> BodyDecomposition(padding 0.01) in the
        # body frame, then PosedBodyPointDecomposition(bd, pose) (cpp:1782-1784, types.cpp:368-379)
        _, pts, _ = oracle.body_decomposition([(box[0], box[1], oracle.pose12())], 0.02, 0.01)
        pts = pts + np.array([0.0, 0.0, z])  # identity rotation: pose * p = p + t
        return pts

    q = np.zeros((1, 0))
    pts0 = set_box(0.0)
    env.world.add_points(pts0)
    assert env.check_batch(q, mode=1)[0] == 1
    pts1 = set_box(0.08)  # setObjectPose -> MOVE_SHAPE: remove old points, add new (cpp:1735-1739)
    env.world.remove_points(pts0)
    env.world.add_points(pts1)
    assert env.check_batch(q, mode=1)[0] == 1
    pts2 = set_box(0.1)
    env.world.remove_points(pts1)
    env.world.add_points(pts2)
    assert env.check_batch(q, mode=1)[0] == 0


def test_q19_lattice_worked_example(oracle):
    """box 0.02^3 + 0.01 padding -> only the 3rd and 4th lattice values pass |x| <= 0.02."""
    _, pts, _ = oracle.body_decomposition([(oracle.SHAPE_BOX, (0.02, 0.02, 0.02), oracle.pose12())], 0.02, 0.01)
    xs = sorted(set(pts[:, 0]))
    assert len(xs) == 2 and len(pts) == 8
    assert xs[0] == -0.019999999999999993 and abs(xs[1]) < 1e-16


# ---- robot_state_test.cpp:432-559 OneRobot.FK --------------------------------------------------------------------
def _one_robot(oracle):
    r = rd.one_robot()
    kin = r.kinematics_arrays()
    L = r.n_links
    kin.update(spheres=[np.zeros((0, 4))] * L, points=[np.zeros((0, 3))] * L, bsphere=[np.zeros(4)] * L,
               has_geometry=[0] * L)
    return r, oracle.Model(kin)


def _quat(R):
    """Eigen::Quaterniond(rotation matrix) for the well-conditioned (trace > 0) case used by the test."""
    t = np.trace(R)
    assert t > 0
    s = np.sqrt(t + 1.0) * 2
    return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])


def test_one_robot_fk(oracle):
    r, model = _one_robot(oracle)
    assert r.nvar == 7
    assert r.variable_names == ["base_joint/x", "base_joint/y", "base_joint/theta", "joint_a", "joint_c", "mim_f",
                                "joint_f"]
    name = {n: i for i, n in enumerate(r.link_names)}

    def state(**kw):
        q = np.zeros(7)
        for k, v in kw.items():
            q[r.variable_names.index(k.replace("__", "/"))] = v
        # RobotState keeps mimic variables consistent (robot_state.h:1854-1863)
        q[r.variable_names.index("mim_f")] = 1.5 * q[r.variable_names.index("joint_f")] + 0.1
        return q

    T = model.fk(state(base_joint__x=1.0, base_joint__y=1.0, base_joint__theta=0.5, joint_a=-0.5, joint_c=0.08))
    exp = {
        "base_link": ((1, 1, 0), (0.0, 0.0, 0.247404, 0.968912)),
        "link_a": ((1, 1, 0), (0.0, 0.0, 0.0, 1.0)),
        "link_b": ((1, 1.5, 0), (0.0, -0.2084598, 0.0, 0.97803091)),
        "link_c": ((1.08, 1.4, 0), (0.0, 0.0, 0.0, 1.0)),
    }
    for link, (trans, quat) in exp.items():
        M = T[name[link]]
        assert np.allclose(M[:, 3], trans, atol=1e-5), link
        assert np.allclose(_quat(M[:, :3]), quat, atol=1e-5), link

    # mimic joints, default values
    T = model.fk(state())
    assert np.allclose(T[name["link_c"]][:, 3], (0.0, 0.4, 0), atol=1e-5)
    assert np.allclose(T[name["link_d"]][:, 3], (0.2, 0.5, 0), atol=1e-5)
    assert np.allclose(T[name["link_e"]][:, 3], (0.3, 0.6, 0), atol=1e-5)
    T = model.fk(state(joint_f=1.0))
    assert np.allclose(T[name["link_c"]][:, 3], (0.0, 0.4, 0), atol=1e-5)
    assert np.allclose(T[name["link_d"]][:, 3], (1.7, 0.5, 0), atol=1e-5)
    assert np.allclose(T[name["link_e"]][:, 3], (2.8, 0.6, 0), atol=1e-5)
    # the engine derives the mimic value itself: a stale mim_f entry in q must not matter
    q = state(joint_f=1.0)
    q[5] = 123.0
    assert np.array_equal(model.fk(q), T)


def test_updated_link_names_one_robot():
    """robot_state_test.cpp:462-466: every group of the fixture updates all seven links."""
    r = rd.one_robot()
    all_links = list(range(r.n_links))
    assert r.descendant_links_of_joints({"base_joint", "joint_a", "joint_c"}) == all_links
    assert r.descendant_links_of_joints({"base_joint", "joint_a", "joint_b"}) == all_links


# ---- test_collision_common_panda.h:54-66,122-136 qualitative sanity (FCL/Bullet on meshes in the reference) -----
def test_panda_home_free_and_folded_self_colliding(oracle):
    from moveit_b200.engine import assemble_model, resolve_acm
    r = rd.panda()
    md = assemble_model(r, oracle.body_decomposition)
    model = oracle.Model(md)
    env = oracle.Env(model, (3.0, 3.0, 4.0), (0.0, 0.0, 0.0))
    env.set_group(r.group_links(""), r.acm_entry_matrix(), np.zeros(r.nvar))
    home = np.zeros(r.nvar)
    home[[1, 3, 5, 6]] = (-0.785, -2.356, 1.571, 0.785)
    assert env.check_batch(home, mode=3)[0] == 0
    folded = np.zeros(r.nvar)
    folded[1], folded[3] = 0.15, -3.0
    # the reference test applies j2=0.15, j4=-3.0 on top of the home pose
    folded2 = home.copy()
    folded2[1], folded2[3] = 0.15, -3.0
    assert env.check_batch(folded2, mode=3)[0] == 1
    # host ACM resolution matches the oracle's restatement of generateDistanceFieldCacheEntry
    se, ie = resolve_acm(md, r.group_links(""), r.acm_entry_matrix())
    ose, oie = env.group_tables()
    assert np.array_equal(se, ose)
    iu = np.triu_indices(len(se), 1)
    assert np.array_equal(ie[iu], oie[iu])
