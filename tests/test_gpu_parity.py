"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Everything here needs a B200."""
import numpy as np
import pytest

from moveit_b200 import robot_description as rd
from moveit_b200 import scenarios
from moveit_b200.binding import CHECK_INTRA, CHECK_ROBOT, CHECK_SELF

from common import EngineSetup, OracleSetup, oracle_world_points, self_field_points

pytestmark = pytest.mark.gpu

SMALL = dict(size=(1.0, 1.0, 1.0), resolution=0.1, origin=(0.0, 0.0, 0.0), max_distance=0.3)
POINT1, POINT2, POINT3 = (0.1, 0.0, 0.0), (0.0, 0.1, 0.2), (0.4, 0.0, 0.0)


def small_pair(oracle, signed=False):
    from moveit_b200 import engine
    g = engine.Field(SMALL["size"], SMALL["resolution"], SMALL["origin"], SMALL["max_distance"], signed)
    o = oracle.Field(SMALL["size"], SMALL["resolution"], SMALL["origin"], SMALL["max_distance"], signed)
    return g, o


# ------------------------------------------------------------------------------------------------ kernel 5: field
def test_field_dims_and_empty(gpu, oracle):
    g, o = small_pair(oracle)
    assert g.dims == o.dims and g.max_d2 == o.max_d2 == 9
    assert np.array_equal(g.d2(), o.d2())  # all max_distance_sq_
    d, grad, inb = g.distance_gradient([(1000.0, 1000.0, 1000.0)])
    assert abs(d[0] - 0.3) < 1e-4 and not inb[0] and np.all(grad == 0)


def test_add_update_remove_points_sequence(gpu, oracle):
    """test_distance_field.cpp:329-375 on the GPU: zero sets and d2 equal to the oracle after every step."""
    g, o = small_pair(oracle)
    for f in (g, o):
        f.add_points([POINT1, POINT2])
    assert np.array_equal(g.d2(), o.d2())
    for f in (g, o):
        f.update_points([POINT1], [POINT2, POINT3])
    assert np.array_equal(g.d2(), o.d2())
    for f in (g, o):
        f.remove_points([POINT2])
    assert np.array_equal(g.d2(), o.d2())
    assert int((g.d2() == 0).sum()) == 1


def test_field_gradients_match_oracle(gpu, oracle):
    g, o = small_pair(oracle)
    for f in (g, o):
        f.add_points([POINT1])
    nx, ny, nz = g.dims
    pts = np.array([o.grid_to_world(x, y, z) for x in range(nx) for y in range(ny) for z in range(nz)])
    d, grad, inb = g.distance_gradient(pts)
    for k in range(0, len(pts), 7):
        od, og, oi = o.distance_gradient(*pts[k])
        assert d[k] == od and np.array_equal(grad[k], og) and inb[k] == oi


@pytest.mark.parametrize("shape,max_dist,res", [((40, 33, 27), 0.25, 0.02), ((24, 24, 50), 0.25, 0.01)])
def test_edt_bit_exact_vs_brute_force(gpu, oracle, shape, max_dist, res):
    from moveit_b200 import engine
    size = [s * res + 1e-9 for s in shape]
    g = engine.Field(size, res, (0, 0, 0), max_dist)
    assert g.dims == shape
    rng = np.random.default_rng(5)
    ijk = np.stack([rng.integers(0, s, 60) for s in shape], axis=1)
    g.add_voxels(ijk)
    occ = np.zeros(shape, dtype=np.uint8)
    occ[ijk[:, 0], ijk[:, 1], ijk[:, 2]] = 1
    assert np.array_equal(g.occupancy(), occ)
    ref = oracle.edt_exact(occ, g.max_d2, brute=True)
    assert np.array_equal(g.d2(), ref)
    assert np.array_equal(oracle.edt_exact(occ, g.max_d2), ref)


def test_signed_field_matches_oracle(gpu, oracle):
    """test_distance_field.cpp:431-499 on the GPU (solid block: propagation == exact)."""
    g, o = small_pair(oracle, signed=True)
    lw, hw = o.grid_to_world(1, 1, 1), o.grid_to_world(8, 8, 8)
    pts = [(x, y, z) for x in np.arange(lw[0], hw[0] + 1e-9, 0.1) for y in np.arange(lw[1], hw[1] + 1e-9, 0.1)
           for z in np.arange(lw[2], hw[2] + 1e-9, 0.1)]
    for f in (g, o):
        f.add_points(pts)
        f.remove_points([o.grid_to_world(5, 5, 5)])
    assert np.array_equal(g.d2(), o.d2())
    assert np.array_equal(g.d2(True), o.d2(True))
    q = np.array([o.grid_to_world(x, y, z) for x in range(1, 9) for y in range(1, 9) for z in range(1, 9)])
    d, grad, inb = g.distance_gradient(q)
    for k in range(0, len(q), 5):
        od, og, oi = o.distance_gradient(*q[k])
        assert d[k] == od and np.array_equal(grad[k], og)


def test_pack_unpack_occupancy(gpu, oracle):
    g, o = small_pair(oracle)
    for f in (g, o):
        f.add_points([POINT1, POINT2, POINT3])
    assert np.array_equal(g.pack_occupancy(), o.pack_occupancy())
    g2, _ = small_pair(oracle)
    g2.unpack_occupancy(o.pack_occupancy())
    assert np.array_equal(g2.d2(), o.d2())


# ------------------------------------------------------------------------------------------- rasteriser (a14)
@pytest.mark.parametrize("k", range(3))
def test_rasteriser_matches_oracle(gpu, oracle, k):
    from moveit_b200 import engine
    objs = scenarios.random_scene(seed=100 + k)
    fa = scenarios.FIELD_A
    g = engine.Field.centered(fa["size"], fa["resolution"], fa["center"], fa["max_distance"])
    o = oracle.Field(fa["size"], fa["resolution"], scenarios.field_corner(fa), fa["max_distance"])
    for t, dims, pose in objs:
        g.add_posed_body(t, dims, pose)
    o.add_points(oracle_world_points(oracle, objs, fa["resolution"]))
    assert np.array_equal(g.occupancy(), (o.d2() == 0).astype(np.uint8))
    # direct (field-frame) rasterisation: addShapeToField
    g2 = engine.Field.centered(fa["size"], fa["resolution"], fa["center"], fa["max_distance"])
    o2 = oracle.Field(fa["size"], fa["resolution"], scenarios.field_corner(fa), fa["max_distance"])
    for t, dims, pose in objs:
        g2.add_shape(t, dims, pose)
        o2.add_points(oracle.shape_points(t, dims, pose, fa["resolution"]))
    assert np.array_equal(g2.occupancy(), (o2.d2() == 0).astype(np.uint8))
    g2.remove_shape(*objs[0][:3])
    o2.remove_points(oracle.shape_points(objs[0][0], objs[0][1], objs[0][2], fa["resolution"]))
    assert np.array_equal(g2.occupancy(), (o2.d2() == 0).astype(np.uint8))


def test_body_decomposition_matches_oracle(gpu, oracle):
    from moveit_b200 import engine
    for desc in (rd.panda(), rd.pr2_like()):
        for i in desc.links_with_geometry():
            for pad in (0.0, 0.01):
                s1, p1, b1 = engine.body_decomposition(desc.shapes[i], 0.02, pad)
                s2, p2, b2 = oracle.body_decomposition(desc.shapes[i], 0.02, pad)
                assert np.array_equal(s1, s2) and np.array_equal(p1, p2) and np.array_equal(b1, b2)
    # multi-shape link: merged bounding sphere
    shapes = [(rd.SHAPE_BOX, (0.1, 0.2, 0.3), rd.rpy_xyz_to_pose((0.1, 0, 0), (0.3, 0.2, 0.1))),
              (rd.SHAPE_SPHERE, (0.07,), rd.rpy_xyz_to_pose((-0.2, 0.1, 0.0))),
              (rd.SHAPE_CYLINDER, (0.05, 0.4), rd.rpy_xyz_to_pose((0, 0, 0.3), (1.0, 0, 0)))]
    s1, p1, b1 = engine.body_decomposition(shapes, 0.02, 0.01)
    s2, p2, b2 = oracle.body_decomposition(shapes, 0.02, 0.01)
    assert np.array_equal(s1, s2) and np.array_equal(p1, p2) and np.array_equal(b1, b2)


# ------------------------------------------------------------------------------------------------- kernel 1: FK
@pytest.mark.parametrize("robot", ["one_robot", "panda", "pr2_like"])
def test_fk_matches_oracle(gpu, oracle, robot):
    from moveit_b200 import engine
    desc = getattr(rd, robot)()
    md = assemble_model(desc, oracle)
    gm, om = engine.Model(md), oracle.Model(md)
    q = desc.sample_states(256, 7)
    T = gm.fk(q)
    ref = np.stack([om.fk(qi) for qi in q])
    # identical operation order; only sin/cos may differ in the last ulp between CUDA and glibc
    assert np.max(np.abs(T - ref)) < 1e-13
    T32 = gm.fk(q.astype(np.float32))
    ref32 = np.stack([om.fk(qi) for qi in q.astype(np.float32).astype(np.float64)])
    assert np.max(np.abs(T32 - ref32)) < 1e-13


def assemble_model(desc, oracle, res=0.02):
    from moveit_b200.engine import assemble_model as am
    return am(desc, oracle.body_decomposition, res)


def test_one_robot_fk_known_answers_on_gpu(gpu, oracle):
    """robot_state_test.cpp:484-543 through b200df_model_fk."""
    from moveit_b200 import engine
    desc = rd.one_robot()
    gm = engine.Model(assemble_model(desc, oracle))
    name = {n: i for i, n in enumerate(desc.link_names)}
    q = np.zeros(7)
    q[[0, 1, 2, 3, 4]] = (1.0, 1.0, 0.5, -0.5, 0.08)
    T = gm.fk(q)[0]
    assert np.allclose(T[name["link_b"]][:, 3], (1, 1.5, 0), atol=1e-5)
    assert np.allclose(T[name["link_c"]][:, 3], (1.08, 1.4, 0), atol=1e-5)
    q = np.zeros(7)
    q[6] = 1.0
    T = gm.fk(q)[0]
    assert np.allclose(T[name["link_d"]][:, 3], (1.7, 0.5, 0), atol=1e-5)
    assert np.allclose(T[name["link_e"]][:, 3], (2.8, 0.6, 0), atol=1e-5)


# ------------------------------------------------------------------------------------- kernels 2-4: verdicts
def test_sphere_centers_match_oracle(gpu, oracle):
    desc = rd.panda()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A)
    e = EngineSetup(desc, scenarios.FIELD_A, model_desc=o.model_desc)
    q = desc.sample_states(128, 11)
    c = e.group.sphere_centers(q)
    ref = np.stack([o.env.sphere_centers(qi) for qi in q])
    assert c.shape == ref.shape == (128, 60, 3)
    assert np.max(np.abs(c - ref)) < 1e-13


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_config1_robot_verdicts_bit_exact(gpu, oracle, dtype):
    """BASELINE config 1: Panda, 10k random states, Field A, 8 boxes + 4 cylinders, checkRobotCollision."""
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs)
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    assert np.array_equal(e.world.d2(), o.env.world.d2()), "exact EDT differs from the reference propagation here"
    q = desc.sample_states(10000, scenarios.SEED_STATES, dtype)
    got = e.group.check(q, e.world, flags=CHECK_ROBOT)
    ref = o.env.check_batch(q.astype(np.float64), mode=1)
    assert np.array_equal(got, ref)
    assert 0.05 < ref.mean() < 0.95  # the workload exercises both outcomes


def test_config2_robot_plus_self_verdicts_bit_exact(gpu, oracle):
    """BASELINE config 2 at oracle-sized N: robot + intra-group pairs from the SRDF-style ACM."""
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    q = desc.sample_states(20000, scenarios.SEED_STATES + 7)
    got = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    ref = o.env.check_batch(q, mode=3, n_threads=4)
    assert np.array_equal(got, ref)
    intra_only = e.group.check(q, flags=CHECK_INTRA)
    ref_self = o.env.check_batch(q, mode=2, n_threads=4)
    assert np.array_equal(intra_only, ref_self)
    assert 0.02 < ref_self.mean() < 0.98
    # checkCollision order (self, intra, env) gives the same verdict as the PlanningScene order
    assert np.array_equal(o.env.check_batch(q[:2000], mode=4), ref[:2000])


def test_group_with_self_field(gpu, oracle):
    """Planning group panda_arm: link0 is outside the group -> its points build the self field (cpp:913-973)."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=77)
    q0 = np.zeros(desc.nvar)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, group="panda_arm", self_state=q0)
    sp = self_field_points(o.model_desc, desc, o.group_links, o.model.fk(q0))
    e = EngineSetup(desc, scenarios.FIELD_A, objs, group="panda_arm", self_points=sp)
    assert np.array_equal(e.self_field.d2(), o.env.self_field().d2())
    q = desc.sample_states(8000, 99)
    got = e.group.check(q, e.world, e.self_field, flags=CHECK_ROBOT | CHECK_SELF | CHECK_INTRA)
    ref = o.env.check_batch(q, mode=3, n_threads=4)
    assert np.array_equal(got, ref)
    only_self = e.group.check(q, None, e.self_field, flags=CHECK_SELF)
    assert only_self.sum() > 0  # arm spheres do reach link0's voxels


def test_two_pass_fast_equals_exact_and_oracle(gpu, oracle):
    """Large FAST batches run in two passes (field lookups of every state, then the sphere pairs of the states left at
    0 through a device-side list).  Batches above the switch-over at 2^17 states (2^18 + a ragged tail, fp32 and fp64 rows,
    world only / world + self field) against the fp64 kernel on every state and against the oracle on a slice; the
    colliding set must not depend on which form ran (the same states in a batch below the switch-over)."""
    from moveit_b200.binding import PRECISION_FAST, PRECISION_EXACT
    desc = rd.panda()
    objs = scenarios.random_scene(seed=77)
    q0 = np.zeros(desc.nvar)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, group="panda_arm", self_state=q0)
    sp = self_field_points(o.model_desc, desc, o.group_links, o.model.fk(q0))
    e = EngineSetup(desc, scenarios.FIELD_A, objs, group="panda_arm", self_points=sp)
    n = (1 << 18) + 12345
    q32 = desc.sample_states(n, 4242, np.float32)
    for flags, fields in ((CHECK_ROBOT | CHECK_INTRA, (e.world, None)),
                          (CHECK_ROBOT | CHECK_SELF | CHECK_INTRA, (e.world, e.self_field)),
                          (CHECK_SELF | CHECK_INTRA, (None, e.self_field))):
        fast = e.group.check(q32, fields[0], fields[1], flags=flags, precision=PRECISION_FAST)
        exact = e.group.check(q32, fields[0], fields[1], flags=flags, precision=PRECISION_EXACT)
        assert np.array_equal(fast, exact), flags
        assert set(np.unique(fast)) <= {0, 1}
        if flags == (CHECK_ROBOT | CHECK_INTRA):
            assert 0.05 < fast.mean() < 0.95 and 0 < (fast & ~e.group.check(q32, e.world, flags=CHECK_ROBOT)).sum()
        below = e.group.check(q32[:100_000], fields[0], fields[1], flags=flags, precision=PRECISION_FAST)  # one sweep
        assert np.array_equal(below, fast[:100_000])
    q64 = q32.astype(np.float64) + 1e-9  # not representable in fp32: the undecided states are redone on these values
    flags = CHECK_ROBOT | CHECK_SELF | CHECK_INTRA
    fast64 = e.group.check(q64, e.world, e.self_field, flags=flags, precision=PRECISION_FAST)
    assert np.array_equal(fast64, e.group.check(q64, e.world, e.self_field, flags=flags, precision=PRECISION_EXACT))
    ref = o.env.check_batch(q64[-20000:], mode=3, n_threads=8)
    assert np.array_equal(fast64[-20000:], ref)


def test_signed_world_field_verdicts(gpu, oracle):
    """use_signed_distance_field: the verdict kernels skip the negative field when it cannot matter (radius >
    tolerance: in an occupied cell both predicates collide, in a free cell the negative distance is 0) - every kernel
    against the oracle's full signed predicate and against the debug sweep, which keeps the fp64 signed evaluation;
    with a tolerance above the radii the negative field does matter and the fp64 path is taken."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=5)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, signed=True)
    e = EngineSetup(desc, scenarios.FIELD_A, objs, signed=True)
    assert np.array_equal(e.world.d2(), o.env.world.d2())
    assert np.array_equal(e.world.d2(True), o.env.world.d2(True))
    q = desc.sample_states(5000, 3)
    assert np.array_equal(e.group.check(q, e.world, flags=CHECK_ROBOT), o.env.check_batch(q, mode=1))
    big = desc.sample_states(200_000, 4, np.float32)
    ref = o.env.check_batch(big[:20000].astype(np.float64), mode=1, n_threads=8)
    sweep = e.group.check(big, e.world, flags=CHECK_ROBOT, precision=2)
    assert np.array_equal(sweep[:20000], ref) and 0 < ref.sum() < len(ref)
    for precision in (0, 1, 4):
        assert np.array_equal(e.group.check(big, e.world, flags=CHECK_ROBOT, precision=precision), sweep), precision
    raw = e.group.check(big, e.world, flags=CHECK_ROBOT, precision=3)  # FAST now serves signed fields
    assert np.array_equal(raw[raw != 2], sweep[raw != 2]) and (raw == 2).mean() < 0.05
    # tolerance 0.08 > every radius... the spheres inside obstacles still collide through the negative distance
    ot = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, signed=True, tol=0.08)
    et = EngineSetup(desc, scenarios.FIELD_A, objs, signed=True, tol=0.08)
    reft = ot.env.check_batch(big[:20000].astype(np.float64), mode=1, n_threads=8)
    for precision in (0, 1, 2, 4):
        got = et.group.check(big[:20000], et.world, flags=CHECK_ROBOT, precision=precision)
        assert np.array_equal(got, reft), precision


def test_tolerance_and_no_acm(gpu, oracle):
    desc = rd.panda()
    objs = scenarios.random_scene(seed=9)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, use_acm=False, tol=0.015, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, scenarios.FIELD_A, objs, use_acm=False, tol=0.015)
    q = desc.sample_states(3000, 4)
    assert np.array_equal(e.group.check(q, e.world, flags=CHECK_ROBOT), o.env.check_batch(q, mode=1))
    # without an ACM adjacent links are tested too -> (nearly) every state self-collides in both
    assert np.array_equal(e.group.check(q, flags=CHECK_INTRA), o.env.check_batch(q, mode=2))


def test_derived_spheres_from_bounding_cylinders(gpu, oracle):
    """determineCollisionSpheres route (no link_body_decompositions override), a3."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=21)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, sphere_overrides=False, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, scenarios.FIELD_A, objs, sphere_overrides=False)
    q = desc.sample_states(4000, 8)
    assert np.array_equal(e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA),
                          o.env.check_batch(q, mode=3, n_threads=4))


def test_config4_pr2_like_full_acm(gpu, oracle):
    """BASELINE config 4: PR2-like dual-arm tree, all pairs enabled except parent-child."""
    desc = rd.pr2_like()
    objs = scenarios.random_scene(seed=31)
    fd = scenarios.FIELD_DEFAULT
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, fd, objs)
    q = desc.sample_states(6000, 12)
    got = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    ref = o.env.check_batch(q, mode=3, n_threads=4)
    assert np.array_equal(got, ref)
    arms = desc.group_links("arms")
    assert len(arms) < desc.n_links


@pytest.mark.parametrize("robot", ["panda", "pr2_like"])
def test_fused_kernel_agrees_with_independent_sweep(gpu, oracle, robot):
    """The fused verdict kernel (pairs folded into the FK sweep, sparse joint products, tight link cull) against the
    independently written round-1 sweep (precision value 2) on 1 M states, every flag combination."""
    desc = rd.panda() if robot == "panda" else rd.pr2_like()
    fd = scenarios.FIELD_A if robot == "panda" else scenarios.FIELD_DEFAULT
    e = EngineSetup(desc, fd, scenarios.random_scene(seed=17))
    q = desc.sample_states(1_000_000 if robot == "panda" else 200_000, 4242, np.float32)
    for flags in (CHECK_ROBOT, CHECK_INTRA, CHECK_ROBOT | CHECK_INTRA):
        a = e.group.check(q, e.world, flags=flags, precision=0)
        b = e.group.check(q, e.world, flags=flags, precision=2)
        assert np.array_equal(a, b), "flags=%d: %d differences" % (flags, int((a != b).sum()))
        assert a.sum() > 0


def test_host_buffer_entry_is_chunk_invariant(gpu, oracle):
    """b200df_check_batch pipelines 2^20-state chunks over two streams: ragged sizes around the chunk edge give the
    same verdicts as the device entry point on one launch."""
    import torch
    from moveit_b200 import binding as B
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    n = (1 << 21) + 12345
    q = desc.sample_states(n, 777, np.float32)
    host = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    qd = torch.from_numpy(q).cuda()
    vd = torch.empty(n, dtype=torch.uint8, device="cuda")
    e.group.check_device(qd.data_ptr(), B.Q_F32, n, vd.data_ptr(), e.world, None, CHECK_ROBOT | CHECK_INTRA,
                         B.PRECISION_EXACT, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(host, vd.cpu().numpy())
    for m in (1, 31, 33, (1 << 20) - 1, (1 << 20) + 1):
        assert np.array_equal(e.group.check(q[:m], e.world, flags=CHECK_ROBOT | CHECK_INTRA), host[:m])


def test_fast_precision_equals_exact_and_decides_most_states(gpu, oracle):
    """PRECISION_FAST = fp32 screening + exact re-run of the undecided states.  Its verdicts must be the exact
    path's bit for bit (2 M states, every flag combination); the raw screening pass (precision value 3) may only
    answer 0 / 1 where the exact kernel agrees, and must leave only a small fraction undecided."""
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    q = desc.sample_states(2_000_000, 9090, np.float32)
    for flags in (CHECK_ROBOT, CHECK_INTRA, CHECK_ROBOT | CHECK_INTRA):
        exact = e.group.check(q, e.world, flags=flags, precision=0)
        fast = e.group.check(q, e.world, flags=flags, precision=1)
        assert np.array_equal(exact, fast), "flags=%d: %d differences" % (flags, int((exact != fast).sum()))
        raw = e.group.check(q, e.world, flags=flags, precision=3)
        decided = raw != 2
        assert np.array_equal(raw[decided], exact[decided])
        undecided = 1.0 - decided.mean()
        print("flags=%d undecided fraction %.4f" % (flags, undecided))
        assert undecided < 0.05
    # out-of-range joint values leave the analysed range: screened as undecided, answered by the exact kernel
    q2 = q[:1000].copy()
    q2[::7, 0] = 500.0
    raw = e.group.check(q2, e.world, flags=CHECK_ROBOT, precision=3)
    assert (raw[::7] == 2).all()
    assert np.array_equal(e.group.check(q2, e.world, flags=CHECK_ROBOT, precision=1),
                          e.group.check(q2, e.world, flags=CHECK_ROBOT, precision=0))


def test_fast_precision_with_self_field_and_fine_field(gpu, oracle):
    """FAST with a self field (planning group) and on the 1 cm / u16 field (smaller voxels -> more face proximity)."""
    desc = rd.panda()
    objs = scenarios.random_scene(seed=77)
    q0 = np.zeros(desc.nvar)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, group="panda_arm", self_state=q0)
    sp = self_field_points(o.model_desc, desc, o.group_links, o.model.fk(q0))
    e = EngineSetup(desc, scenarios.FIELD_A, objs, group="panda_arm", self_points=sp)
    q = desc.sample_states(300_000, 31, np.float32)
    flags = CHECK_ROBOT | CHECK_SELF | CHECK_INTRA
    assert np.array_equal(e.group.check(q, e.world, e.self_field, flags=flags, precision=1),
                          e.group.check(q, e.world, e.self_field, flags=flags, precision=0))
    fine = dict(size=(2.0, 2.0, 2.0), resolution=0.01, center=(0.0, 0.0, 0.5), max_distance=0.25)
    e2 = EngineSetup(desc, fine, objs)
    assert e2.world.max_d2 == 625
    a = e2.group.check(q, e2.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=1)
    b = e2.group.check(q, e2.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=0)
    assert np.array_equal(a, b) and 0 < a.sum() < len(a)


def test_mobile_arm_every_joint_type_fast_exact_oracle(gpu, oracle):
    """Planar base, prismatic lift, revolute about -z / y / x / an oblique axis, a mimic joint with offset and a
    branch: the oracle, the exact kernel, the independent round-1 sweep and the FAST path must all agree."""
    desc = rd.mobile_arm()
    fd = scenarios.FIELD_DEFAULT
    objs = scenarios.random_scene(seed=57, n_boxes=3, n_cylinders=1)
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, fd, objs)
    q = desc.sample_states(200_000, 606, np.float32)
    flags = CHECK_ROBOT | CHECK_INTRA
    exact = e.group.check(q, e.world, flags=flags, precision=0)
    assert np.array_equal(exact[:30000], o.env.check_batch(q[:30000].astype(np.float64), mode=3, n_threads=8))
    assert np.array_equal(exact, e.group.check(q, e.world, flags=flags, precision=2))
    assert np.array_equal(exact, e.group.check(q, e.world, flags=flags, precision=1))
    raw = e.group.check(q, e.world, flags=flags, precision=3)  # the FAST tables do cover this robot
    assert np.array_equal(raw[raw != 2], exact[raw != 2]) and (raw == 2).mean() < 0.05
    assert 0 < exact.sum() < len(exact)
    T = e.model.fk(q[:2000].astype(np.float64))
    ref = np.stack([o.model.fk(x) for x in q[:2000].astype(np.float64)])
    assert np.max(np.abs(T - ref)) < 1e-12


def test_config4_pr2_like_fast_equals_exact(gpu, oracle):
    """The PR2-like tree (33 links, 124 derived spheres, ~500 enabled link pairs) also fits the FAST tables."""
    desc = rd.pr2_like()
    e = EngineSetup(desc, scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=31))
    q = desc.sample_states(200_000, 12, np.float32)
    for flags in (CHECK_ROBOT, CHECK_ROBOT | CHECK_INTRA):
        exact = e.group.check(q, e.world, flags=flags, precision=0)
        assert np.array_equal(exact, e.group.check(q, e.world, flags=flags, precision=1))
        raw = e.group.check(q, e.world, flags=flags, precision=3)
        assert np.array_equal(raw[raw != 2], exact[raw != 2])
        print("pr2_like flags=%d undecided %.4f colliding %.3f" % (flags, (raw == 2).mean(), exact.mean()))
        assert (raw == 2).mean() < 0.1
    # the self-collision part alone, where the verdicts of this tree are not a foregone conclusion: FAST runs the pair
    # tests one warp per state for a robot of this size (validity_fast_warp_pairs_kernel), on its own and behind the
    # field pass in an empty world; fp32 and fp64 rows; against the fp64 kernel on every state and the oracle on a slice
    o = OracleSetup(oracle, desc, scenarios.FIELD_DEFAULT, [], self_state=np.zeros(desc.nvar))
    empty = EngineSetup(desc, scenarios.FIELD_DEFAULT, [])
    exact = empty.group.check(q, None, flags=CHECK_INTRA, precision=0)
    assert 0.3 < exact.mean() < 0.95
    assert np.array_equal(exact, empty.group.check(q, None, flags=CHECK_INTRA, precision=1))
    raw = empty.group.check(q, None, flags=CHECK_INTRA, precision=3)
    assert np.array_equal(raw[raw != 2], exact[raw != 2]) and (raw == 2).mean() < 0.1
    both = empty.group.check(q, empty.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=1)
    assert np.array_equal(both, exact)  # nothing to hit in an empty world
    q64 = q[:70_000].astype(np.float64) + 1e-9
    f64 = empty.group.check(q64, None, flags=CHECK_INTRA, precision=1)
    assert np.array_equal(f64, empty.group.check(q64, None, flags=CHECK_INTRA, precision=0))
    assert np.array_equal(f64[:5000], o.env.check_batch(q64[:5000], mode=2, n_threads=8))


def _interpolate_reference(qa, qb, K, angular):
    """JointModel::interpolate in numpy doubles (revolute_joint_model.cpp:125-148, planar_joint_model.cpp:180-205)."""
    M, nv = qa.shape
    out = np.empty((M, K, nv))
    for j in range(1, K + 1):
        t = float(j) / float(K)
        diff = qb - qa
        lin = qa + diff * t
        d2 = np.where(diff > 0.0, 2.0 * np.pi - diff, -2.0 * np.pi - diff)
        r = qa - d2 * t
        r = np.where(r > np.pi, r - 2.0 * np.pi, np.where(r < -np.pi, r + 2.0 * np.pi, r))
        use_arc = (angular[None, :] == 1) & (np.abs(diff) > np.pi)
        out[:, j - 1, :] = np.where(use_arc, r, lin)
        for v in np.nonzero(angular == 2)[0]:  # FloatingJointModel::interpolate (floating_joint_model.cpp:137-157)
            f, g = qa[:, v:v + 4], qb[:, v:v + 4]
            dq = np.abs(((f[:, 0] * g[:, 0] + f[:, 1] * g[:, 1]) + f[:, 2] * g[:, 2]) + f[:, 3] * g[:, 3])
            theta = np.where(dq + np.finfo(np.float64).eps >= 1.0, 0.0, np.arccos(np.minimum(dq, 1.0)))
            big = theta > np.finfo(np.float64).eps
            with np.errstate(divide="ignore", invalid="ignore"):
                d = 1.0 / np.sin(theta)
                s0, s1 = np.sin((1.0 - t) * theta), np.sin(t * theta)
                sl = (f * s0[:, None] + g * s1[:, None]) * d[:, None]
            out[:, j - 1, v:v + 4] = np.where(big[:, None], sl, f)
    return out


@pytest.mark.parametrize("robot", ["panda", "mobile_arm"])
def test_batched_edge_check_matches_per_state_oracle(gpu, oracle, robot):
    """b200df_check_edges: motions interpolated on the device with the reference's JointModel::interpolate arithmetic,
    first colliding sample per motion == what checking the same interpolated states one by one with the oracle gives."""
    desc = rd.panda() if robot == "panda" else rd.mobile_arm()
    fd = scenarios.FIELD_A if robot == "panda" else scenarios.FIELD_DEFAULT
    objs = scenarios.random_scene() if robot == "panda" else scenarios.random_scene(seed=57, n_boxes=3, n_cylinders=1)
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, fd, objs)
    M, K = 3000, 12
    qa = desc.sample_states(M, 71)
    qb = desc.sample_states(M, 72)
    angular = np.zeros(desc.nvar, dtype=np.uint8)
    if robot == "mobile_arm":
        angular[2] = 1  # planar joint's theta takes the shortest arc
    flags = CHECK_ROBOT | CHECK_INTRA
    got = e.group.check_edges(qa, qb, K, e.world, flags=flags, precision=0, var_is_angular=angular)
    states = _interpolate_reference(qa, qb, K, angular)
    ref = o.env.check_batch(states.reshape(-1, desc.nvar), mode=3, n_threads=8).reshape(M, K)
    want = np.where(ref.any(axis=1), ref.argmax(axis=1) + 1, 0)
    assert np.array_equal(got, want)
    assert 0 < (got == 0).sum() < M
    fast = e.group.check_edges(qa, qb, K, e.world, flags=flags, precision=1, var_is_angular=angular)
    assert np.array_equal(fast, got)
    # enough states for the FAST screening pass to be taken (>= 64 k): it screens the fp64 states in fp32 and re-runs
    # the undecided ones unrounded
    qa2, qb2 = desc.sample_states(20_000, 73), desc.sample_states(20_000, 74)
    assert np.array_equal(e.group.check_edges(qa2, qb2, K, e.world, flags=flags, precision=1, var_is_angular=angular),
                          e.group.check_edges(qa2, qb2, K, e.world, flags=flags, precision=0, var_is_angular=angular))
    if robot == "mobile_arm":
        assert (np.abs(qb[:, 2] - qa[:, 2]) > np.pi).any()  # the shortest-arc branch was exercised


def test_device_resident_cloud_and_non_finite_joint_values(gpu, oracle):
    """b200df_field_add_points_device == add_points; NaN / inf joint values give the oracle's verdicts on every
    precision (a NaN centre fails every range test -> free, like the reference's dead out-of-bounds branch, Q4)."""
    import torch
    from moveit_b200 import engine
    fd = dict(size=(1.0, 1.0, 1.0), resolution=0.02, center=(0.0, 0.0, 0.5), max_distance=0.25)
    pts = scenarios.synthetic_cloud(20000, field=fd)
    a = engine.Field.centered(fd["size"], fd["resolution"], fd["center"], fd["max_distance"])
    b = engine.Field.centered(fd["size"], fd["resolution"], fd["center"], fd["max_distance"])
    a.add_points(pts)
    pd = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    b.add_points_device(pd.data_ptr(), len(pts), torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(a.d2(), b.d2()) and (a.d2() == 0).sum() > 1000
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    q = desc.sample_states(4096, 5, np.float32)
    q[::9, 3] = np.nan
    q[1::9, 0] = np.inf
    q[2::9, 6] = -np.inf
    q[3::9, 7] = np.nan  # prismatic finger
    ref = o.env.check_batch(q.astype(np.float64), mode=3, n_threads=4)
    for precision in (0, 1, 2):
        got = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=precision)
        assert np.array_equal(got, ref), "precision %d: %d differences" % (precision, int((got != ref).sum()))
    # fp64 joint values through FAST: screened rounded to fp32, undecided states re-run on the fp64 values
    q64 = desc.sample_states(300_000, 88, np.float64)
    a = e.group.check(q64, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=1)
    b = e.group.check(q64, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=0)
    raw = e.group.check(q64, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=3)
    assert np.array_equal(a, b) and np.array_equal(raw[raw != 2], b[raw != 2]) and (raw == 2).mean() < 0.01


def test_padding_sphere_known_answer_on_gpu(gpu, oracle):
    """test_collision_distance_field.cpp:336-383 through the engine: collide, collide, free."""
    from moveit_b200 import engine
    r = rd.RobotDescription("test")
    r.add_link("base", None, "root_joint", rd.JOINT_FIXED)
    r.add_shape("base", rd.SHAPE_SPHERE, (0.04,))
    md = engine.assemble_model(r, engine.body_decomposition, 0.02, link_padding=0.04)
    assert md["spheres"][0][0, 3] == 0.08
    model = engine.Model(md)
    grp = engine.Group(model, [0], [1], [[1]], 0.02, 0.0, 0.25)
    world = engine.Field.centered((0.1, 0.1, 0.2), 0.02, (0, 0, 0), 0.25)
    assert world.dims == (5, 5, 10)
    q = np.zeros((1, 0))
    box = (rd.SHAPE_BOX, (0.02, 0.02, 0.02))
    expected = [(0.0, 1), (0.08, 1), (0.1, 0)]
    for z, verdict in expected:
        world.reset()  # MOVE_SHAPE: old points removed, new points added
        world.add_posed_body(box[0], box[1], rd.rpy_xyz_to_pose((0, 0, z)))
        assert grp.check(q, world, flags=CHECK_ROBOT)[0] == verdict, z


def test_ragged_and_empty_batches(gpu, oracle):
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    assert len(e.group.check(np.zeros((0, desc.nvar)), e.world)) == 0
    for n in (1, 31, 33, 127, 129, 1000):
        q = desc.sample_states(n, 1000 + n)
        assert np.array_equal(e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA),
                              o.env.check_batch(q, mode=3))
    # states far outside the field are "free" (Q4), NaN joints do not crash
    q = desc.sample_states(4, 1)
    q[0, :] = np.nan
    e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA)


# ------------------------------------------------------------------------------------------------ gradients
def test_config5_gradients_within_tolerance(gpu, oracle):
    """BASELINE config 5 at oracle-sized N: getCollisionGradients (self field + intra + environment)."""
    desc = rd.panda()
    objs = scenarios.random_scene()
    q0 = np.zeros(desc.nvar)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=q0)
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    traj = scenarios.chomp_trajectories(desc, 12, 100).reshape(-1, desc.nvar)
    got = e.group.gradients(traj, e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    ref = o.env.gradients_batch(traj, flags=2 | 4)
    assert np.array_equal(got["type"], ref["type"].astype(np.uint8))
    fin = ref["dist"] < 1e300
    assert np.array_equal(fin, got["dist"] < 1e300)
    assert np.max(np.abs(got["dist"][fin] - ref["dist"][fin])) < 1e-5   # north_star tolerance: 1e-5 m
    assert np.max(np.abs(got["grad"] - ref["grad"])) < 1e-5
    assert np.max(np.abs(got["centers"] - ref["centers"])) < 1e-12
    assert (got["type"] == 3).any() and (got["type"] == 2).any()


def test_config5_full_size_1000x100_with_link_fields_and_f32_outputs(gpu, oracle):
    """BASELINE config 5 at its full size: 1 000 trajectories x 100 waypoints, getCollisionGradients with the per-link
    fields, environment and intra-group terms against the oracle (types exact, values within north_star's 1e-5 m),
    through the fp64-output call and the single-precision-output call."""
    from moveit_b200 import engine
    from moveit_b200.binding import CHECK_LINK_FIELDS
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    o.env.build_link_fields()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    fields, enabled = engine.link_distance_fields(e.model_desc, e.group_links, desc.acm_entry_matrix(),
                                                  scenarios.FIELD_A["resolution"], scenarios.FIELD_A["max_distance"])
    e.group.set_link_fields(fields, enabled)
    traj = scenarios.chomp_trajectories(desc, 1000, 100).reshape(-1, desc.nvar)
    assert len(traj) == 100_000
    fl = CHECK_ROBOT | CHECK_INTRA | CHECK_LINK_FIELDS
    got = e.group.gradients(traj, e.world, flags=fl)
    ref = o.env.gradients_batch(traj, flags=7)
    typ = ref["type"].astype(np.uint8)
    assert np.array_equal(got["type"], typ)
    fin = ref["dist"] < 1e300
    assert np.array_equal(fin, got["dist"] < 1e300)
    assert np.max(np.abs(got["dist"][fin] - ref["dist"][fin])) < 1e-5   # north_star tolerance: 1e-5 m
    assert np.max(np.abs(got["grad"] - ref["grad"])) < 1e-5
    assert np.max(np.abs(got["centers"] - ref["centers"])) < 1e-12
    assert all((typ == t).any() for t in (0, 1, 2, 3))
    g32 = e.group.gradients(traj, e.world, flags=fl, out_dtype=np.float32)
    assert g32["dist"].dtype == np.float32 and np.array_equal(g32["type"], typ)
    assert np.array_equal(fin, g32["dist"] < 1e30) and np.all(g32["dist"][~fin] == np.finfo(np.float32).max)
    assert np.max(np.abs(g32["dist"][fin] - ref["dist"][fin])) < 1e-5
    assert np.max(np.abs(g32["grad"] - ref["grad"])) < 1e-5
    assert np.max(np.abs(g32["centers"] - ref["centers"])) < 1e-5
    # the rounded values are exactly the fp64 results rounded once
    assert np.array_equal(g32["grad"], got["grad"].astype(np.float32))
    assert np.array_equal(g32["dist"][fin], got["dist"][fin].astype(np.float32))


def test_device_resident_gradient_entry_points_match_the_host_buffer_ones(gpu):
    """b200df_gradients_batch_device / _device_f32 (results stay on the GPU, nothing synchronises) against the host-buffer
    calls on the same states, for a small batch (CTA-per-state variant) and a large one, into pageable arrays (staged
    by the library) and page-locked ones."""
    import ctypes as C
    import torch
    from moveit_b200 import binding as B
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    fl = CHECK_ROBOT | CHECK_INTRA
    for n in (100, 40_000):
        q = scenarios.chomp_trajectories(desc, n // 100, 100).reshape(-1, desc.nvar)
        S = int(e.group.n_spheres)
        qd = torch.from_numpy(q).cuda()
        st = torch.cuda.current_stream().cuda_stream
        for dt, fn, kw in ((torch.float64, B.load().b200df_gradients_batch_device, {}),
                           (torch.float32, B.load().b200df_gradients_batch_device_f32, dict(out_dtype=np.float32))):
            dd = torch.empty((n, S), dtype=dt, device="cuda")
            gd = torch.empty((n, S, 3), dtype=dt, device="cuda")
            td = torch.empty((n, S), dtype=torch.uint8, device="cuda")
            cd = torch.empty((n, S, 3), dtype=dt, device="cuda")
            B.check(fn(e.group.h, e.world.h, None, C.c_void_p(qd.data_ptr()), B.Q_F64, n, fl, C.c_void_p(dd.data_ptr()),
                       C.c_void_p(gd.data_ptr()), C.c_void_p(td.data_ptr()), C.c_void_p(cd.data_ptr()), C.c_void_p(st)))
            torch.cuda.synchronize()
            host = e.group.gradients(q, e.world, flags=fl, **kw)  # pageable outputs
            pin = {k: torch.empty(v.shape, dtype=torch.from_numpy(v).dtype).pin_memory().numpy() for k, v in host.items()}
            e.group.gradients(q, e.world, flags=fl, out=pin, **kw)
            for got in (host, pin):
                assert np.array_equal(got["dist"], dd.cpu().numpy()) and np.array_equal(got["grad"], gd.cpu().numpy())
                assert np.array_equal(got["type"], td.cpu().numpy()) and np.array_equal(got["centers"], cd.cpu().numpy())


def test_gradients_with_link_distance_fields(gpu, oracle):
    """getSelfProximityGradients' per-link PosedDistanceField term (:395-418), including the reference's quirks: the
    gradient is mapped with the full pose (rotation + translation, Q6) and a link's sphere loop stops at the first
    sphere outside the other link's field (Q7)."""
    from moveit_b200 import engine
    from moveit_b200.binding import CHECK_LINK_FIELDS
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    o.env.build_link_fields()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    fields, enabled = engine.link_distance_fields(e.model_desc, e.group_links, desc.acm_entry_matrix(),
                                                  scenarios.FIELD_A["resolution"], scenarios.FIELD_A["max_distance"])
    e.group.set_link_fields(fields, enabled)
    traj = scenarios.chomp_trajectories(desc, 30, 100).reshape(-1, desc.nvar)
    got = e.group.gradients(traj, e.world, flags=CHECK_ROBOT | CHECK_INTRA | CHECK_LINK_FIELDS)
    ref = o.env.gradients_batch(traj, flags=7)
    assert np.array_equal(got["type"], ref["type"].astype(np.uint8))
    fin = ref["dist"] < 1e300
    assert np.array_equal(fin, got["dist"] < 1e300)
    assert np.max(np.abs(got["dist"][fin] - ref["dist"][fin])) < 1e-5
    assert np.max(np.abs(got["grad"] - ref["grad"])) < 1e-5
    # the term does change results: some spheres now carry a SELF entry that the run without link fields lacks
    plain = e.group.gradients(traj, e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    assert (got["type"] == 1).sum() > 0 and not np.array_equal(plain["type"], got["type"])


def test_distance_batch_definition(gpu, oracle):
    """distanceRobot is a stub in the reference (F5); the engine defines min over spheres of dist - radius."""
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs)
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    q = desc.sample_states(200, 17)
    got = e.group.distance(q, e.world)
    radii = np.concatenate([s[:, 3] for s in o.model_desc["spheres"] if len(s)])
    for k in range(0, 200, 9):
        c = o.env.sphere_centers(q[k])
        d = np.array([o.env.world.distance_gradient(*p)[0] for p in c]) - radii
        assert abs(got[k] - d.min()) < 1e-12


# ------------------------------------------------------------------------------------------------ full sizes
def test_full_size_properties_10m_states(gpu, oracle):
    """Config 2 at BASELINE size through size-independent properties: chunk invariance (the verdict of a state
    does not depend on its position in the batch), robot|intra == robot OR intra, and exact agreement with the
    oracle on a strided sample of the same batch."""
    desc = rd.panda()
    objs = scenarios.random_scene()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    n = 10_000_000
    q = desc.sample_states(n, scenarios.SEED_STATES, np.float32)
    both = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    robot = e.group.check(q, e.world, flags=CHECK_ROBOT)
    intra = e.group.check(q, flags=CHECK_INTRA)
    assert np.array_equal(both, robot | intra)
    sl = slice(3_333_333, 3_343_333)
    assert np.array_equal(e.group.check(q[sl], e.world, flags=CHECK_ROBOT | CHECK_INTRA), both[sl])
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    idx = np.arange(0, n, 47)  # 212 766 states through the CPU oracle
    ref = o.env.check_batch(q[idx].astype(np.float64), mode=3, n_threads=8)
    assert np.array_equal(both[idx], ref)
    fast = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=1)
    assert np.array_equal(fast, both)


def test_config3_field_b_cloud(gpu, oracle):
    """Config 3: 1M-point cloud into the 400^3 field; exactness checked against a KD-tree on 300 000 sampled voxels.
    (The mismatch census against the reference's propagation on this field is a 30 s CPU job: 30 761 of 64 M voxels,
    always reference > exact, at most 8 in d^2 - DESIGN.md F4 table.)"""
    from scipy.spatial import cKDTree
    from moveit_b200 import engine
    fb = scenarios.FIELD_B
    pts = scenarios.synthetic_cloud(1_000_000)
    g = engine.Field.centered(fb["size"], fb["resolution"], fb["center"], fb["max_distance"])
    assert g.dims == (400, 400, 400) and g.max_d2 == 625
    g.add_points(pts)
    ms = g.rebuild()
    d2 = g.d2()
    occ = g.occupancy()
    assert (d2 == 0).sum() == occ.sum() > 100_000
    assert np.array_equal(d2 == 0, occ == 1)
    vox = np.argwhere(occ)
    tree = cKDTree(vox)
    rng = np.random.default_rng(0)
    sample = rng.integers(0, 400, size=(300_000, 3))
    dist, _ = tree.query(sample, k=1)
    want = np.minimum(np.rint(dist * dist).astype(np.int64), 625)
    assert np.array_equal(d2[sample[:, 0], sample[:, 1], sample[:, 2]], want)
    print("field B rebuild: %.3f ms" % ms)


def test_concurrent_callers_share_one_group(gpu, oracle):
    """planning_scene/test/test_multi_threaded.cpp, batched: several host threads call the C ABI on the same model,
    group and field at once (verdicts in both precisions, gradients, edges) and every one must get the answer a lone
    caller gets."""
    import threading
    desc = rd.panda()
    objs = scenarios.random_scene()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    flags = CHECK_ROBOT | CHECK_INTRA
    n_threads, rounds = 6, 12
    batches = [desc.sample_states(20_000 + 3_000 * t, 900 + t, np.float32) for t in range(n_threads)]
    want = [e.group.check(b, e.world, flags=flags, precision=0) for b in batches]
    want_grad = [e.group.gradients(b[:500], e.world) for b in batches]
    want_edge = [e.group.check_edges(b[:2000].astype(np.float64), b[2000:4000].astype(np.float64), 8, e.world,
                                     flags=flags) for b in batches]
    errors = []

    def worker(t):
        try:
            for k in range(rounds):
                got = e.group.check(batches[t], e.world, flags=flags, precision=k % 2)
                if not np.array_equal(got, want[t]):
                    errors.append("thread %d round %d: %d verdicts differ" % (t, k, int((got != want[t]).sum())))
                if k % 4 == 0:
                    g = e.group.gradients(batches[t][:500], e.world)
                    for key in ("dist", "grad", "type", "centers"):
                        if not np.array_equal(g[key], want_grad[t][key]):
                            errors.append("thread %d round %d: gradients[%s] differ" % (t, k, key))
                            break
                if k % 4 == 2:
                    ed = e.group.check_edges(batches[t][:2000].astype(np.float64),
                                             batches[t][2000:4000].astype(np.float64), 8, e.world, flags=flags)
                    if not np.array_equal(ed, want_edge[t]):
                        errors.append("thread %d round %d: edge results differ" % (t, k))
        except Exception as ex:  # noqa: BLE001 - reported through the assertion below
            errors.append("thread %d: %r" % (t, ex))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:5]


# ------------------------------------------------------------------------------- small-batch (one CTA per state) kernel
PREC_SMALL = 4  # B200DF_PRECISION_DEBUG_SMALL: the kernel EXACT / FAST pick for small batches, forced at any batch size


@pytest.mark.parametrize("robot", ["panda", "mobile_arm", "pr2_like"])
def test_small_batch_kernel_equals_batched_exact_and_oracle(gpu, oracle, robot):
    """validity_small.cu spreads one state over a CTA; the transforms and predicates are the same arithmetic, so the
    verdicts equal the one-thread-per-state kernel's (and the oracle's) on every robot, flag set and joint dtype."""
    desc = {"panda": rd.panda, "mobile_arm": rd.mobile_arm, "pr2_like": rd.pr2_like}[robot]()
    fd = scenarios.FIELD_A if robot == "panda" else scenarios.FIELD_DEFAULT
    objs = scenarios.random_scene() if robot == "panda" else scenarios.random_scene(seed=31)
    e = EngineSetup(desc, fd, objs)
    n = 60_000 if robot != "pr2_like" else 30_000
    for dtype in (np.float32, np.float64):
        q = desc.sample_states(n, 17, dtype)
        for flags in (CHECK_ROBOT, CHECK_INTRA, CHECK_ROBOT | CHECK_INTRA):
            batched = e.group.check(q, e.world, flags=flags, precision=0)  # n > small-batch limit: batched kernel
            small = e.group.check(q, e.world, flags=flags, precision=PREC_SMALL)
            assert np.array_equal(small, batched), (robot, dtype, flags, int((small != batched).sum()))
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    q = desc.sample_states(3000, 18)
    got = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=PREC_SMALL)
    assert np.array_equal(got, o.env.check_batch(q, mode=3, n_threads=8))


def test_small_batch_kernel_self_field_signed_fine_and_non_finite(gpu, oracle):
    desc = rd.panda()
    objs = scenarios.random_scene(seed=77)
    q0 = np.zeros(desc.nvar)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, group="panda_arm", self_state=q0)
    sp = self_field_points(o.model_desc, desc, o.group_links, o.model.fk(q0))
    e = EngineSetup(desc, scenarios.FIELD_A, objs, group="panda_arm", self_points=sp)
    q = desc.sample_states(20_000, 99)
    flags = CHECK_ROBOT | CHECK_SELF | CHECK_INTRA
    got = e.group.check(q, e.world, e.self_field, flags=flags, precision=PREC_SMALL)
    assert np.array_equal(got, e.group.check(q, e.world, e.self_field, flags=flags, precision=0))
    assert np.array_equal(got[:4000], o.env.check_batch(q[:4000], mode=3, n_threads=8))
    # signed world field (fp64 predicate instead of the integer threshold)
    es = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene(seed=5), signed=True)
    assert np.array_equal(es.group.check(q, es.world, flags=CHECK_ROBOT, precision=PREC_SMALL),
                          es.group.check(q, es.world, flags=CHECK_ROBOT, precision=0))
    # 1 cm / u16 field
    fine = dict(size=(2.0, 2.0, 2.0), resolution=0.01, center=(0.0, 0.0, 0.5), max_distance=0.25)
    ef = EngineSetup(desc, fine, objs)
    a = ef.group.check(q, ef.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=PREC_SMALL)
    assert np.array_equal(a, ef.group.check(q, ef.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=0))
    assert 0 < a.sum() < len(a)
    # NaN / inf joint values
    qn = desc.sample_states(12_000, 5, np.float32)
    qn[::9, 3] = np.nan
    qn[1::9, 0] = np.inf
    qn[2::9, 6] = -np.inf
    qn[3::9, 7] = np.nan
    ew = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    assert np.array_equal(ew.group.check(qn, ew.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=PREC_SMALL),
                          ew.group.check(qn, ew.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=0))


def test_small_batches_take_the_low_latency_path_and_agree(gpu, oracle, monkeypatch):
    """n = 1 ... a few thousand states: EXACT and FAST route to the small-batch kernel (host entry: through one mapped
    pinned buffer).  Same verdicts with the route switched off, and the oracle's."""
    import torch
    from moveit_b200.binding import Q_F32
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    flags = CHECK_ROBOT | CHECK_INTRA
    q = desc.sample_states(6000, 41, np.float32)
    ref = o.env.check_batch(q.astype(np.float64), mode=3, n_threads=8)
    for n in (1, 2, 31, 32, 33, 500, 2048, 2049, 6000):
        for precision in (0, 1):
            assert np.array_equal(e.group.check(q[:n], e.world, flags=flags, precision=precision), ref[:n]), (n, precision)
        assert np.array_equal(e.group.check(q[:n].astype(np.float64), e.world, flags=flags), ref[:n])
    # one state at a time, the way isValid() is called
    single = np.array([e.group.check(q[i:i + 1], e.world, flags=flags)[0] for i in range(300)], dtype=np.uint8)
    assert np.array_equal(single, ref[:300])
    # device-resident entry
    qd = torch.from_numpy(q).cuda()
    vd = torch.empty(len(q), dtype=torch.uint8, device="cuda")
    e.group.check_device(qd.data_ptr(), Q_F32, 777, vd.data_ptr(), e.world, flags=flags, precision=1,
                         stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(vd[:777].cpu().numpy(), ref[:777])
    # the batched kernels still serve small batches when the small-batch route is off
    monkeypatch.setenv("B200DF_SMALL_MAX_STATES", "0")
    monkeypatch.setenv("B200DF_SMALL_BATCH", "0")
    for n in (1, 33, 2048):
        for precision in (0, 1):
            assert np.array_equal(e.group.check(q[:n], e.world, flags=flags, precision=precision), ref[:n])


def test_gradient_call_sizes_and_pinned_outputs_agree(gpu, oracle):
    """One CHOMP iteration hands over one trajectory (~100 waypoints): such calls take the warp-per-state FK variant
    of the gradient kernel and the mapped pinned staging; results must be bit-identical to one big call (32 states
    per CTA, staged copies), with pageable or pinned result arrays, with and without the per-link fields."""
    import torch
    from moveit_b200 import engine
    from moveit_b200.binding import CHECK_LINK_FIELDS
    desc = rd.panda()
    objs = scenarios.random_scene()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    fields, enabled = engine.link_distance_fields(e.model_desc, e.group_links, desc.acm_entry_matrix(),
                                                  scenarios.FIELD_A["resolution"], scenarios.FIELD_A["max_distance"])
    e.group.set_link_fields(fields, enabled)
    traj = scenarios.chomp_trajectories(desc, 60, 100).reshape(-1, desc.nvar)  # 6000 states: the big-call variant
    S = e.group.n_spheres
    for flags in (CHECK_ROBOT | CHECK_INTRA, CHECK_ROBOT | CHECK_INTRA | CHECK_LINK_FIELDS):
        big = e.group.gradients(traj, e.world, flags=flags)
        for lo, n in ((0, 100), (100, 100), (1234, 1), (2000, 37), (3000, 1500)):
            part = e.group.gradients(traj[lo:lo + n], e.world, flags=flags)
            for key in ("dist", "grad", "type", "centers"):
                assert np.array_equal(part[key], big[key][lo:lo + n]), (flags, lo, n, key)
        n = 100
        pin = dict(dist=torch.empty((n, S), dtype=torch.float64).pin_memory().numpy(),
                   grad=torch.empty((n, S, 3), dtype=torch.float64).pin_memory().numpy(),
                   type=torch.empty((n, S), dtype=torch.uint8).pin_memory().numpy(),
                   centers=torch.empty((n, S, 3), dtype=torch.float64).pin_memory().numpy())
        got = e.group.gradients(traj[500:600], e.world, flags=flags, out=pin)
        for key in ("dist", "grad", "type", "centers"):
            assert np.array_equal(got[key], big[key][500:600]), (flags, "pinned", key)
    # sphere centres / clearance / FK through the same small-call staging
    c_big = e.group.sphere_centers(traj)
    assert np.array_equal(e.group.sphere_centers(traj[77:78]), c_big[77:78])
    d_big = e.group.distance(traj, e.world)
    assert np.array_equal(e.group.distance(traj[300:420], e.world), d_big[300:420])


def test_robot_too_large_for_the_batched_kernels_runs_on_the_cta_per_state_kernel(gpu, oracle):
    """378 collision spheres on the PR2-like tree: a state's shared-memory column of the one-thread-per-state kernels
    no longer fits an SM (and FAST's tables overflow), so every batch size is served by the one-CTA-per-state
    kernel.  Verdicts against the oracle."""
    desc = rd.pr2_like()
    n_sph = 0
    for li, (name, shapes) in enumerate(zip(desc.link_names, desc.shapes)):
        if not shapes:
            continue
        rng = np.random.default_rng(1000 + li)  # irregular positions: a regular lattice makes exact distance ties
        desc.set_spheres(name, [(rng.uniform(-0.2, 0.2), rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03),
                                 rng.uniform(0.03, 0.045)) for _ in range(14)])
        n_sph += 14
    assert n_sph > 300
    fd = scenarios.FIELD_DEFAULT
    objs = scenarios.random_scene(seed=31)
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, fd, objs)
    assert e.group.n_spheres == n_sph
    q = desc.sample_states(40_000, 77, np.float32)  # above the small-batch limit: the fallback, not the routing
    flags = CHECK_ROBOT | CHECK_INTRA
    got = e.group.check(q, e.world, flags=flags, precision=0)
    assert np.array_equal(got, e.group.check(q, e.world, flags=flags, precision=1))
    ref = o.env.check_batch(q[:3000].astype(np.float64), mode=3, n_threads=8)
    assert np.array_equal(got[:3000], ref)
    robot_only = e.group.check(q, e.world, flags=CHECK_ROBOT, precision=0)
    assert np.array_equal(robot_only[:3000], o.env.check_batch(q[:3000].astype(np.float64), mode=1, n_threads=8))
    assert 0 < robot_only.sum() < len(q)
    # the gradient kernel likewise: its 32-state tile does not fit, the 4-state variant serves any batch size
    qg = desc.sample_states(5000, 78)
    got = e.group.gradients(qg, e.world, flags=flags)
    ref = o.env.gradients_batch(qg[:300], flags=2 | 4)
    assert np.array_equal(got["type"][:300], ref["type"].astype(np.uint8))
    fin = ref["dist"] < 1e300
    assert np.max(np.abs(got["dist"][:300][fin] - ref["dist"][fin])) < 1e-5
    assert np.max(np.abs(got["grad"][:300] - ref["grad"])) < 1e-5


def test_c_abi_error_behaviour(gpu, oracle):
    """The reference never throws across its API; the C ABI likewise answers every misuse with a status + message and
    stays usable: null pointers, unknown dtype / precision, a flag without its field, fields that do not fit the
    group, the screening-only hook on a call it cannot serve."""
    import ctypes as C
    from moveit_b200 import binding as B
    from moveit_b200 import engine
    lib = B.load()
    desc = rd.panda()
    objs = scenarios.random_scene()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    q = desc.sample_states(64, 3, np.float32)
    v = np.zeros(64, dtype=np.uint8)
    qp, vp = q.ctypes.data_as(C.c_void_p), B.ptr(v, C.c_uint8)
    good = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA)

    def expect(code, rc):
        assert rc == code, (rc, lib.b200df_last_error())
        assert len(lib.b200df_last_error()) > 0

    expect(B.ERR_INVALID, lib.b200df_check_batch(None, e.world.h, None, qp, B.Q_F32, 64, CHECK_ROBOT, 0, vp))
    expect(B.ERR_INVALID, lib.b200df_check_batch(e.group.h, e.world.h, None, None, B.Q_F32, 64, CHECK_ROBOT, 0, vp))
    expect(B.ERR_INVALID, lib.b200df_check_batch(e.group.h, e.world.h, None, qp, 7, 64, CHECK_ROBOT, 0, vp))
    expect(B.ERR_INVALID, lib.b200df_check_batch(e.group.h, e.world.h, None, qp, B.Q_F32, 64, CHECK_ROBOT, 99, vp))
    expect(B.ERR_INVALID, lib.b200df_check_batch(e.group.h, e.world.h, None, qp, B.Q_F32, -1, CHECK_ROBOT, 0, vp))
    expect(B.ERR_INVALID, lib.b200df_check_batch(e.group.h, None, None, qp, B.Q_F32, 64, CHECK_ROBOT, 0, vp))  # no world
    coarse = engine.Field.centered((1.0, 1.0, 1.0), 0.05, (0.0, 0.0, 0.5), 0.25)  # other resolution than the group's
    expect(B.ERR_INVALID, lib.b200df_check_batch(e.group.h, coarse.h, None, qp, B.Q_F32, 64, CHECK_ROBOT, 0, vp))
    signed = engine.Field.centered(scenarios.FIELD_A["size"], 0.02, scenarios.FIELD_A["center"], 0.25, True)
    q64 = q.astype(np.float64)
    # screening hook on a call it cannot serve: a signed field whose negative distances matter (tolerance >= radius)
    et = EngineSetup(desc, scenarios.FIELD_A, objs, tol=1.0)
    expect(B.ERR_UNSUPPORTED, lib.b200df_check_batch(et.group.h, signed.h, None, q64.ctypes.data_as(C.c_void_p), B.Q_F64,
                                                     64, CHECK_ROBOT, 3, vp))
    expect(B.ERR_INVALID, lib.b200df_field_add_points(None, None, 5))
    expect(B.ERR_INVALID, lib.b200df_gradients_batch(e.group.h, e.world.h, None, qp, B.Q_F32, 4, CHECK_ROBOT, None,
                                                     None, None, None))
    assert lib.b200df_check_batch(e.group.h, e.world.h, None, qp, B.Q_F32, 0, CHECK_ROBOT, 0, None) == B.OK  # n = 0
    # none of the above left the engine in a bad state
    assert np.array_equal(e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA), good)
    with pytest.raises(B.B200dfError):
        e.group.check(q, None, flags=CHECK_ROBOT)


# ----------------------------------------------------------------------------------------------- random robots (fuzz)
from common import FUZZ_CASES as _FUZZ, random_robot as _random_robot  # noqa: E402


@pytest.mark.parametrize("seed,n_links,floating,kw", _FUZZ)
def test_random_robots_every_kernel_against_the_oracle(gpu, oracle, seed, n_links, floating, kw):
    """Differential fuzzing: random trees (up to 40 links with geometry and a full ACM -> more than 32 partners per
    link, i.e. two partner chunks) through the oracle, the batched exact kernel, the debug sweep, the
    one-CTA-per-state kernel and FAST (where the robot qualifies), for every flag set."""
    from moveit_b200 import binding as B
    desc = _random_robot(100 + seed, n_links, floating, **kw)
    fd = scenarios.FIELD_DEFAULT
    objs = scenarios.random_scene(seed=200 + seed, n_boxes=6, n_cylinders=3)
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, fd, objs)
    q = desc.sample_states(70_000, 300 + seed)
    n_ref = 2500
    for flags, mode in ((CHECK_ROBOT, 1), (CHECK_INTRA, 2), (CHECK_ROBOT | CHECK_INTRA, 3)):
        ref = o.env.check_batch(q[:n_ref], mode=mode, n_threads=8)
        exact = e.group.check(q, e.world, flags=flags, precision=0)  # 70 k states: the batched kernel
        assert np.array_equal(exact[:n_ref], ref), (seed, flags, int((exact[:n_ref] != ref).sum()))
        assert 0 < exact.sum() < len(exact)  # both outcomes occur
        for precision in (1, 2, 4):
            try:
                got = e.group.check(q, e.world, flags=flags, precision=precision)
            except B.B200dfError as ex:  # the debug sweep keeps every centre in shared memory: big robots do not fit
                assert precision == 2 and ex.code == B.ERR_UNSUPPORTED
                continue
            assert np.array_equal(got, exact), (seed, flags, precision, int((got != exact).sum()))
        q32 = q.astype(np.float32)
        e32 = e.group.check(q32, e.world, flags=flags, precision=0)
        assert np.array_equal(e.group.check(q32, e.world, flags=flags, precision=1), e32)
        assert np.array_equal(e.group.check(q32, e.world, flags=flags, precision=4), e32)
        try:  # the screening pass alone, where the robot fits FAST's tables
            raw = e.group.check(q32, e.world, flags=flags, precision=3)
            assert np.array_equal(raw[raw != 2], e32[raw != 2]) and (raw == 2).mean() < 0.2
        except B.B200dfError as ex:
            assert ex.code == B.ERR_UNSUPPORTED and seed not in (5, 7)  # those two are built to fit FAST's tables
    if seed in (5, 7):  # a floating-root tree / a tree with two partner chunks: FAST at a size where margins get hit
        big = desc.sample_states(1_000_000, 400 + seed, np.float32)
        fl = CHECK_ROBOT | CHECK_INTRA
        ex_big = e.group.check(big, e.world, flags=fl, precision=0)
        assert np.array_equal(e.group.check(big, e.world, flags=fl, precision=1), ex_big)
        raw = e.group.check(big, e.world, flags=fl, precision=3)
        assert np.array_equal(raw[raw != 2], ex_big[raw != 2])
        print("fuzz %d: FAST undecided %.4f" % (seed, (raw == 2).mean()))
        assert (raw == 2).mean() < 0.2
    # gradients (random geometry: no exact distance ties)
    got = e.group.gradients(q[:300], e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    refg = o.env.gradients_batch(q[:300], flags=2 | 4)
    assert np.array_equal(got["type"], refg["type"].astype(np.uint8))
    fin = refg["dist"] < 1e300
    assert np.max(np.abs(got["dist"][fin] - refg["dist"][fin])) < 1e-5
    assert np.max(np.abs(got["grad"] - refg["grad"])) < 1e-5
    # ... and with the per-link PosedDistanceFields of getSelfProximityGradients (Q6 / Q7), which take at most 32 links
    if len(desc.links_with_geometry()) <= 32:
        from moveit_b200 import engine
        from moveit_b200.binding import CHECK_LINK_FIELDS
        o.env.build_link_fields()
        fields, enabled = engine.link_distance_fields(e.model_desc, e.group_links, desc.acm_entry_matrix(),
                                                      fd["resolution"], fd["max_distance"])
        e.group.set_link_fields(fields, enabled)
        got = e.group.gradients(q[:300], e.world, flags=CHECK_ROBOT | CHECK_INTRA | CHECK_LINK_FIELDS)
        refg = o.env.gradients_batch(q[:300], flags=7)
        assert np.array_equal(got["type"], refg["type"].astype(np.uint8))
        fin = refg["dist"] < 1e300
        assert np.array_equal(fin, got["dist"] < 1e300)
        assert np.max(np.abs(got["dist"][fin] - refg["dist"][fin])) < 1e-5
        assert np.max(np.abs(got["grad"] - refg["grad"])) < 1e-5


# radius = ceil(max_distance / resolution); the EDT dispatch depends on it and on nz:
#   nz odd -> windowed fallback kernels; nz % 8 != 0 -> pair-wise z pass; nz % 16 != 0 -> word-wise bit packing;
#   radius <= 13 / 25 / 31 -> marching instantiations 12 / 24 / 30; radius > 31 -> windowed fallback
_EDT_CASES = [
    ((31, 29, 27), 0.02, 0.10, False),   # nz odd, radius 5, u8
    ((20, 22, 30), 0.02, 0.26, False),   # nz even, not % 8, radius 13 (march<12>), u8 (169)
    ((18, 20, 40), 0.02, 0.27, False),   # nz % 8 == 0, not % 16, radius 14 -> march<24>, u8 (196)
    ((24, 18, 48), 0.01, 0.25, False),   # nz % 16 == 0, radius 25 (march<24>), u16 (625)
    ((30, 26, 32), 0.01, 0.255, False),  # radius 26 -> march<30>, u16
    ((40, 36, 64), 0.01, 0.31, False),   # radius 31 (march<30>), u16 (961)
    ((40, 36, 64), 0.01, 0.33, False),   # radius 33 -> windowed fallback
    ((22, 24, 32), 0.02, 0.26, True),    # signed: the negative field runs the same passes on the complement
    ((17, 19, 26), 0.05, 0.02, False),   # radius 1: max_d2 = 1
    ((9, 70, 8), 0.02, 0.25, False),     # a thin grid
]


@pytest.mark.parametrize("shape,res,max_dist,signed", _EDT_CASES)
def test_edt_dispatch_variants_bit_exact(gpu, oracle, shape, res, max_dist, signed):
    """Every branch of the EDT dispatch against brute force (sparse random voxels plus a solid blob, so that both the
    INF-skip and the dense paths run), after an incremental sequence of edits as well."""
    from moveit_b200 import engine
    size = [s * res + 1e-9 for s in shape]
    g = engine.Field(size, res, (0, 0, 0), max_dist, signed)
    assert g.dims == shape
    rng = np.random.default_rng(sum(shape))
    n_pts = max(4, int(np.prod(shape)) // 400)
    ijk = np.stack([rng.integers(0, s, n_pts) for s in shape], axis=1)
    c = [s // 2 for s in shape]
    blob = np.array([(c[0] + a, c[1] + b, c[2] + d) for a in range(-2, 3) for b in range(-2, 3) for d in range(-2, 3)
                     if 0 <= c[0] + a < shape[0] and 0 <= c[1] + b < shape[1] and 0 <= c[2] + d < shape[2]])
    occ = np.zeros(shape, dtype=np.uint8)
    for batch in (ijk[: n_pts // 2], blob, ijk[n_pts // 2:]):  # a rebuild after every edit
        g.add_voxels(batch)
        occ[batch[:, 0], batch[:, 1], batch[:, 2]] = 1
        assert np.array_equal(g.occupancy(), occ)
        want = oracle.edt_exact(occ, g.max_d2, brute=max(shape) <= 48)
        assert np.array_equal(g.d2(), want), shape
        if signed:
            assert np.array_equal(g.d2(True), oracle.edt_exact(1 - occ, g.max_d2, brute=max(shape) <= 48))
    g.reset()
    assert (g.d2() == g.max_d2).all()


@pytest.mark.parametrize("seed,signed", [(1, False), (2, False), (3, True)])
def test_random_field_edit_sequences(gpu, oracle, seed, signed):
    """Forty random edits (points in and out of bounds, removals of present and absent points, updatePointsInField,
    shapes added / removed / moved with random poses, octree leaves, voxel lists, resets): after every edit the
    occupancy equals the oracle's zero set and d^2 (and the negative field) equal the exact EDT of it."""
    from moveit_b200 import engine
    rng = np.random.default_rng(700 + seed)
    size, res, corner, maxd = (0.64, 0.56, 0.48), 0.02, (-0.3, -0.2, 0.1), 0.17
    g = engine.Field(size, res, corner, maxd, signed)
    o = oracle.Field(size, res, corner, maxd, signed)
    lo, hi = np.array(corner) - 0.1, np.array(corner) + np.array(size) + 0.1  # some points fall outside the grid
    held = np.zeros((0, 3))

    def rand_pose():
        qv = rng.normal(size=4)
        qv /= np.linalg.norm(qv)
        x, y, z, w = qv
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        return oracle.pose12(rng.uniform(lo + 0.15, hi - 0.15), R)

    shapes = []
    for step in range(40):
        op = rng.integers(0, 8)
        if op == 0 or len(held) == 0:
            pts = rng.uniform(lo, hi, size=(int(rng.integers(1, 60)), 3))
            g.add_points(pts)
            o.add_points(pts)
            held = np.concatenate([held, pts])
        elif op == 1:
            pick = held[rng.random(len(held)) < 0.3]
            extra = rng.uniform(lo, hi, size=(5, 3))  # points that were never added
            pts = np.concatenate([pick, extra])
            g.remove_points(pts)
            o.remove_points(pts)
        elif op == 2:
            old = held[rng.random(len(held)) < 0.5]
            new = np.concatenate([old[: len(old) // 2], rng.uniform(lo, hi, size=(int(rng.integers(1, 30)), 3))])
            g.update_points(old, new)
            o.update_points(old, new)
            held = new
        elif op == 3:
            t = [rd.SHAPE_SPHERE, rd.SHAPE_BOX, rd.SHAPE_CYLINDER][rng.integers(0, 3)]
            dims = {rd.SHAPE_SPHERE: (rng.uniform(0.03, 0.1),), rd.SHAPE_BOX: tuple(rng.uniform(0.04, 0.2, 3)),
                    rd.SHAPE_CYLINDER: (rng.uniform(0.03, 0.08), rng.uniform(0.05, 0.25))}[t]
            pose = rand_pose()
            g.add_shape(t, dims, pose)
            o.add_points(oracle.shape_points(t, dims, pose, res))
            shapes.append((t, dims, pose))
        elif op == 4 and shapes:
            t, dims, pose = shapes.pop(int(rng.integers(0, len(shapes))))
            g.remove_shape(t, dims, pose)
            o.remove_points(oracle.shape_points(t, dims, pose, res))
        elif op == 5:
            leaves = np.concatenate([rng.uniform(lo + 0.1, hi - 0.1, size=(4, 3)), rng.choice([0.01, 0.02, 0.05], (4, 1))],
                                    axis=1)
            g.add_octree_leaves(leaves)
            o.add_points(oracle.octree_points(res, [tuple(l) for l in leaves]))
        elif op == 6:
            ijk = np.stack([rng.integers(-2, n + 2, 20) for n in g.dims], axis=1)  # some outside the grid
            g.add_voxels(ijk)
            o.add_voxels(ijk)
        elif op == 7 and rng.random() < 0.3:
            g.reset()
            o.reset()
            held = np.zeros((0, 3))
            shapes = []
        occ = (o.d2() == 0).astype(np.uint8)
        assert np.array_equal(g.occupancy(), occ), (seed, step, int(op))
        assert np.array_equal(g.d2(), oracle.edt_exact(occ, g.max_d2)), (seed, step, int(op))
        if signed:
            assert np.array_equal(g.d2(True), oracle.edt_exact(1 - occ, g.max_d2)), (seed, step, int(op))


def test_edge_check_slerps_floating_joints(gpu, oracle):
    """b200df_check_edges on a floating-root tree: the quaternion block is interpolated like
    FloatingJointModel::interpolate (slerp on |dot|, translation linear)."""
    seed, n_links, floating, kw = _FUZZ[4]
    desc = _random_robot(100 + seed, n_links, floating, **kw)
    assert desc.joint_type[0] == rd.JOINT_FLOATING
    fd = scenarios.FIELD_DEFAULT
    objs = scenarios.random_scene(seed=200 + seed, n_boxes=6, n_cylinders=3)
    o = OracleSetup(oracle, desc, fd, objs, self_state=np.zeros(desc.nvar))
    e = EngineSetup(desc, fd, objs)
    M, K = 2500, 10
    qa, qb = desc.sample_states(M, 81), desc.sample_states(M, 82)
    qb[:50, 3:7] = qa[:50, 3:7]  # identical orientations: the theta <= eps branch
    angular = np.zeros(desc.nvar, dtype=np.uint8)
    angular[3], angular[4:7] = 2, 3
    flags = CHECK_ROBOT | CHECK_INTRA
    got = e.group.check_edges(qa, qb, K, e.world, flags=flags, precision=0, var_is_angular=angular)
    states = _interpolate_reference(qa, qb, K, angular)
    ref = o.env.check_batch(states.reshape(-1, desc.nvar), mode=3, n_threads=8).reshape(M, K)
    assert np.array_equal(got, np.where(ref.any(axis=1), ref.argmax(axis=1) + 1, 0))
    assert 0 < (got == 0).sum() < M
    lin = e.group.check_edges(qa, qb, K, e.world, flags=flags, precision=0)  # plain lerp of the quaternion: not the same
    assert not np.array_equal(lin, got)


def test_check_route_reports_the_kernel_and_why_fast_is_not_used(gpu, oracle):
    """b200df_check_route: the EXACT fall-back of a FAST call is visible to the caller, never silent."""
    import ctypes as C
    from moveit_b200 import binding as B
    from common import FUZZ_CASES, random_robot

    def route(e, n, precision=B.PRECISION_FAST, flags=CHECK_ROBOT | CHECK_INTRA, dtype=B.Q_F32, world=None):
        r, w = C.c_int32(-1), C.c_int32(-1)
        B.check(B.load().b200df_check_route(e.group.h, (world or e.world).h, None, dtype, n, flags, precision,
                                            C.byref(r), C.byref(w)))
        return r.value, w.value

    e = EngineSetup(rd.panda(), scenarios.FIELD_A, scenarios.random_scene())
    assert route(e, 10_000_000) == (B.ROUTE_FAST, 0)
    assert route(e, 10_000_000, precision=B.PRECISION_EXACT) == (B.ROUTE_EXACT, 0)
    assert route(e, 100) == (B.ROUTE_SMALL, B.NOFAST_BATCH)
    assert route(e, 40_000) == (B.ROUTE_EXACT, B.NOFAST_BATCH)
    # a robot beyond FAST's constant-bank tables (> 128 spheres): exact kernels, and the reason says so
    seed, n_links, floating, kw = FUZZ_CASES[3]
    big = random_robot(seed, n_links, floating, **kw)
    eb = EngineSetup(big, scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=9, n_boxes=2, n_cylinders=1))
    r, w = route(eb, 1_000_000)
    assert r != B.ROUTE_FAST and (w & (B.NOFAST_TABLES | B.NOFAST_SMEM))
    # a signed world field with a collision tolerance above some radius needs the fp64 signed predicate
    es = EngineSetup(rd.panda(), scenarios.FIELD_A, scenarios.random_scene(), signed=True, tol=0.2)
    r, w = route(es, 1_000_000)
    assert r == B.ROUTE_EXACT and (w & B.NOFAST_SIGNED)
    es2 = EngineSetup(rd.panda(), scenarios.FIELD_A, scenarios.random_scene(), signed=True)
    assert route(es2, 1_000_000) == (B.ROUTE_FAST, 0)  # radius > tolerance everywhere: the negative field is irrelevant


def test_bounds_checked_build_runs_every_kernel_family(gpu):
    """compute-sanitizer is closed on the pool, so the library is also built with -DB200DF_BOUNDS_CHECK
    (moveit_b200/libb200df_checked.so, __graft_entry__.build): every computed voxel / list index is tested on the device
    before use and a violation traps.  tools/sanitize_smoke.py drives every kernel family on that build."""
    import os
    import subprocess
    import sys
    from moveit_b200 import build as b
    lib = b.OUT_CHECKED
    if not os.path.exists(lib):
        pytest.skip("libb200df_checked.so is not built (python -m moveit_b200.build --checked)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200DF_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_smoke.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "sanitize_smoke ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
    assert "bounds check failed" not in r.stdout + r.stderr
