"""Single-process multi-GPU (b200df_multi) and cross-device field replication.  The correctness tests need two visible
devices and skip otherwise (the driver's GPU test box has one; run with `gpurun --gpus 2`); the one-device test runs the
same code path with a single replica."""
import numpy as np
import pytest

from moveit_b200 import robot_description as rd
from moveit_b200 import scenarios
from moveit_b200.binding import CHECK_INTRA, CHECK_ROBOT, CHECK_SELF

from common import EngineSetup, OracleSetup, self_field_points

pytestmark = pytest.mark.gpu


def _setup(devices):
    from moveit_b200 import binding, engine
    from moveit_b200.engine import resolve_acm
    binding.check(binding.load().b200df_set_device(devices[0]))
    desc = rd.panda()
    objs = scenarios.random_scene()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    se, ie = resolve_acm(e.model_desc, e.group_links, desc.acm_entry_matrix())
    m = engine.MultiGroup(e.model_desc, e.group_links, se, ie, devices, scenarios.FIELD_A["resolution"], 0.0,
                          scenarios.FIELD_A["max_distance"])
    m.set_field(0, e.world)
    return desc, objs, e, m


def test_multi_with_one_replica_equals_the_plain_call(gpu, oracle):
    desc, objs, e, m = _setup([0])
    q = desc.sample_states(150_000, 99, np.float32)
    flags = CHECK_ROBOT | CHECK_INTRA
    want = e.group.check(q, e.world, flags=flags, precision=1)
    assert np.array_equal(m.check(q, flags=flags, precision=1), want)
    assert np.array_equal(m.check(q[:777], flags=flags, precision=0), want[:777])
    assert np.array_equal(m.replica_d2(0, 0, e.world.dims), e.world.d2())
    # a change of the source only reaches the replica through set_field
    e.world.add_points([(0.3, 0.0, 0.5)])
    assert not np.array_equal(m.replica_d2(0, 0, e.world.dims), e.world.d2())
    m.set_field(0, e.world)
    assert np.array_equal(m.replica_d2(0, 0, e.world.dims), e.world.d2())


def test_two_gpus_field_replication_and_sharded_verdicts(gpu, oracle):
    """Field built on device 0, replicated to device 1 by b200df_field_copy_from (NVLink peer copy); a batch sharded
    over both devices in one process gives the oracle's verdicts, state by state, on BOTH halves."""
    from moveit_b200 import binding, engine
    if binding.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    desc, objs, e, m = _setup([0, 1])
    for slot in (0, 1):
        assert np.array_equal(m.replica_d2(0, slot, e.world.dims), e.world.d2())
    q = desc.sample_states(40_000, 4321)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
    ref = o.env.check_batch(q, mode=3, n_threads=8)
    flags = CHECK_ROBOT | CHECK_INTRA
    for precision in (0, 1):
        got = m.check(q, flags=flags, precision=precision)
        assert np.array_equal(got[:20_000], ref[:20_000]) and np.array_equal(got[20_000:], ref[20_000:])
    big = desc.sample_states(400_000, 5, np.float32)
    assert np.array_equal(m.check(big, flags=flags, precision=1), e.group.check(big, e.world, flags=flags, precision=1))
    # ragged: fewer states than devices, and an odd count
    assert np.array_equal(m.check(q[:1], flags=flags), ref[:1])
    assert np.array_equal(m.check(q[:1001], flags=flags), ref[:1001])
    # explicit copy_from between two fields on different devices, then a check on the second device alone
    binding.check(binding.load().b200df_set_device(1))
    fa = scenarios.FIELD_A
    other = engine.Field.centered(fa["size"], fa["resolution"], fa["center"], fa["max_distance"])
    other.copy_from(e.world)
    assert np.array_equal(other.d2(), e.world.d2()) and np.array_equal(other.occupancy(), e.world.occupancy())
    binding.check(binding.load().b200df_set_device(0))
    # self field replicas as well
    qs = np.zeros(desc.nvar)
    ep = EngineSetup(desc, fa, objs, group="panda_arm",
                     self_points=self_field_points(e.model_desc, desc, desc.group_links("panda_arm"),
                                                   e.model.fk(qs[None, :])[0]))
    se, ie = engine.resolve_acm(ep.model_desc, ep.group_links, desc.acm_entry_matrix())
    mp = engine.MultiGroup(ep.model_desc, ep.group_links, se, ie, [0, 1])
    mp.set_field(0, ep.world)
    mp.set_field(1, ep.self_field)
    op = OracleSetup(oracle, desc, fa, objs, group="panda_arm", self_state=qs)
    allf = CHECK_ROBOT | CHECK_SELF | CHECK_INTRA
    assert np.array_equal(mp.check(q, flags=allf), op.env.check_batch(q, mode=4, n_threads=8))


def test_cpp_env_shards_batches_over_two_gpus(gpu, oracle, tmp_path, monkeypatch):
    """CollisionEnvB200::setDevices({0, 1}): checkCollisionBatch of the C++ mirror through b200df_multi, world change
    re-replicated through the observer, everything against the oracle."""
    from moveit_b200 import binding
    if binding.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from test_host_mirror import _run
    from common import oracle_world_points
    monkeypatch.setenv("B200DF_DRIVER_DEVICES", "0,1")
    desc = rd.panda()
    objs = scenarios.random_scene(seed=41)
    fa = scenarios.FIELD_A
    q = desc.sample_states(4001, 2024)
    prefix = _run(tmp_path, desc, fa, objs, q, group="panda_arm", n_single=8, mutate=True)
    o = OracleSetup(oracle, desc, fa, objs, group="panda_arm", self_state=np.zeros(desc.nvar))
    robot = np.fromfile(prefix + ".robot.u8", dtype=np.uint8)
    assert np.array_equal(robot, o.env.check_batch(q, mode=1, n_threads=4))
    assert np.array_equal(np.fromfile(prefix + ".self.u8", dtype=np.uint8), o.env.check_batch(q, mode=2, n_threads=4))
    assert np.array_equal(np.fromfile(prefix + ".all.u8", dtype=np.uint8), o.env.check_batch(q, mode=4, n_threads=4))
    pts = [oracle_world_points(oracle, [ob], fa["resolution"]) for ob in objs]
    moved = (objs[0][0], objs[0][1], np.array(objs[0][2]) + np.array([[0, 0, 0, 0.1]] * 3) * np.array([[1], [0], [0]]))
    o.env.world.remove_points(pts[0])
    o.env.world.add_points(oracle_world_points(oracle, [moved], fa["resolution"]))
    o.env.world.remove_points(pts[1])
    o.env.world.make_exact()
    robot2 = np.fromfile(prefix + ".robot2.u8", dtype=np.uint8)
    assert np.array_equal(robot2, o.env.check_batch(q, mode=1, n_threads=4)) and not np.array_equal(robot2, robot)
