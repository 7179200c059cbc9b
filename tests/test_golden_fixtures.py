"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle): the oracle
must keep reproducing them on the CPU, and the CUDA path must reproduce them through the C ABI on the GPU."""
import os

import numpy as np
import pytest

from moveit_b200 import robot_description as rd
from moveit_b200 import scenarios
from moveit_b200.binding import CHECK_INTRA, CHECK_ROBOT

from common import EngineSetup, OracleSetup

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _unpack(a, n):
    return np.unpackbits(a)[:n]


def test_oracle_reproduces_config1_fixture(oracle):
    fx = np.load(os.path.join(G, "config1_panda_verdicts.npz"))
    desc = rd.panda()
    n = int(fx["n"])
    q = desc.sample_states(n, int(fx["seed"]), np.float64)
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, scenarios.random_scene(), self_state=np.zeros(desc.nvar))
    assert np.array_equal(o.env.check_batch(q, mode=1, n_threads=4), _unpack(fx["robot_exact"], n))
    assert np.array_equal(o.env.check_batch(q, mode=3, n_threads=4), _unpack(fx["all_exact"], n))
    d2 = np.load(os.path.join(G, "field_a_scene_d2.npz"))
    assert np.array_equal(o.env.world.d2().astype(np.uint8), d2["exact"])
    # the reference's own (inexact) propagation result and what it changes: a few hundred voxels, two verdicts
    o2 = OracleSetup(oracle, desc, scenarios.FIELD_A, scenarios.random_scene(), self_state=np.zeros(desc.nvar),
                     exact_field=False)
    assert np.array_equal(o2.env.world.d2().astype(np.uint8), d2["propagation"])
    diff = d2["propagation"].astype(np.int32) - d2["exact"].astype(np.int32)
    assert diff.min() >= 0 and 0 < (diff != 0).sum() < 2000
    assert (_unpack(fx["robot_exact"], n) != _unpack(fx["robot_propagation"], n)).sum() <= 5


def test_oracle_reproduces_small_field_sequence(oracle):
    fx = np.load(os.path.join(G, "small_field_sequence.npz"))
    f = oracle.Field((1.0, 1.0, 1.0), 0.1, (0.0, 0.0, 0.0), 0.3)
    p1, p2, p3 = (0.1, 0.0, 0.0), (0.0, 0.1, 0.2), (0.4, 0.0, 0.0)
    f.add_points([p1, p2])
    assert np.array_equal(f.d2().astype(np.uint8), fx["add_p1_p2"])
    f.update_points([p1, p2], [p2, p3])
    assert np.array_equal(f.d2().astype(np.uint8), fx["update_to_p2_p3"])
    f.remove_points([p2])
    assert np.array_equal(f.d2().astype(np.uint8), fx["remove_p2"])


@pytest.mark.gpu
def test_engine_reproduces_config1_fixture(gpu):
    fx = np.load(os.path.join(G, "config1_panda_verdicts.npz"))
    desc = rd.panda()
    n = int(fx["n"])
    q = desc.sample_states(n, int(fx["seed"]), np.float64)
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    assert np.array_equal(e.group.check(q, e.world, flags=CHECK_ROBOT), _unpack(fx["robot_exact"], n))
    assert np.array_equal(e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA), _unpack(fx["all_exact"], n))
    d2 = np.load(os.path.join(G, "field_a_scene_d2.npz"))
    assert np.array_equal(e.world.d2().astype(np.uint8), d2["exact"])
    # the same states rounded to fp32 (what the FAST path takes): FAST == EXACT, and all but a handful of states
    # keep the fixture's verdict
    q32 = q.astype(np.float32)
    fast = e.group.check(q32, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=1)
    assert np.array_equal(fast, e.group.check(q32, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=0))
    assert (fast != _unpack(fx["all_exact"], n)).sum() < 10


@pytest.mark.gpu
def test_engine_reproduces_small_field_and_mobile_arm_fixtures(gpu):
    from moveit_b200 import engine
    fx = np.load(os.path.join(G, "small_field_sequence.npz"))
    f = engine.Field((1.0, 1.0, 1.0), 0.1, (0.0, 0.0, 0.0), 0.3)
    p1, p2, p3 = (0.1, 0.0, 0.0), (0.0, 0.1, 0.2), (0.4, 0.0, 0.0)
    f.add_points([p1, p2])
    assert np.array_equal(f.d2().astype(np.uint8), fx["add_p1_p2"])  # few points: propagation == exact here
    f.update_points([p1, p2], [p2, p3])
    assert np.array_equal(f.d2().astype(np.uint8), fx["update_to_p2_p3"])
    f.remove_points([p2])
    assert np.array_equal(f.d2().astype(np.uint8), fx["remove_p2"])
    fm = np.load(os.path.join(G, "mobile_arm_verdicts.npz"))
    ma = rd.mobile_arm()
    n = int(fm["n"])
    q = ma.sample_states(n, int(fm["seed"]), np.float32)
    e = EngineSetup(ma, scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=57, n_boxes=3, n_cylinders=1))
    for precision in (0, 1):
        got = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=precision)
        assert np.array_equal(got, _unpack(fm["all_exact"], n))
