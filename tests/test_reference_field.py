"""Parity pinned on the REFERENCE's own field values.

The engine's default field is the exact EDT; the reference's PropagationDistanceField::propagatePositive
(propagation_distance_field.cpp:389-444) over-estimates d^2 on a few voxels in 10^4.  Here the
reference's propagated d^2 (committed fixtures, produced by the oracle's faithful restatement) is uploaded into an
engine field (b200df_field_upload_d2, "reference-field mode") and every verdict kernel and the gradient kernel must
reproduce what getCollisionSphereCollision / getCollisionGradients (collision_distance_field_types.cpp:206-231,
collision_env_distance_field.cpp:1536-1557) give on exactly those values.
"""
import hashlib
import os

import numpy as np
import pytest

from moveit_b200 import robot_description as rd
from moveit_b200 import scenarios
from moveit_b200.binding import CHECK_INTRA, CHECK_ROBOT, B200dfError

from common import EngineSetup, OracleSetup

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRECISIONS = dict(exact=0, fast=1, small=4)


def _unpack(a, n):
    return np.unpackbits(a)[:n]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_config3_census_fixture_matches_the_oracle(oracle):
    """Config 3 through the oracle (faithful propagation and exact EDT, ~30 s): the committed census is what DESIGN.md
    quotes - 30 761 of 64 M voxels differ, the reference always larger, by at most 8."""
    fx = np.load(os.path.join(G, "config3_census.npz"))
    fb = scenarios.FIELD_B
    f = oracle.Field(fb["size"], fb["resolution"], scenarios.field_corner(fb), fb["max_distance"])
    f.add_points(scenarios.synthetic_cloud(int(fx["n_points"])))
    ref = f.d2().astype(np.int32)
    assert _sha(ref) == str(fx["sha_reference"])
    f.make_exact()
    exact = f.d2().astype(np.int32)
    assert _sha(exact) == str(fx["sha_exact"])
    diff = (ref - exact).ravel()
    idx = np.flatnonzero(diff)
    assert np.array_equal(idx, fx["index"]) and np.array_equal(diff[idx], fx["delta"])
    assert len(idx) == 30761 and int(diff.max()) == int(fx["max_delta"]) == 8 and diff.min() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", sorted(PRECISIONS))
def test_verdicts_on_the_reference_propagated_field(gpu, kernel):
    """config1_panda_verdicts.npz: robot_propagation / all_propagation were produced on the reference's propagated
    d^2 (field_a_scene_d2.npz: propagation).  EXACT, FAST and the one-CTA-per-state kernel, bit for bit."""
    fx = np.load(os.path.join(G, "config1_panda_verdicts.npz"))
    d2 = np.load(os.path.join(G, "field_a_scene_d2.npz"))
    desc = rd.panda()
    n = int(fx["n"])
    q = desc.sample_states(n, int(fx["seed"]), np.float64)
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    assert np.array_equal(e.world.d2().astype(np.uint8), d2["exact"])
    assert not e.world.values_uploaded()
    e.world.upload_d2(d2["propagation"])
    assert e.world.values_uploaded()
    assert np.array_equal(e.world.d2().astype(np.uint8), d2["propagation"])
    p = PRECISIONS[kernel]
    robot = e.group.check(q, e.world, flags=CHECK_ROBOT, precision=p)
    both = e.group.check(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA, precision=p)
    assert np.array_equal(robot, _unpack(fx["robot_propagation"], n))
    assert np.array_equal(both, _unpack(fx["all_propagation"], n))
    # the fixture really distinguishes the two fields
    assert (robot != _unpack(fx["robot_exact"], n)).sum() > 0


@pytest.mark.gpu
def test_gradients_on_the_reference_propagated_field(gpu, oracle):
    d2 = np.load(os.path.join(G, "field_a_scene_d2.npz"))
    desc = rd.panda()
    objs = scenarios.random_scene()
    o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar), exact_field=False)
    assert np.array_equal(o.env.world.d2().astype(np.uint8), d2["propagation"])
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    e.world.upload_d2(o.env.world.d2())
    q = desc.sample_states(3000, 4242, np.float64)
    g = e.group.gradients(q, e.world, flags=CHECK_ROBOT | CHECK_INTRA)
    r = o.env.gradients_batch(q, flags=2 | 4)
    assert np.array_equal(g["type"], r["type"].astype(np.uint8))
    fin = r["dist"] < 1e300
    assert np.max(np.abs(g["dist"][fin] - r["dist"][fin])) <= 1e-5
    assert np.max(np.abs(g["grad"][fin] - r["grad"][fin])) <= 1e-5
    # and the exact field gives (slightly) different distances somewhere: the test would notice a silent rebuild
    e2 = EngineSetup(desc, scenarios.FIELD_A, objs)
    g2 = e2.group.gradients(q, e2.world, flags=CHECK_ROBOT | CHECK_INTRA)
    assert np.max(np.abs(g2["dist"][fin] - r["dist"][fin])) > 1e-5


@pytest.mark.gpu
def test_uploaded_values_are_dropped_by_the_next_edit_and_bad_uploads_are_refused(gpu, oracle):
    from moveit_b200 import engine
    f = engine.Field((1.0, 1.0, 1.0), 0.1, (0.0, 0.0, 0.0), 0.3)
    o = oracle.Field((1.0, 1.0, 1.0), 0.1, (0.0, 0.0, 0.0), 0.3)
    for x in (f, o):
        x.add_points([(0.1, 0.0, 0.0), (0.5, 0.5, 0.5)])
    exact = f.d2()
    fake = np.minimum(exact + (exact > 0) * 1, f.max_d2)  # "over-estimated" everywhere outside the obstacles
    f.upload_d2(fake)
    assert f.values_uploaded() and np.array_equal(f.d2(), fake)
    d, _, inb = f.distance_gradient([o.grid_to_world(3, 3, 3)])
    assert inb[0] and d[0] == np.sqrt(float(fake[3, 3, 3])) * 0.1
    assert np.array_equal(f.occupancy(), (fake == 0).astype(np.uint8))
    f.add_points([(0.9, 0.9, 0.9)])  # an edit: back to the engine's own exact EDT of the new occupancy
    o.add_points([(0.9, 0.9, 0.9)])
    o.make_exact()
    assert not f.values_uploaded() and np.array_equal(f.d2(), o.d2())
    bad = exact.copy()
    bad[0, 0, 0] = f.max_d2 + 1
    with pytest.raises(B200dfError):
        f.upload_d2(bad)
    assert np.array_equal(f.d2(), o.d2())  # refused upload left the field untouched
    s = engine.Field((1.0, 1.0, 1.0), 0.1, (0.0, 0.0, 0.0), 0.3, True)
    with pytest.raises(B200dfError):
        s.upload_d2(exact)  # a signed field needs both arrays
    s.upload_d2(exact, np.zeros_like(exact))
    assert np.array_equal(s.d2(), exact) and np.array_equal(s.d2(True), np.zeros_like(exact))


@pytest.mark.gpu
def test_config3_reference_values_from_the_census_fixture(gpu):
    """Field B on the GPU: the engine's exact EDT has the fixture's hash; exact + census delta IS the reference's
    propagated d^2 (hash pinned by the CPU test above); uploaded, the field serves exactly those 64 M values."""
    from moveit_b200 import engine
    fx = np.load(os.path.join(G, "config3_census.npz"))
    fb = scenarios.FIELD_B
    g = engine.Field.centered(fb["size"], fb["resolution"], fb["center"], fb["max_distance"])
    g.add_points(scenarios.synthetic_cloud(int(fx["n_points"])))
    d2 = g.d2()
    assert _sha(g.occupancy()) == str(fx["sha_occupancy"])
    assert _sha(d2) == str(fx["sha_exact"])
    ref = d2.copy().ravel()
    ref[fx["index"]] += fx["delta"]
    assert _sha(ref) == str(fx["sha_reference"])
    g.upload_d2(ref)
    assert _sha(g.d2()) == str(fx["sha_reference"])
    # a lookup that lands on a differing voxel sees the reference's value
    k = int(fx["index"][len(fx["index"]) // 2])
    x, y, z = np.unravel_index(k, g.dims)
    if 0 < x < 399 and 0 < y < 399 and 0 < z < 399:
        c = np.array(scenarios.field_corner(fb)) + fb["resolution"] * np.array([x, y, z])
        d, _, inb = g.distance_gradient([c])
        assert inb[0] and d[0] == np.sqrt(float(ref[k])) * fb["resolution"]


@pytest.mark.gpu
def test_nearest_cells(gpu, oracle):
    """getNearestCell (propagation_distance_field.h:411-431): distance as the reference, position = an obstacle
    (free, for cells inside obstacles of a signed field) voxel at exactly that squared distance."""
    from moveit_b200 import engine
    f = engine.Field((1.0, 1.0, 1.0), 0.05, (0.0, 0.0, 0.0), 0.3, True)
    rng = np.random.default_rng(3)
    vox = rng.integers(2, 18, size=(12, 3))
    block = np.array([(x, y, z) for x in range(8, 12) for y in range(8, 12) for z in range(8, 12)])
    f.add_voxels(np.concatenate([vox, block]))
    d2, neg, occ = f.d2(), f.d2(True), f.occupancy()
    cells = np.array([(x, y, z) for x in range(20) for y in range(20) for z in range(0, 20, 3)], dtype=np.int32)
    dist, pos, found = f.nearest_cells(np.concatenate([cells, [[-1, 0, 0], [20, 5, 5]]]))
    assert not found[-1] and not found[-2]
    for c, d, p, ok in zip(cells, dist[:-2], pos[:-2], found[:-2]):
        v, w = int(d2[tuple(c)]), int(neg[tuple(c)])
        if v > 0:
            assert d == np.sqrt(float(v)) * 0.05
            assert ok == (v < f.max_d2)
            if ok:
                assert occ[tuple(p)] == 1 and int(((p - c) ** 2).sum()) == v
        else:
            assert w > 0 and d == -np.sqrt(float(w)) * 0.05
            if ok:
                assert occ[tuple(p)] == 0 and int(((p - c) ** 2).sum()) == w
