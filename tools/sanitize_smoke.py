"""Every kernel family once on small inputs, for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`:
field edits + EDT (TMA form and the fallback), EXACT / FAST (fused and two-pass) / CTA-per-state verdicts, gradients
(both output types, with link fields), motion validation, octree leaves, occupancy packing.  Results are compared with
each other, not with the oracle (tests/ does that); the point is the memory checker's verdict."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from moveit_b200 import binding as B
    from moveit_b200 import engine, scenarios
    from moveit_b200 import robot_description as rd
    from common import EngineSetup
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    fl = B.CHECK_ROBOT | B.CHECK_INTRA
    n = 1 << 18  # large enough for the two-pass FAST form (>= 2^17 states)
    q = desc.sample_states(n + 1000, scenarios.SEED_STATES, np.float32)
    fast = e.group.check(q[:n], e.world, flags=fl, precision=B.PRECISION_FAST)
    exact = e.group.check(q[:n], e.world, flags=fl, precision=B.PRECISION_EXACT)
    assert np.array_equal(fast, exact)
    odd = e.group.check(q[: n - 37], e.world, flags=fl, precision=B.PRECISION_FAST)
    assert np.array_equal(odd, exact[: n - 37])
    small = e.group.check(q[:3000], e.world, flags=fl, precision=B.PRECISION_FAST)
    assert np.array_equal(small, exact[:3000])
    q64 = q[:70000].astype(np.float64)
    assert np.array_equal(e.group.check(q64, e.world, flags=fl, precision=B.PRECISION_FAST), exact[:70000])
    g64 = e.group.gradients(q64[:5000], e.world, flags=fl)
    g32 = e.group.gradients(q64[:5000], e.world, flags=fl, out_dtype=np.float32)
    assert np.array_equal(g64["type"], g32["type"])
    e.group.gradients(q64[:100], e.world, flags=fl)
    first = e.group.check_edges(q64[:2000], q64[2000:4000], 8, e.world, flags=fl, precision=B.PRECISION_FAST)
    assert first.shape == (2000,)
    # Field B-like grid (nz a multiple of 16: the TMA form) and an odd one (the fallback), edits and a rebuild each
    for size, res in (((1.28, 1.28, 1.28), 0.01), ((0.9, 0.7, 0.5), 0.02)):
        f = engine.Field.centered(size, res, (0.0, 0.0, 0.0), 0.25)
        pts = (np.random.default_rng(5).random((20000, 3)) - 0.5) * np.array(size) * 0.9
        f.add_points(pts)
        f.rebuild()
        d2 = f.d2()
        f.remove_points(pts[:5000])
        f.add_octree_leaves(np.array([[0.0, 0.0, 0.0, 0.08], [0.2, 0.1, 0.0, res]]))
        f.rebuild()
        assert f.d2().shape == d2.shape
        packed = f.pack_occupancy() if hasattr(f, "pack_occupancy") else None
        del packed
    print("sanitize_smoke ok: %d kernel launches" % B.load().b200df_launch_count())


if __name__ == "__main__":
    main()
