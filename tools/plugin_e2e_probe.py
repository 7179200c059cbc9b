"""b200df_check_batch from pageable fp64 rows (the shape a CollisionEnv caller has), 10 M Panda states: the e2e_plugin
number of bench.py on its own, for sweeps of the staging knobs (B200DF_NARROW_LANES)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    from moveit_b200 import binding as B
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios
    from common import EngineSetup
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    q = desc.sample_states(n, scenarios.SEED_STATES, np.float32).astype(np.float64)
    fl = B.CHECK_ROBOT | B.CHECK_INTRA
    out = np.zeros(n, dtype=np.uint8)
    e.group.check(q, e.world, flags=fl, precision=B.PRECISION_FAST, out=out)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        e.group.check(q, e.world, flags=fl, precision=B.PRECISION_FAST, out=out)
        best = min(best, time.perf_counter() - t0)
    print("lanes=%s: %.2f ms, %.4e states/s, %d colliding" % (os.environ.get("B200DF_NARROW_LANES", "default"),
                                                              best * 1e3, n / best, int(out.sum())))


if __name__ == "__main__":
    main()
