"""b200df_check_batch from page-locked fp32 rows into a page-locked verdict array (bench.py's `e2e`), 10 M Panda states, on
its own: for sweeps of the pipeline knobs (B200DF_CHUNK_STATES)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch
    from moveit_b200 import binding as B
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios
    from common import EngineSetup
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    q = torch.from_numpy(desc.sample_states(n, scenarios.SEED_STATES, np.float32)).pin_memory().numpy()
    out = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    fl = B.CHECK_ROBOT | B.CHECK_INTRA
    e.group.check(q, e.world, flags=fl, precision=B.PRECISION_FAST, out=out)
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter()
        e.group.check(q, e.world, flags=fl, precision=B.PRECISION_FAST, out=out)
        best = min(best, time.perf_counter() - t0)
    print("chunk=%s: %.3f ms, %.4e states/s, %d colliding" % (os.environ.get("B200DF_CHUNK_STATES", "default"), best * 1e3,
                                                             n / best, int(out.sum())))


if __name__ == "__main__":
    main()
