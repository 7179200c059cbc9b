"""A short loop of small calls (1 state, 256 states, a 100-waypoint gradient batch, one edge) for ncu captures of the
CTA-per-state kernels.  Prints nothing but a one-line summary."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    from moveit_b200 import binding as B
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios
    from common import EngineSetup
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    fl = B.CHECK_ROBOT | B.CHECK_INTRA
    q = desc.sample_states(4096, scenarios.SEED_STATES, np.float32)
    traj = scenarios.chomp_trajectories(desc, 1, 100).reshape(-1, desc.nvar)
    hits = 0
    for k in range(10):
        hits += int(e.group.check(q[k:k + 1], e.world, flags=fl)[0])
        hits += int(e.group.check(q[:256], e.world, flags=fl).sum())
        hits += int(e.group.check(q[:2048], e.world, flags=fl).sum())
        e.group.gradients(traj, e.world, flags=fl)
        e.group.check_edges(q[:1].astype(np.float64), q[1:2].astype(np.float64), 16, e.world, flags=fl)
    print("small_call_loop ok, %d hits, %d launches" % (hits, B.load().b200df_launch_count()))


if __name__ == "__main__":
    main()
