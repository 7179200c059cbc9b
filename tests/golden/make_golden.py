#!/usr/bin/env python
"""Generates the committed golden fixtures from the CPU oracle (the faithful restatement of the reference; the
reference itself cannot be built in this image).  Run from the repo root:

    python tests/golden/make_golden.py

Fixtures (small, deterministic, seeded from the portable SplitMix64 stream):
  config1_panda_verdicts.npz   BASELINE config 1: 10 000 Panda states, Field A scene, checkRobotCollision verdicts
                               (exact-EDT field and reference-propagation field), + robot|self verdicts (config 2 at
                               oracle size), bit-packed
  field_a_scene_d2.npz         d^2 of the Field A bench scene: exact EDT and the reference's propagation, as uint8
  small_field_sequence.npz     the add / update / remove point sequence of test_distance_field.cpp:329-429 (d^2 after
                               each step, reference propagation)
  mobile_arm_verdicts.npz      20 000 states of the mixed-joint robot, robot|intra verdicts
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import oracle_py
    from common import OracleSetup
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios

    desc = rd.panda()
    objs = scenarios.random_scene()
    q = desc.sample_states(10000, scenarios.SEED_STATES, np.float64)
    out = {}
    for name, exact in (("exact", True), ("propagation", False)):
        o = OracleSetup(oracle_py, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar), exact_field=exact)
        out["robot_" + name] = np.packbits(o.env.check_batch(q, mode=1, n_threads=8))
        out["all_" + name] = np.packbits(o.env.check_batch(q, mode=3, n_threads=8))
        if exact:
            d2_exact = o.env.world.d2().astype(np.uint8)
        else:
            d2_prop = o.env.world.d2().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "config1_panda_verdicts.npz"), n=10000, seed=scenarios.SEED_STATES, **out)
    np.savez_compressed(os.path.join(HERE, "field_a_scene_d2.npz"), exact=d2_exact, propagation=d2_prop)

    f = oracle_py.Field((1.0, 1.0, 1.0), 0.1, (0.0, 0.0, 0.0), 0.3)
    p1, p2, p3 = (0.1, 0.0, 0.0), (0.0, 0.1, 0.2), (0.4, 0.0, 0.0)
    steps = {}
    f.add_points([p1, p2])
    steps["add_p1_p2"] = f.d2().astype(np.uint8)
    f.update_points([p1, p2], [p2, p3])
    steps["update_to_p2_p3"] = f.d2().astype(np.uint8)
    f.remove_points([p2])
    steps["remove_p2"] = f.d2().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "small_field_sequence.npz"), **steps)

    ma = rd.mobile_arm()
    o = OracleSetup(oracle_py, ma, scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=57, n_boxes=3, n_cylinders=1),
                    self_state=np.zeros(ma.nvar))
    qm = ma.sample_states(20000, 606, np.float32).astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "mobile_arm_verdicts.npz"), n=20000, seed=606,
                        all_exact=np.packbits(o.env.check_batch(qm, mode=3, n_threads=8)))
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")


if __name__ == "__main__":
    main()
