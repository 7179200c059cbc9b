#!/usr/bin/env python
"""Field B (config 3) rebuilt a few times: the command the EDT ncu captures run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moveit_b200 import engine, scenarios  # noqa: E402

fb = scenarios.FIELD_B
pts = scenarios.synthetic_cloud(1_000_000)
g = engine.Field.centered(fb["size"], fb["resolution"], fb["center"], fb["max_distance"])
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    g.reset()
    g.add_points(pts)
    print("rebuild %.4f ms" % g.rebuild())
