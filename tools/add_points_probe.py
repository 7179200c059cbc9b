import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from moveit_b200 import engine, scenarios
fb = scenarios.FIELD_B
pts = scenarios.synthetic_cloud(1_000_000)
g = engine.Field.centered(fb["size"], fb["resolution"], fb["center"], fb["max_distance"])
best = 1e9
for _ in range(12):
    g.reset()
    t0 = time.perf_counter(); g.add_points(pts); t = time.perf_counter() - t0
    best = min(best, t)
print("workers=%s add_points 1M host: %.3f ms" % (os.environ.get("B200DF_ADD_POINTS_WORKERS", "default"), best * 1e3))
