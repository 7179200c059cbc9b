#!/usr/bin/env python
"""EDT rebuild timings (Field A from the bench scene, Field B from the 1 M-point cloud) without the validity bench."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from moveit_b200 import engine, scenarios  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(bench.edt_timings(engine, scenarios, None)))
