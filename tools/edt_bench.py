#!/usr/bin/env python
"""EDT rebuild timings (device events inside b200df_field_rebuild) for Field A / Field B over the segment settings of
edt.cu, each result checked against the committed fixtures (hash of the exact EDT of config 3, d^2 of the Field A
scene).  Usage: python tools/edt_bench.py [--reps 5] [--segs 1,2,3,4]"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moveit_b200 import engine, scenarios  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--combos", default=";sy=1,sx=1;sy=2,sx=2;sy=3,sx=3;lin=1;lin=1,sy=2,sx=2",
                    help="semicolon-separated settings, each a comma list of sy/sx (segments), my/mx (dynamic shared "
                         "memory bytes: an occupancy knob), lin=1 (no tiled footprint)")
    args = ap.parse_args()
    out = {}
    fx = np.load(os.path.join(G, "config3_census.npz"))
    fb = scenarios.FIELD_B
    pts = scenarios.synthetic_cloud(1_000_000)
    g = engine.Field.centered(fb["size"], fb["resolution"], fb["center"], fb["max_distance"])
    fa = scenarios.FIELD_A
    a = engine.Field.centered(fa["size"], fa["resolution"], fa["center"], fa["max_distance"])
    objs = scenarios.random_scene()
    d2a = np.load(os.path.join(G, "field_a_scene_d2.npz"))["exact"]
    combos = []
    for spec in args.combos.split(";"):
        env = dict(kv.split("=") for kv in spec.split(",") if kv)
        combos.append(env)
    keys = ["B200DF_EDT_SEG_Y", "B200DF_EDT_SEG_X", "B200DF_EDT_SMEM_Y", "B200DF_EDT_SMEM_X", "B200DF_EDT_LINEAR",
            "B200DF_EDT_SLABS"]
    short = dict(sy=keys[0], sx=keys[1], my=keys[2], mx=keys[3], lin=keys[4], sl=keys[5])
    for env in combos:
        for k in keys:
            os.environ.pop(k, None)
        for k, v in env.items():
            os.environ[short[k]] = v
        tb = []
        for _ in range(args.reps):
            g.reset()
            g.add_points(pts)
            tb.append(g.rebuild())
        ok_b = sha(g.d2()) == str(fx["sha_exact"])
        ta = []
        for _ in range(args.reps):
            a.reset()
            for t, dims, pose in objs:
                a.add_posed_body(t, dims, pose)
            ta.append(a.rebuild())
        ok_a = bool(np.array_equal(a.d2().astype(np.uint8), d2a))
        name = ",".join("%s=%s" % kv for kv in sorted(env.items())) or "default"
        out[name] = dict(field_b_ms=min(tb), field_b_ms_all=tb, field_b_exact=ok_b, field_a_ms=min(ta),
                         field_a_exact=ok_a, field_b_frac=192e6 / (min(tb) * 1e-3) / 6546.2e9)
        print("%-40s B %.4f ms (%s)  A %.4f ms (%s)" % (name, min(tb), ok_b, min(ta), ok_a), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
