#!/usr/bin/env python
"""Generates tests/golden/config3_census.npz: BASELINE config 3 (Field B, 400^3 at 1 cm, the 1M-point synthetic
cloud) through the CPU oracle's faithful restatement of the reference's propagation AND its exact EDT.

    python tests/golden/make_config3_census.py        (about a minute, ~6 GB of host memory)

Stored (small): the flat indices of the voxels where the reference's propagated d^2 differs from the exact EDT, the
differences (reference - exact, always > 0), SHA-256 of both full int32 arrays and of the occupancy, and the counts
DESIGN.md quotes.  With it a GPU test can reconstruct the reference's own d^2 from the engine's exact EDT (exact +
delta) without running the 22 s propagation, check the hash, and upload it as a reference-mode field.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import oracle_py
    from moveit_b200 import scenarios
    fb = scenarios.FIELD_B
    pts = scenarios.synthetic_cloud(1_000_000)
    f = oracle_py.Field(fb["size"], fb["resolution"], scenarios.field_corner(fb), fb["max_distance"])
    f.add_points(pts)
    ref = f.d2().astype(np.int32)
    f.make_exact()
    exact = f.d2().astype(np.int32)
    diff = ref - exact
    idx = np.flatnonzero(diff)
    assert diff.min() >= 0
    out = dict(index=idx.astype(np.int32), delta=diff.ravel()[idx].astype(np.uint8), n_points=1_000_000,
               n_voxels=ref.size, n_occupied=int((exact == 0).sum()), max_delta=int(diff.max()),
               sha_reference=sha(ref), sha_exact=sha(exact), sha_occupancy=sha((exact == 0).astype(np.uint8)))
    np.savez_compressed(os.path.join(HERE, "config3_census.npz"), **out)
    print("config 3 census: %d of %d voxels differ, max +%d; occupied %d" % (len(idx), ref.size, out["max_delta"],
                                                                           out["n_occupied"]))


if __name__ == "__main__":
    main()
