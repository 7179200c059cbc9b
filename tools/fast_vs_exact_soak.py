#!/usr/bin/env python
"""Soak test of the FAST precision's exactness claim (and of the small-batch kernel's): many batches of fresh random
states (different seeds, three scenes, three robots incl. a floating-root tree), FAST and the one-CTA-per-state kernel vs EXACT verdicts compared on
the device, undecided fraction of the screening pass reported.
Prints one JSON line; the committed result lives in profiles/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(batches=10, n=10_000_000):
    import torch
    from moveit_b200 import binding as B
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios
    from moveit_b200.rng import uniform01
    from common import EngineSetup
    out = {"batches": [], "states": 0, "differences": 0}
    stream = torch.cuda.current_stream().cuda_stream
    cases = [("panda", rd.panda(), scenarios.FIELD_A, scenarios.random_scene()),
             ("panda_fine_field", rd.panda(),
              dict(size=(2.0, 2.0, 2.0), resolution=0.01, center=(0.0, 0.0, 0.5), max_distance=0.25),
              scenarios.random_scene(seed=1234)),
             ("mobile_arm", rd.mobile_arm(), scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=57, n_boxes=3, n_cylinders=1))]
    from common import FUZZ_CASES, random_robot
    seed, n_links, floating, kw = FUZZ_CASES[4]  # the floating-root tree of the fuzz tests (quaternion joint in FAST)
    cases.append(("fuzz_floating_root", random_robot(100 + seed, n_links, floating, **kw), scenarios.FIELD_DEFAULT,
                  scenarios.random_scene(seed=200 + seed, n_boxes=6, n_cylinders=3)))
    # the PR2-like tree, self-collision only (72 % colliding): FAST runs its pair tests one warp per state there
    cases.append(("pr2_like_self_collision", rd.pr2_like(), scenarios.FIELD_DEFAULT, None))
    n_full = n
    for name, desc, field, objs in cases:
        self_only = objs is None
        n = max(1, n_full // 10) if self_only else n_full  # the fp64 kernels take 40 ms per million states of this tree
        e = EngineSetup(desc, field, objs or [])
        lo = np.array([l[0] for l in desc.limits], dtype=np.float32)
        hi = np.array([l[1] for l in desc.limits], dtype=np.float32)
        lo_t, hi_t = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
        flags = B.CHECK_INTRA if self_only else B.CHECK_ROBOT | B.CHECK_INTRA
        for b in range(batches):
            g = torch.Generator(device="cuda")
            g.manual_seed(1000 * (b + 1) + len(name))
            q = lo_t + (hi_t - lo_t) * torch.rand((n, desc.nvar), generator=g, device="cuda", dtype=torch.float32)
            for i, m in enumerate(desc.mimic):
                if m is not None:
                    src = desc.var_index[desc.joint_names.index(m[0])]
                    q[:, desc.var_index[i]] = m[1] * q[:, src] + m[2]
            q = q.contiguous()
            v = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
            # 4 = the one-CTA-per-state kernel (what small batches run on), forced at this batch size
            for k, prec in enumerate((B.PRECISION_EXACT, B.PRECISION_FAST, 3, 4)):
                e.group.check_device(q.data_ptr(), B.Q_F32, n, v[k].data_ptr(), None if self_only else e.world, None, flags,
                                     prec, stream)
            torch.cuda.synchronize()
            diff = int((v[0] != v[1]).sum().item()) + int((v[0] != v[3]).sum().item())
            und = float((v[2] == 2).sum().item()) / n
            wrong_decided = int(((v[2] != 2) & (v[2] != v[0])).sum().item())
            out["batches"].append({"case": name, "seed": 1000 * (b + 1) + len(name), "states": n, "fast_or_small_ne_exact": diff,
                                   "screening_undecided": und, "screening_decided_wrong": wrong_decided,
                                   "colliding": float(v[0].float().mean().item())})
            out["states"] += n
            out["differences"] += diff + wrong_decided
    print(json.dumps(out))
    return 0 if out["differences"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 10, int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000))
