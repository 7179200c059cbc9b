"""Config 4 (PR2-like tree, 33 links / 124 spheres / ~500 link pairs, robot + intra-group) device-resident: FAST and
EXACT rates for a batch size given on the command line, with the route the call takes."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch
    from moveit_b200 import binding as B
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios
    from common import EngineSetup
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    pr2 = rd.pr2_like()
    e = EngineSetup(pr2, scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=31))
    fl = int(os.environ.get("PR2_LOOP_FLAGS", B.CHECK_ROBOT | B.CHECK_INTRA))  # 1 robot, 4 intra-group
    q = pr2.sample_states(n, 12, np.float32)
    qd = torch.from_numpy(q).cuda()
    vd = torch.empty(n, dtype=torch.uint8, device="cuda")
    ve = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for name, prec, out in (("fast", B.PRECISION_FAST, vd), ("exact", B.PRECISION_EXACT, ve)):
        run = lambda: B.check(B.load().b200df_check_batch_device(
            e.group.h, e.world.h, None, C.c_void_p(qd.data_ptr()), B.Q_F32, n, fl, prec, C.c_void_p(out.data_ptr()),
            C.c_void_p(st)))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print("pr2_loop %s: %d states, %.3f ms, %.4e states/s, %d colliding" % (name, n, ms, n / (ms * 1e-3),
                                                                                int(out.sum().item())))
    assert torch.equal(vd, ve)


if __name__ == "__main__":
    main()
