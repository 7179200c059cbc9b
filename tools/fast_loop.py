"""Config 2 (Panda, Field A, robot + intra-group) through the device-resident verdict entry a few times: the command the
FAST kernel's ncu captures and occupancy experiments run.  Usage: python tools/fast_loop.py [states] [reps]"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    fl = int(os.environ.get("FAST_LOOP_FLAGS", B.CHECK_ROBOT | B.CHECK_INTRA))  # 1 robot, 4 intra-group
    q = desc.sample_states(n, scenarios.SEED_STATES, np.float32)
    qd = torch.from_numpy(q).cuda()
    vd = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: B.check(B.load().b200df_check_batch_device(
        e.group.h, e.world.h, None, C.c_void_p(qd.data_ptr()), B.Q_F32, n, fl, int(os.environ.get("FAST_LOOP_PRECISION", B.PRECISION_FAST)),
        C.c_void_p(vd.data_ptr()), C.c_void_p(st)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("fast_loop ok: %d states, %.3f ms per call, %.4e states/s, %d colliding" % (
        n, ms, n / (ms * 1e-3), int(vd.sum().item())))


if __name__ == "__main__":
    main()
