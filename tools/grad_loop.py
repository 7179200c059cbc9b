"""CHOMP configuration (1000 trajectories x 100 waypoints) through the device-resident gradient entry, a few times:
the command ncu captures gradients_gather_kernel on."""
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
    desc = rd.panda()
    e = EngineSetup(desc, scenarios.FIELD_A, scenarios.random_scene())
    fl = B.CHECK_ROBOT | B.CHECK_INTRA
    traj = scenarios.chomp_trajectories(desc, 1000, 100).reshape(-1, desc.nvar)
    N, S = len(traj), int(e.group.n_spheres)
    qd = torch.from_numpy(traj).cuda()
    dd = torch.empty((N, S), dtype=torch.float64, device="cuda")
    gd = torch.empty((N, S, 3), dtype=torch.float64, device="cuda")
    td = torch.empty((N, S), dtype=torch.uint8, device="cuda")
    cd = torch.empty((N, S, 3), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: B.check(B.load().b200df_gradients_batch_device(
        e.group.h, e.world.h, None, C.c_void_p(qd.data_ptr()), B.Q_F64, N, fl, C.c_void_p(dd.data_ptr()),
        C.c_void_p(gd.data_ptr()), C.c_void_p(td.data_ptr()), C.c_void_p(cd.data_ptr()), C.c_void_p(st)))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        run()
    b.record()
    torch.cuda.synchronize()
    print("grad_loop ok: %d states x %d spheres, %.3f ms per call" % (N, S, a.elapsed_time(b) / 10))


if __name__ == "__main__":
    main()
