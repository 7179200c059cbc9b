"""Call latency of b200df_check_batch for small batches (the per-state isValid() usage of the reference's callers).
Times the bare C-ABI call (preallocated buffers, no numpy work per call) and, beside it, the kernel alone with CUDA
events on device-resident inputs, with the small-batch routes on and off (B200DF_SMALL_MAX_STATES / B200DF_SMALL_BATCH)."""
import ctypes as C
import json
import os
import sys
import time

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
    lib = B.load()
    flags = B.CHECK_ROBOT | B.CHECK_INTRA
    out = {"host_call_us": {}, "host_call_us_small_paths_off": {}, "kernel_us": {}}
    q_all = desc.sample_states(1 << 17, scenarios.SEED_STATES, np.float32)

    def host_calls(key):
        for n in (1, 32, 256, 2048, 4096, 8192):
            for name, arr, dt, prec in (("f32_fast", q_all[:n].copy(), B.Q_F32, B.PRECISION_FAST),
                                        ("f64_exact", q_all[:n].astype(np.float64), B.Q_F64, B.PRECISION_EXACT)):
                v = np.empty(n, dtype=np.uint8)
                qp, vp = arr.ctypes.data_as(C.c_void_p), B.ptr(v, C.c_uint8)
                call = lambda: lib.b200df_check_batch(e.group.h, e.world.h, None, qp, dt, n, flags, prec, vp)
                for _ in range(50):
                    assert call() == 0
                reps = 2000 if n <= 256 else 500
                t0 = time.perf_counter()
                for _ in range(reps):
                    call()
                out[key]["%s_n%d" % (name, n)] = round(1e6 * (time.perf_counter() - t0) / reps, 2)

    host_calls("host_call_us")
    os.environ["B200DF_SMALL_MAX_STATES"] = "0"
    os.environ["B200DF_SMALL_BATCH"] = "0"
    host_calls("host_call_us_small_paths_off")

    # kernels alone on device-resident fp32 joint values (small-batch routing still off: precision picks the kernel)
    st = torch.cuda.current_stream().cuda_stream
    for n in (1, 32, 256, 2048, 8192, 32768, 131072):
        qd = torch.from_numpy(q_all[:n].copy()).cuda()
        vd = torch.empty(n, dtype=torch.uint8, device="cuda")
        for name, prec in (("small", 4), ("batched_exact", B.PRECISION_EXACT), ("batched_fast", B.PRECISION_FAST)):
            run = lambda: e.group.check_device(qd.data_ptr(), B.Q_F32, n, vd.data_ptr(), e.world, flags=flags,
                                               precision=prec, stream=st)
            for _ in range(20):
                run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200 if n <= 8192 else 50
            a.record()
            for _ in range(reps):
                run()
            b.record()
            torch.cuda.synchronize()
            out["kernel_us"]["%s_n%d" % (name, n)] = round(1e3 * a.elapsed_time(b) / reps, 2)

    # the other per-call entry points a planner drives with small inputs: CHOMP evaluates the gradients of one
    # trajectory (~100 waypoints) per iteration, OMPL's MotionValidator checks one edge per call
    os.environ.pop("B200DF_SMALL_MAX_STATES")
    os.environ.pop("B200DF_SMALL_BATCH")
    out["gradients_host_call_us"], out["gradients_kernel_us"], out["edges_host_call_us"] = {}, {}, {}
    S = int(e.group.n_spheres)
    for n in (1, 100, 1000):
        qn = q_all[:n].astype(np.float64)
        bufs = dict(dist=np.empty((n, S)), grad=np.empty((n, S, 3)), type=np.empty((n, S), dtype=np.uint8),
                    centers=np.empty((n, S, 3)))
        call = lambda: e.group.gradients(qn, e.world, flags=flags, out=bufs)
        for _ in range(20):
            call()
        t0 = time.perf_counter()
        for _ in range(300):
            call()
        out["gradients_host_call_us"]["n%d" % n] = round(1e6 * (time.perf_counter() - t0) / 300, 2)
        qd = torch.from_numpy(qn).cuda()
        dd = torch.empty((n, S), dtype=torch.float64, device="cuda")
        gd = torch.empty((n, S, 3), dtype=torch.float64, device="cuda")
        td = torch.empty((n, S), dtype=torch.uint8, device="cuda")
        cd = torch.empty((n, S, 3), dtype=torch.float64, device="cuda")
        run = lambda: B.check(lib.b200df_gradients_batch_device(
            e.group.h, e.world.h, None, C.c_void_p(qd.data_ptr()), B.Q_F64, n, flags, C.c_void_p(dd.data_ptr()),
            C.c_void_p(gd.data_ptr()), C.c_void_p(td.data_ptr()), C.c_void_p(cd.data_ptr()), C.c_void_p(st)))
        for _ in range(20):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200):
            run()
        b.record()
        torch.cuda.synchronize()
        out["gradients_kernel_us"]["n%d" % n] = round(1e3 * a.elapsed_time(b) / 200, 2)
    for m, seg in ((1, 16), (1, 64), (64, 16)):
        qa, qb = q_all[:m].astype(np.float64), q_all[m:2 * m].astype(np.float64)
        call = lambda: e.group.check_edges(qa, qb, seg, e.world, flags=flags, precision=B.PRECISION_FAST)
        for _ in range(20):
            call()
        t0 = time.perf_counter()
        for _ in range(300):
            call()
        out["edges_host_call_us"]["m%d_seg%d" % (m, seg)] = round(1e6 * (time.perf_counter() - t0) / 300, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
