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
    print(json.dumps(out))


if __name__ == "__main__":
    main()
