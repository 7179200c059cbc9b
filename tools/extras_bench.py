#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 1, 4, 5) timed through the host-buffer C ABI, with the CPU oracle
beside them on a bounded sample.  Not the headline metric; bench.py embeds the result under "extras"."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(with_oracle=True):
    from moveit_b200 import binding as B
    from moveit_b200 import engine, scenarios
    from moveit_b200 import robot_description as rd
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import EngineSetup, OracleSetup
    out = {}
    oracle = None
    if with_oracle:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as oracle
    cores = (oracle.hardware_concurrency() if oracle else 0) or os.cpu_count() or 1

    def best_of(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best

    # config 1: 10 k Panda states, checkRobotCollision
    desc = rd.panda()
    objs = scenarios.random_scene()
    e = EngineSetup(desc, scenarios.FIELD_A, objs)
    q = desc.sample_states(10000, scenarios.SEED_STATES, np.float32)
    e.group.check(q, e.world, flags=B.CHECK_ROBOT, precision=B.PRECISION_FAST)
    t = best_of(lambda: e.group.check(q, e.world, flags=B.CHECK_ROBOT, precision=B.PRECISION_FAST), 5)
    out["config1_10k_robot_ms"] = 1e3 * t

    # checkRobotCollision alone (no intra-group pairs) on device-resident states: the kernel variant without the pair code
    try:
        import torch
        nq = 4_000_000
        qd = torch.from_numpy(desc.sample_states(nq, 777, np.float32)).cuda()
        vd = torch.empty(nq, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        res = {}
        for name, fl_ in (("robot_only", B.CHECK_ROBOT), ("robot_and_intra", B.CHECK_ROBOT | B.CHECK_INTRA)):
            run = lambda: e.group.check_device(qd.data_ptr(), B.Q_F32, nq, vd.data_ptr(), e.world, flags=fl_,
                                               precision=B.PRECISION_FAST, stream=st)
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                run()
            e1.record()
            torch.cuda.synchronize()
            res[name + "_states_per_s"] = nq * 5 / (e0.elapsed_time(e1) * 1e-3)
        out["fast_device_resident_4M"] = res
    except ImportError:
        pass

    # what a C++ caller with RobotState doubles in ordinary (pageable) vectors sees: 2 M states, fp64, FAST
    qp = desc.sample_states(2_000_000, 4242, np.float64)
    flp = B.CHECK_ROBOT | B.CHECK_INTRA
    e.group.check(qp[:100000], e.world, flags=flp, precision=B.PRECISION_FAST)
    tpg = best_of(lambda: e.group.check(qp, e.world, flags=flp, precision=B.PRECISION_FAST), 3)
    out["check_2M_f64_pageable"] = {"ms": 1e3 * tpg, "states_per_s": len(qp) / tpg, "h2d_bytes": int(qp.nbytes)}

    # one state per call, the way OMPL's StateValidityChecker::isValid drives a CollisionEnv (latency, not throughput)
    fl1 = B.CHECK_ROBOT | B.CHECK_INTRA
    one32, one64 = q[:1].copy(), q[:1].astype(np.float64)
    for name, arr, prec in (("f32_fast", one32, B.PRECISION_FAST), ("f64_exact", one64, B.PRECISION_EXACT)):
        for _ in range(20):
            e.group.check(arr, e.world, flags=fl1, precision=prec)
        t0 = time.perf_counter()
        for _ in range(500):
            e.group.check(arr, e.world, flags=fl1, precision=prec)
        out["single_state_call_us_" + name] = 1e6 * (time.perf_counter() - t0) / 500

    # config 5: CHOMP gradients, 1000 trajectories x 100 waypoints
    traj = scenarios.chomp_trajectories(desc, 1000, 100).reshape(-1, desc.nvar)
    fl = B.CHECK_ROBOT | B.CHECK_INTRA
    e.group.gradients(traj[:1000], e.world, flags=fl)
    Sg = int(e.group.n_spheres)
    # result arrays in ordinary (pageable) memory, allocated and touched once like a caller's reused buffers
    pg = dict(dist=np.zeros((len(traj), Sg)), grad=np.zeros((len(traj), Sg, 3)),
              type=np.zeros((len(traj), Sg), dtype=np.uint8), centers=np.zeros((len(traj), Sg, 3)))
    t = best_of(lambda: e.group.gradients(traj, e.world, flags=fl, out=pg), 3)
    out["config5_chomp_gradients"] = {"states": len(traj), "ms": 1e3 * t, "states_per_s": len(traj) / t,
                                      "spheres": int(e.group.n_spheres),
                                      "d2h_bytes": len(traj) * int(e.group.n_spheres) * (8 + 24 + 1 + 24)}
    try:  # the same with pinned result buffers, and the device-resident entry point
        import torch
        N, S = len(traj), int(e.group.n_spheres)
        pin = dict(dist=torch.empty((N, S), dtype=torch.float64).pin_memory().numpy(),
                   grad=torch.empty((N, S, 3), dtype=torch.float64).pin_memory().numpy(),
                   type=torch.empty((N, S), dtype=torch.uint8).pin_memory().numpy(),
                   centers=torch.empty((N, S, 3), dtype=torch.float64).pin_memory().numpy())
        tq = torch.from_numpy(traj).pin_memory().numpy()
        e.group.gradients(tq, e.world, flags=fl, out=pin)
        t = best_of(lambda: e.group.gradients(tq, e.world, flags=fl, out=pin), 3)
        out["config5_chomp_gradients"]["pinned_ms"] = 1e3 * t
        out["config5_chomp_gradients"]["pinned_states_per_s"] = N / t
        import ctypes as C
        qd = torch.from_numpy(traj).cuda()
        dd = torch.empty((N, S), dtype=torch.float64, device="cuda")
        gd = torch.empty((N, S, 3), dtype=torch.float64, device="cuda")
        td = torch.empty((N, S), dtype=torch.uint8, device="cuda")
        cd = torch.empty((N, S, 3), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream()

        def dev():
            B.check(B.load().b200df_gradients_batch_device(e.group.h, e.world.h, None, C.c_void_p(qd.data_ptr()), B.Q_F64,
                                                           N, fl, C.c_void_p(dd.data_ptr()), C.c_void_p(gd.data_ptr()),
                                                           C.c_void_p(td.data_ptr()), C.c_void_p(cd.data_ptr()),
                                                           C.c_void_p(st.cuda_stream)))
        dev()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dev()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out["config5_chomp_gradients"]["device_resident_ms"] = ms
        out["config5_chomp_gradients"]["device_resident_states_per_s"] = N / (ms * 1e-3)
        assert np.array_equal(dd.cpu().numpy(), pin["dist"]) and np.array_equal(td.cpu().numpy(), pin["type"])
        # single-precision outputs (b200df_gradients_batch_f32): 29 instead of 57 bytes per sphere over PCIe
        pin32 = dict(dist=torch.empty((N, S), dtype=torch.float32).pin_memory().numpy(),
                     grad=torch.empty((N, S, 3), dtype=torch.float32).pin_memory().numpy(),
                     type=torch.empty((N, S), dtype=torch.uint8).pin_memory().numpy(),
                     centers=torch.empty((N, S, 3), dtype=torch.float32).pin_memory().numpy())
        e.group.gradients(tq, e.world, flags=fl, out=pin32, out_dtype=np.float32)
        t = best_of(lambda: e.group.gradients(tq, e.world, flags=fl, out=pin32, out_dtype=np.float32), 3)
        c5 = out["config5_chomp_gradients"]
        c5["f32_out_pinned_ms"] = 1e3 * t
        c5["f32_out_d2h_bytes"] = N * S * (4 + 12 + 1 + 12)
        pg32 = dict(dist=np.zeros((N, S), np.float32), grad=np.zeros((N, S, 3), np.float32),
                    type=np.zeros((N, S), np.uint8), centers=np.zeros((N, S, 3), np.float32))
        t = best_of(lambda: e.group.gradients(traj, e.world, flags=fl, out=pg32, out_dtype=np.float32), 3)
        c5["f32_out_pageable_ms"] = 1e3 * t
        assert np.array_equal(pg32["dist"], pin32["dist"]) and np.array_equal(pg32["type"], pin32["type"]) and \
            np.array_equal(pg["grad"], pin["grad"]) and np.array_equal(pg["centers"], pin["centers"])
        fin = pin["dist"] < 1e300
        c5["f32_out_max_abs_diff_vs_f64"] = float(max(np.max(np.abs(pin32["dist"][fin] - pin["dist"][fin])),
                                                       np.max(np.abs(pin32["grad"] - pin["grad"])),
                                                       np.max(np.abs(pin32["centers"] - pin["centers"]))))
        assert np.array_equal(pin32["type"], pin["type"]) and c5["f32_out_max_abs_diff_vs_f64"] < 1e-5
        d32 = torch.empty((N, S), dtype=torch.float32, device="cuda")
        g32 = torch.empty((N, S, 3), dtype=torch.float32, device="cuda")
        c32 = torch.empty((N, S, 3), dtype=torch.float32, device="cuda")

        def dev32():
            B.check(B.load().b200df_gradients_batch_device_f32(
                e.group.h, e.world.h, None, C.c_void_p(qd.data_ptr()), B.Q_F64, N, fl, C.c_void_p(d32.data_ptr()),
                C.c_void_p(g32.data_ptr()), C.c_void_p(td.data_ptr()), C.c_void_p(c32.data_ptr()),
                C.c_void_p(st.cuda_stream)))
        dev32()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            dev32()
        e1.record()
        torch.cuda.synchronize()
        c5["f32_out_device_resident_ms"] = e0.elapsed_time(e1) / 5
    except ImportError:
        pass

    # CollisionEnvB200::checkCollisionBatch (the C++ env of the plugin) at scale, from pageable fp64 rows
    try:
        import subprocess
        import tempfile
        from moveit_b200 import scenario_io
        import runpy
        exe = runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "moveit_b200",
                                          "host", "build_host.py"))["build"]()
        with tempfile.TemporaryDirectory() as td:
            scen, states, prefix = os.path.join(td, "scen.txt"), os.path.join(td, "q.f64"), os.path.join(td, "out")
            scenario_io.write_scenario(scen, desc, scenarios.FIELD_A, scenarios.random_scene())
            nb = 4_000_000
            qb = desc.sample_states(nb, scenarios.SEED_STATES, np.float32).astype(np.float64)
            qb.tofile(states)
            r = subprocess.run([exe, scen, states, prefix], capture_output=True, text=True,
                               env=dict(os.environ, B200DF_DRIVER_BENCH_REPS="5"))
            line = [l for l in r.stdout.splitlines() if l.startswith("bench_batch")]
            if r.returncode == 0 and line:
                kv = dict(t.split("=") for t in line[0].split()[1:])
                got = np.fromfile(prefix + ".all.u8", dtype=np.uint8)
                ref = e.group.check(qb, e.world, flags=B.CHECK_ROBOT | B.CHECK_INTRA, precision=B.PRECISION_FAST)
                out["cpp_env_checkCollisionBatch"] = {
                    "states": nb, "states_per_s": float(kv["states_per_s"]), "ms": 1e3 * float(kv["best_s"]),
                    "input": "pageable fp64 rows through CollisionEnvB200::checkCollisionBatch (host_driver)",
                    "verdicts_match_c_abi_call": bool(np.array_equal(got, ref)), "errors": int(kv["errors"])}
            else:
                out["cpp_env_checkCollisionBatch"] = {"failed": (r.stderr or r.stdout)[-300:]}
    except Exception as ex:  # the extras never take the bench line down
        out["cpp_env_checkCollisionBatch"] = {"failed": repr(ex)[:300]}

    # small calls, the granularity the reference's callers use (tools/latency_probe.py has the full sweep)
    small = {}
    q256 = q[:256].copy()
    t100 = traj[:100].copy()
    qa1, qb1 = q[:1].astype(np.float64), q[1:2].astype(np.float64)
    for name, fn in (("check_256_states", lambda: e.group.check(q256, e.world, flags=fl1, precision=B.PRECISION_FAST)),
                     ("gradients_100_waypoints", lambda: e.group.gradients(t100, e.world, flags=fl)),
                     ("edge_1_motion_16_segments",
                      lambda: e.group.check_edges(qa1, qb1, 16, e.world, flags=fl, precision=B.PRECISION_FAST))):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        for _ in range(300):
            fn()
        small[name + "_us"] = 1e6 * (time.perf_counter() - t0) / 300
    out["small_call_latency_incl_python_binding"] = small

    # planner edge checks: 1 M random motions x 10 segments, only the end states cross PCIe
    M, K = 1_000_000, 10
    qa = desc.sample_states(M, 901)
    qb = desc.sample_states(M, 902)
    e.group.check_edges(qa[:1000], qb[:1000], K, e.world, flags=fl, precision=B.PRECISION_FAST)
    t = best_of(lambda: e.group.check_edges(qa, qb, K, e.world, flags=fl, precision=B.PRECISION_FAST), 3)
    out["edge_checks"] = {"motions": M, "segments": K, "ms": 1e3 * t, "motions_per_s": M / t,
                          "states_per_s": M * K / t, "h2d_bytes": 2 * M * desc.nvar * 8, "d2h_bytes": 4 * M}
    try:  # the same with the end points in page-locked memory (the 144 MB H2D then runs at PCIe speed)
        import torch
        pa = torch.from_numpy(qa).pin_memory().numpy()
        pb = torch.from_numpy(qb).pin_memory().numpy()
        tp = best_of(lambda: e.group.check_edges(pa, pb, K, e.world, flags=fl, precision=B.PRECISION_FAST), 3)
        out["edge_checks"]["pinned_ms"] = 1e3 * tp
        out["edge_checks"]["pinned_states_per_s"] = M * K / tp
    except ImportError:
        pass

    # config 4: PR2-like dual-arm tree, full ACM
    pr2 = rd.pr2_like()
    e4 = EngineSetup(pr2, scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=31))
    q4 = pr2.sample_states(200_000, 12, np.float32)
    e4.group.check(q4[:1000], e4.world, flags=fl, precision=B.PRECISION_FAST)
    t = best_of(lambda: e4.group.check(q4, e4.world, flags=fl, precision=B.PRECISION_FAST), 3)
    tx = best_of(lambda: e4.group.check(q4, e4.world, flags=fl, precision=B.PRECISION_EXACT), 3)
    out["config4_pr2_like"] = {"states": len(q4), "ms": 1e3 * t, "states_per_s": len(q4) / t,
                               "exact_states_per_s": len(q4) / tx,
                               "links": pr2.n_links, "spheres": int(e4.group.n_spheres)}
    # the self-collision part on its own (the "full ACM self-collision pairs" of BASELINE config 4), 1 M states: random
    # states of this tree nearly always touch the scene, the pair tests are where its verdicts are not a foregone
    # conclusion (adjacent and always-colliding pairs disabled like an SRDF does)
    q4b = pr2.sample_states(1_000_000, 13, np.float32)
    v4 = e4.group.check(q4b, None, flags=B.CHECK_INTRA, precision=B.PRECISION_FAST)
    ts = best_of(lambda: e4.group.check(q4b, None, flags=B.CHECK_INTRA, precision=B.PRECISION_FAST), 3)
    out["config4_pr2_like"].update({"self_collision_states": len(q4b), "self_collision_states_per_s": len(q4b) / ts,
                                    "self_collision_colliding_fraction": float(v4.mean()),
                                    "robot_and_intra_colliding_fraction": float(e4.group.check(
                                        q4, e4.world, flags=fl, precision=B.PRECISION_FAST).mean())})
    if oracle:
        o = OracleSetup(oracle, desc, scenarios.FIELD_A, objs, self_state=np.zeros(desc.nvar))
        sample = traj[:20000]
        t0 = time.perf_counter()
        o.env.gradients_batch(sample, flags=2 | 4)
        out["config5_chomp_gradients"]["cpu_oracle_states_per_s_1_thread"] = len(sample) / (time.perf_counter() - t0)
        o4 = OracleSetup(oracle, pr2, scenarios.FIELD_DEFAULT, scenarios.random_scene(seed=31),
                         self_state=np.zeros(pr2.nvar))
        s4 = q4[:20000].astype(np.float64)
        _, tt = o4.env.check_batch(s4, mode=3, n_threads=cores, return_time=True)
        out["config4_pr2_like"]["cpu_oracle_states_per_s"] = len(s4) / tt
        out["config4_pr2_like"]["cpu_cores"] = cores
    return out


if __name__ == "__main__":
    print(json.dumps(run("--no-oracle" not in sys.argv)))
