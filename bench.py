#!/usr/bin/env python
"""bench.py — state-validity checks/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 2 — Panda (12 links, 9 variables, 60 collision spheres), robot-vs-world
against Field A (2 m cube @ 2 cm, 8 boxes + 4 cylinders) plus intra-group sphere pairs from the SRDF-style ACM,
10 M uniformly random states PER GPU (weak scaling; the state batch shards with no data-path collective).  A step is
one pass of the hot path over the rank's batch.

  value : whole-job states/s, inputs already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e   : the same metric through the host-buffer C-ABI call (b200df_check_batch): pinned host joint values in,
          host verdicts out, copies inside the timed region
  cpu_baseline / --impl reference : the CPU oracle (faithful restatement of the reference; the reference itself
          cannot be compiled here) with one thread per host core on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STATES_PER_GPU = 10_000_000
ALGO_BYTES_PER_STATE = 4 * 9 + 1  # fp32 joint values in, one verdict byte out


def ncu_counters(path, precision):
    """Counters of the dominant kernel from the committed ncu --set full capture of this command
    (profiles/summarize.py traffic): DRAM bytes, L2 sectors and warp instructions per state.  None when the file is
    missing - the bench then reports traffic = null rather than a stale constant."""
    try:
        with open(path) as f:
            d = json.load(f)
        k = d.get(precision)
        if not k or not k.get("states_per_launch"):
            return None
        n = float(k["states_per_launch"])
        return {"dram_bytes_per_state": (k["dram_bytes_read"] + k["dram_bytes_write"]) / n,
                "l2_sector_bytes_per_state": 32.0 * (k.get("lts_sectors_read", 0) + k.get("lts_sectors_write", 0)) / n,
                "warp_instructions_per_state": k.get("warp_instructions", 0) / n, "kernel": k.get("kernel"),
                "passes": [{"kernel": p.get("kernel", "")[:110], "share_of_call": p["duration_us_under_ncu"] /
                            max(k.get("duration_us_under_ncu", 0.0), 1e-9),
                            "issue_active_pct": p.get("issue_active_pct"), "warps_active_pct": p.get("warps_active_pct"),
                            "registers": p.get("registers")} for p in k.get("passes", [])],
                "source": os.path.relpath(path, ROOT) + " <- " + str(d.get("report"))}
    except Exception:
        return None


def issue_roofline(counters, states_per_s_per_gpu, clocks):
    """The resource that actually binds the validity kernel: warp-instruction issue slots.  achieved = (warp
    instructions per state, counted by ncu on the committed capture) x (states/s measured in this run on one GPU);
    peak = 148 SMs x 4 schedulers x the SM clock sampled during the timed region."""
    per_state = (counters or {}).get("warp_instructions_per_state")
    mhz = (clocks or {}).get("sm_mhz")
    if not per_state or not mhz:
        return None
    peak = 148 * 4 * mhz * 1e6
    achieved = per_state * states_per_s_per_gpu
    return {"achieved": achieved, "peak": peak, "unit": "warp-instructions/s", "frac": achieved / peak,
            "warp_instructions_per_state": per_state, "source": counters["source"] + " x live rate"}


def l2_roofline(counters, states_per_s_per_gpu, clocks):
    """north_star asks for the lookup kernels' achieved L2 GB/s against the chip's peak: L2 sectors x 32 B per state
    (ncu lts__t_sectors of the committed capture) x the live rate, against the L2 throughput ceiling the microarchitecture
    guide measured (~6300 B per clock for the whole chip) at the SM clock sampled in this run."""
    per_state = (counters or {}).get("l2_sector_bytes_per_state")
    mhz = (clocks or {}).get("sm_mhz")
    if not per_state or not mhz:
        return None
    peak = 6300.0 * mhz * 1e6 / 1e9
    achieved = per_state * states_per_s_per_gpu / 1e9
    return {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "l2_bytes_per_state": per_state,
            "peak_source": "B300_MICROARCH.md LTS cap 6300 B/clk x sampled SM clock (guide figure, not measured here)",
            "source": counters["source"] + " x live rate"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--states", type=int, default=STATES_PER_GPU, help="states per GPU per step")
    ap.add_argument("--precision", default="fast", choices=["fast", "exact", "v1"],
                    help="fast = fp32 screening + exact fp64 re-run of the undecided states (same verdicts as exact)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-edt", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (BASELINE configs 1, 4, 5)")
    ap.add_argument("--single-process", action="store_true",
                    help="drive --gpus N devices from ONE process through b200df_multi (how a MoveIt process would) "
                         "instead of one torchrun rank per GPU")
    ap.add_argument("--traffic-json", default=os.path.join(ROOT, "profiles", "r02_validity_traffic.json"),
                    help="per-launch DRAM / L2 counters of the dominant kernel from an ncu --set full capture of this "
                         "command (written by `python profiles/summarize.py traffic <rep> <json>`)")
    return ap.parse_args()


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed regions: NVML in-process every 5 ms (nvidia_ml_py),
    falling back to polling nvidia-smi when NVML is not importable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []  # (sm_mhz, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.source = "nvml"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            self.source = "nvidia-smi"

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                pass
        return index

    def _sample_nvml(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        self.samples.append((mhz, bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        p = [x.strip() for x in out.strip().split(",")]
        if len(p) >= 7:
            act = [x.lower().startswith("active") for x in p[3:7]]
            self.samples.append((float(p[0]), act[0], act[1], act[2], act[3]))
            self.max_mhz = float(p[1])

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.005 if self.nv is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "samples": 0}
        sm = sorted(s[0] for s in self.samples)
        reasons = [name for k, name in ((1, "hw_slowdown"), (2, "hw_thermal_slowdown"), (3, "sw_thermal_slowdown"),
                                        (4, "sw_power_cap")) if any(s[k] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples), "source": self.source}


def build_oracle_env(n_threads_hint=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py
    from common import OracleSetup
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios
    desc = rd.panda()
    setup = OracleSetup(oracle_py, desc, scenarios.FIELD_A, scenarios.random_scene(),
                        self_state=np.zeros(desc.nvar), exact_field=False)
    return oracle_py, desc, setup


def cpu_rate(setup, desc, oracle_py, target_seconds=12.0):
    """states/s of the oracle with one thread per host core on a bounded sample of the bench workload: the first
    4 M states of the bench batch, passed over repeatedly until about target_seconds of wall time are spent."""
    from moveit_b200 import scenarios
    cores = oracle_py.hardware_concurrency() or os.cpu_count() or 1
    n = 4_000_000
    q = desc.sample_states(n, scenarios.SEED_STATES, np.float32).astype(np.float64)
    setup.env.check_batch(q[:200000], mode=3, n_threads=cores)  # warm the caches / thread pool
    total_t, passes = 0.0, 0
    while total_t < target_seconds and passes < 64:
        _, t = setup.env.check_batch(q, mode=3, n_threads=cores, return_time=True)
        total_t += t
        passes += 1
    return n * passes / total_t, cores, n * passes, total_t


def run_reference(args, rank, world):
    if rank != 0:
        return
    oracle_py, desc, setup = build_oracle_env()
    from moveit_b200 import scenarios
    cores = oracle_py.hardware_concurrency() or os.cpu_count() or 1
    probe = desc.sample_states(20000, scenarios.SEED_STATES, np.float32).astype(np.float64)
    _, t = setup.env.check_batch(probe, mode=3, n_threads=cores, return_time=True)
    per_step = int(min(max(len(probe) / t * 3.0, 20000), 2_000_000))  # ~3 s of CPU work per step
    q = desc.sample_states(per_step, scenarios.SEED_STATES, np.float32).astype(np.float64)
    for _ in range(args.warmup):
        setup.env.check_batch(q, mode=3, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        setup.env.check_batch(q, mode=3, n_threads=cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "state_validity_checks_per_sec", "value": value, "unit": "states/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64" if args.precision == "fast" else "f64", "data": "synthetic",
        "config": workload_config(per_step),
        "cpu_baseline": {"value": value, "unit": "states/s", "cores": cores, "kind": "port",
                         "sample": "%d states per step of the bench batch (same seed), CPU oracle = faithful restatement "
                                   "of CollisionEnvDistanceField (reference not buildable here), one thread per core"
                                   % per_step},
        "e2e": {"value": value, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(states_per_gpu):
    return {"workload": "BASELINE config 2: Panda (12 links, 9 vars, 60 spheres) checkRobotCollision vs Field A "
                        "(100^3 voxels @ 2 cm, 8 boxes + 4 cylinders) + intra-group sphere pairs (SRDF ACM), "
                        "uniform random states",
            "states_per_gpu_per_step": states_per_gpu, "checks": "robot+intra", "q_dtype": "f32",
            "l2_policy": "inputs larger than L2 (360 MB of joint values per step per GPU)",
            "parallelism": "states sharded across ranks, field replicated once by NCCL broadcast"}


def edt_cpu_baseline(scenarios):
    """The reference's own (single-threaded) propagation on Field A through the oracle's restatement:
    part of the cpu_baseline leg.  Field B takes ~20 s the same way and is left out of the default
    run."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as oracle
    from common import oracle_world_points
    fa = scenarios.FIELD_A
    corner = [fa["center"][i] - 0.5 * fa["size"][i] for i in range(3)]
    pts = oracle_world_points(oracle, scenarios.random_scene(), fa["resolution"])
    best = 1e9
    for _ in range(3):
        f = oracle.Field(fa["size"], fa["resolution"], corner, fa["max_distance"])
        t0 = time.perf_counter()
        f.add_points(pts)
        best = min(best, time.perf_counter() - t0)
    return {"field_a_reference_propagation_ms": 1e3 * best, "cores": 1, "kind": "port",
            "sample": "Field A, the bench scene's 12 objects in one addPointsToField call (one propagatePositive sweep)"}


def edt_timings(engine, scenarios, B):
    """EDT rebuild ms (BASELINE metric's second half): Field A from the bench scene, Field B from the 1 M-point cloud."""
    out = {}
    fa = scenarios.FIELD_A
    f = engine.Field.centered(fa["size"], fa["resolution"], fa["center"], fa["max_distance"])
    best = 1e9
    objs = scenarios.random_scene()
    for _ in range(5):
        f.reset()
        for t, dims, pose in objs:
            f.add_posed_body(t, dims, pose)
        best = min(best, f.rebuild())
    out["field_a_2cm_100^3_ms"] = best
    fb = scenarios.FIELD_B
    pts = scenarios.synthetic_cloud(1_000_000)
    g = engine.Field.centered(fb["size"], fb["resolution"], fb["center"], fb["max_distance"])
    best = 1e9
    for _ in range(3):
        g.reset()
        g.add_points(pts)
        best = min(best, g.rebuild())
    out["field_b_1cm_400^3_1Mpts_ms"] = best
    peak, src = measured_peak_hbm()
    algo = 400 ** 3 * 3  # occupancy byte in + u16 d^2 out per voxel; the point list is scattered at add time
    out["field_b_roofline"] = {"bound": "hbm", "achieved": algo / (best * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": algo / (best * 1e-3) / 1e9 / peak, "peak_source": src,
                               "algorithmic_bytes": algo,
                               "note": "2 launches per z slab (bit rows -> plane kernel: z distances on the fly + DPX "
                                       "march along y; axis kernel: DPX march along x), 4 slabs on 4 streams"}
    t0 = time.perf_counter()
    g.reset()
    g.add_points(pts)
    out["field_b_add_1Mpts_host_ms"] = 1e3 * (time.perf_counter() - t0)  # 24 MB H2D + scatter kernel, wall clock
    try:  # the same cloud already on the GPU (perception output): scatter + rebuild, device time
        import torch
        pd = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
        st = torch.cuda.current_stream()
        best_s = 1e9
        for _ in range(3):
            g.reset()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            g.add_points_device(pd.data_ptr(), len(pts), st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            best_s = min(best_s, e0.elapsed_time(e1))
            g.rebuild()
        out["field_b_scatter_1Mpts_device_ms"] = best_s
    except ImportError:
        pass
    return out


def run_single_process(args):
    """--single-process: ONE process drives --gpus N devices through b200df_multi, the way a MoveIt process (one
    PlanningScene shared by planner threads) would; no torchrun, no NCCL.  Same workload, same per-device batch
    (weak scaling), same timing rules: CUDA events on every device's stream, the slowest device counts."""
    import torch
    from moveit_b200 import binding as B
    from moveit_b200 import engine, scenarios
    from moveit_b200 import robot_description as rd
    from moveit_b200.rng import uniform01

    G = args.gpus
    if not torch.cuda.is_available() or B.device_count() < G:
        raise SystemExit("bench.py --single-process --gpus %d needs %d CUDA devices (no CPU fallback)" % (G, G))
    devices = list(range(G))
    B.check(B.load().b200df_set_device(0))
    desc = rd.panda()
    md = engine.assemble_model(desc, engine.body_decomposition, scenarios.FIELD_A["resolution"])
    links = desc.group_links("")
    se, ie = engine.resolve_acm(md, links, desc.acm_entry_matrix())
    fa = scenarios.FIELD_A
    multi = engine.MultiGroup(md, links, se, ie, devices, fa["resolution"], 0.0, fa["max_distance"])
    src = engine.Field.centered(fa["size"], fa["resolution"], fa["center"], fa["max_distance"])
    for t, dims, pose in scenarios.random_scene():
        src.add_posed_body(t, dims, pose)
    t0 = time.perf_counter()
    multi.set_field(0, src)  # rebuild on device 0 + one peer copy per replica
    replicate_first_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    multi.set_field(0, src)
    replicate_ms = 1e3 * (time.perf_counter() - t0)
    ref_d2 = src.d2()
    for slot in range(G):
        if not np.array_equal(multi.replica_d2(0, slot, src.dims), ref_d2):
            raise SystemExit("replica %d of the world field differs from the source" % slot)

    n = args.states
    lim_lo = np.array([l[0] for l in desc.limits])
    lim_hi = np.array([l[1] for l in desc.limits])
    q_host = torch.empty((n * G, desc.nvar), dtype=torch.float32).pin_memory()
    chunk = 1_000_000
    for s in range(0, n * G, chunk):
        m = min(chunk, n * G - s)
        u = uniform01(scenarios.SEED_STATES, m * desc.nvar, offset=s * desc.nvar).reshape(m, desc.nvar)
        qq = lim_lo + (lim_hi - lim_lo) * u
        qq[:, 8] = qq[:, 7]
        q_host[s:s + m] = torch.from_numpy(qq.astype(np.float32))
    v_host = torch.empty(n * G, dtype=torch.uint8).pin_memory()
    q_dev, v_dev, streams = [], [], []
    for d in devices:
        with torch.cuda.device(d):
            q_dev.append(q_host[d * n:(d + 1) * n].to("cuda:%d" % d))
            v_dev.append(torch.empty(n, dtype=torch.uint8, device="cuda:%d" % d))
            streams.append(torch.cuda.Stream(device=d))
    flags = B.CHECK_ROBOT | B.CHECK_INTRA
    precision = {"fast": B.PRECISION_FAST, "exact": B.PRECISION_EXACT, "v1": 2}[args.precision]
    counts = [n] * G

    def step_device():
        multi.check_device([q.data_ptr() for q in q_dev], B.Q_F32, counts, [v.data_ptr() for v in v_dev], flags,
                           precision, [s.cuda_stream for s in streams])

    def sync_all():
        for d in devices:
            torch.cuda.synchronize(d)

    for _ in range(max(args.warmup, 3)):
        step_device()
    sync_all()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = B.load().b200df_launch_count()
    ev = []
    for d in devices:
        with torch.cuda.device(d):
            ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
    for d in devices:
        ev[d][0].record(streams[d])
    for _ in range(args.steps):
        step_device()
    for d in devices:
        ev[d][1].record(streams[d])
    sync_all()
    per_dev_ms = [ev[d][0].elapsed_time(ev[d][1]) for d in devices]
    ms_total = max(per_dev_ms)
    launches = B.load().b200df_launch_count() - launches0

    # every device's shard against device 0 running the same shard (a corrupt replica would show here)
    agree = True
    with torch.cuda.device(0):
        probe = torch.empty(min(n, 100_000), dtype=torch.uint8, device="cuda:0")
    m = probe.numel()
    g0 = engine.Group(engine.Model(md), links, se, ie, fa["resolution"], 0.0, fa["max_distance"])
    for d in devices:
        qd = q_host[d * n:d * n + m].to("cuda:0")
        g0.check_device(qd.data_ptr(), B.Q_F32, m, probe.data_ptr(), src, None, flags, B.PRECISION_EXACT, 0)
        torch.cuda.synchronize(0)
        if not torch.equal(probe.cpu(), v_dev[d][:m].cpu()):
            agree = False
    if not agree:
        raise SystemExit("cross-device check failed: a device's verdicts differ from device 0's exact kernel")

    # end to end: ONE b200df_multi_check_batch call on the whole N x G batch in page-locked memory
    def step_host():
        B.check(B.load().b200df_multi_check_batch(multi.h, ctypes.c_void_p(q_host.data_ptr()), B.Q_F32, n * G, flags,
                                                  precision,
                                                  ctypes.cast(v_host.data_ptr(), ctypes.POINTER(ctypes.c_uint8))))

    step_host()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    sync_all()
    e2e_s = time.perf_counter() - t0
    same = all(bool(torch.equal(v_host[d * n:(d + 1) * n], v_dev[d].cpu())) for d in devices)
    clocks = sampler.summary()
    ms_per_step = ms_total / args.steps
    peak, peak_src = measured_peak_hbm()
    line = {
        "metric": "state_validity_checks_per_sec", "value": n * G / (ms_per_step * 1e-3), "unit": "states/s",
        "n_gpus": G, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64" if args.precision == "fast" else "f64", "data": "synthetic",
        "config": dict(workload_config(n), parallelism="ONE process, b200df_multi: states sharded across devices, "
                                                        "field replicated by peer copies (no NCCL, no torchrun)"),
        "launch": "single-process",
        "e2e": {"value": n * G * args.steps / e2e_s, "unit": "states/s", "h2d_bytes_per_step": n * G * desc.nvar * 4,
                "d2h_bytes_per_step": n * G, "steps": args.steps,
                "input": "one page-locked fp32 buffer of all N x G states, one b200df_multi_check_batch call per step",
                "verdicts_match_device_path": same},
        "gpu_launches": int(launches), "clocks": clocks, "per_device_ms_per_step": [x / args.steps for x in per_dev_ms],
        "roofline": {"bound": "hbm", "achieved": ALGO_BYTES_PER_STATE * n / (ms_per_step * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": ALGO_BYTES_PER_STATE * n / (ms_per_step * 1e-3) / 1e9 / peak,
                     "traffic": None, "peak_source": peak_src, "note": "per device; see the N=1 line for the counters"},
        "field_replication": {"bytes_per_replica": int(np.prod(src.dims)) * 2, "first_call_ms": replicate_first_ms,
                              "steady_ms": replicate_ms, "replicas_equal_source": True},
        "cross_device_check": {"devices": G, "states_compared_per_device": m, "agree": agree},
        "precision": args.precision,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Native libraries print there too (the GPU boxes set NCCL_DEBUG=VERSION,
    whose banner goes to stdout), so file descriptor 1 is pointed at stderr for the rest of the process and the JSON
    line is written to the original descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.single_process:
        if world > 1:
            raise SystemExit("--single-process is launched without torchrun")
        run_single_process(args)
        return

    import torch
    import torch.distributed as dist
    from moveit_b200 import binding as B
    from moveit_b200 import engine, scenarios, sharding
    from moveit_b200 import robot_description as rd

    if not torch.cuda.is_available() or B.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    B.check(B.load().b200df_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # ---- set-up (untimed): model, group, world field (built on rank 0, replicated by ONE broadcast)
    desc = rd.panda()
    md = engine.assemble_model(desc, engine.body_decomposition, scenarios.FIELD_A["resolution"])
    model = engine.Model(md)
    links = desc.group_links("")
    se, ie = engine.resolve_acm(md, links, desc.acm_entry_matrix())
    fa = scenarios.FIELD_A
    group = engine.Group(model, links, se, ie, fa["resolution"], 0.0, fa["max_distance"])
    world_field = engine.Field.centered(fa["size"], fa["resolution"], fa["center"], fa["max_distance"])
    replicated_bytes = 0
    if rank == 0:
        for t, dims, pose in scenarios.random_scene():
            world_field.add_posed_body(t, dims, pose)
        world_field.rebuild()
    if world > 1:
        replicated_bytes = sharding.replicate_field(world_field, dist, rank, 0, dev)

    n = args.states
    lo = rank * n  # each rank draws its own slice of one global SplitMix64 stream
    from moveit_b200.rng import uniform01
    lim_lo = np.array([l[0] for l in desc.limits])
    lim_hi = np.array([l[1] for l in desc.limits])
    q_host = torch.empty((n, desc.nvar), dtype=torch.float32).pin_memory()
    chunk = 1_000_000
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        u = uniform01(scenarios.SEED_STATES, m * desc.nvar, offset=(lo + s) * desc.nvar).reshape(m, desc.nvar)
        qq = lim_lo + (lim_hi - lim_lo) * u
        qq[:, 8] = qq[:, 7]  # finger2 mimics finger1
        q_host[s:s + m] = torch.from_numpy(qq.astype(np.float32))
    q_dev = q_host.to(dev)
    v_dev = torch.empty(n, dtype=torch.uint8, device=dev)
    v_host = torch.empty(n, dtype=torch.uint8).pin_memory()
    flags = B.CHECK_ROBOT | B.CHECK_INTRA
    precision = {"fast": B.PRECISION_FAST, "exact": B.PRECISION_EXACT, "v1": 2}[args.precision]
    stream = torch.cuda.current_stream()

    def step_device():
        group.check_device(q_dev.data_ptr(), B.Q_F32, n, v_dev.data_ptr(), world_field, None, flags, precision,
                           stream.cuda_stream)

    def step_host():
        B.check(B.load().b200df_check_batch(group.h, world_field.h, None, ctypes.c_void_p(q_host.data_ptr()), B.Q_F32,
                                            n, flags, precision,
                                            ctypes.cast(v_host.data_ptr(), ctypes.POINTER(ctypes.c_uint8))))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = B.load().b200df_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = B.load().b200df_launch_count() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    colliding = int(v_dev.sum().item())
    undecided = None
    if args.precision == "fast":
        # share of states the fp32 screening pass hands to the exact kernel (raw screening verdict 2), on a sample
        m = min(n, 2_000_000)
        raw = torch.empty(m, dtype=torch.uint8, device=dev)
        group.check_device(q_dev.data_ptr(), B.Q_F32, m, raw.data_ptr(), world_field, None, flags, 3, stream.cuda_stream)
        torch.cuda.synchronize()
        undecided = float((raw == 2).sum().item()) / m
        # and the proof that FAST == EXACT on this very batch
        v_exact = torch.empty(n, dtype=torch.uint8, device=dev)
        group.check_device(q_dev.data_ptr(), B.Q_F32, n, v_exact.data_ptr(), world_field, None, flags,
                           B.PRECISION_EXACT, stream.cuda_stream)
        torch.cuda.synchronize()
        if not torch.equal(v_exact, v_dev):
            raise SystemExit("FAST and EXACT verdicts differ on the bench batch")

    # ---- every rank holds the same field and computes the same verdicts (a rank with a corrupt replica would
    # otherwise go unnoticed: the throughput lines above never compare ranks)
    cross = {"ranks": world}
    nvox = world_field.dims[0] * world_field.dims[1] * world_field.dims[2]
    d2_t = torch.empty(world_field.device_arrays()[2], dtype=torch.uint8, device=dev)
    world_field.export_device(None, d2_t.data_ptr(), None, stream.cuda_stream)
    torch.cuda.synchronize()
    wts = (torch.arange(d2_t.numel(), device=dev, dtype=torch.int64) % 65521) + 1
    field_sum = torch.stack([d2_t.to(torch.int64).sum(), (d2_t.to(torch.int64) * wts).sum()])
    m = min(n, 100_000)
    u = uniform01(scenarios.SEED_STATES, m * desc.nvar, offset=0).reshape(m, desc.nvar)
    q0 = lim_lo + (lim_hi - lim_lo) * u
    q0[:, 8] = q0[:, 7]
    q0_dev = torch.from_numpy(q0.astype(np.float32)).to(dev)
    v0 = torch.empty(m, dtype=torch.uint8, device=dev)
    group.check_device(q0_dev.data_ptr(), B.Q_F32, m, v0.data_ptr(), world_field, None, flags, precision,
                       stream.cuda_stream)
    torch.cuda.synchronize()
    v_sum = torch.stack([v0.to(torch.int64).sum(), (v0.to(torch.int64) * wts[:m]).sum()])
    sums = torch.cat([field_sum, v_sum])
    lo_s, hi_s = sums.clone(), sums.clone()
    if world > 1:
        dist.all_reduce(lo_s, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_s, op=dist.ReduceOp.MAX)
    if not torch.equal(lo_s, hi_s):
        raise SystemExit("cross-rank check failed: ranks disagree on the replicated field or on the verdicts of the "
                         "first %d global states (min %s max %s)" % (m, lo_s.tolist(), hi_s.tolist()))
    if rank == 0 and not torch.equal(v0, v_dev[:m]):
        raise SystemExit("rank 0: verdicts of the first states differ between two calls")
    cross.update(field_checksum=[int(x) for x in field_sum.tolist()], verdict_checksum=[int(x) for x in v_sum.tolist()],
                 states_compared=m, agree=True)

    # ---- end-to-end through the host-buffer C-ABI call (page-locked fp32 rows, the fastest way in): args.steps steps
    def timed_host(fn, steps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        t_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    e2e_steps = args.steps
    e2e_s = timed_host(step_host, e2e_steps)
    same = bool(torch.equal(v_host.to(dev), v_dev))

    # the shape a CollisionEnv caller really has: RobotState doubles in ordinary (pageable) memory, pageable verdicts
    q64 = q_host.numpy().astype(np.float64)
    v_page = np.empty(n, dtype=np.uint8)

    def step_plugin():
        B.check(B.load().b200df_check_batch(group.h, world_field.h, None, q64.ctypes.data_as(ctypes.c_void_p), B.Q_F64,
                                            n, flags, precision, v_page.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))

    plugin_steps = max(2, min(args.steps, 10))
    plugin_s = timed_host(step_plugin, plugin_steps)
    same_plugin = bool(np.array_equal(v_page, v_host.numpy()))
    if not same_plugin:
        # fp32 rows rounded from these very doubles were checked above: the verdicts must agree state by state
        raise SystemExit("pageable fp64 call and pinned fp32 call disagree")

    # what the host's PCIe fabric gives N concurrent plain copies of the same bytes (no kernels): the ceiling of e2e
    def step_copy():
        q_dev.copy_(q_host, non_blocking=True)
        v_host.copy_(v_dev, non_blocking=True)

    copy_s = timed_host(step_copy, max(3, min(args.steps, 10)))
    copy_steps = max(3, min(args.steps, 10))
    clocks = sampler.summary()  # sampled over the timed regions (device-resident and end-to-end)

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = n * world / (ms_per_step * 1e-3)
        peak, peak_src = measured_peak_hbm()
        achieved = ALGO_BYTES_PER_STATE * n / (ms_per_step * 1e-3) / 1e9
        counters = ncu_counters(args.traffic_json, args.precision)
        per_gpu_rate = n * args.steps / (ms_total * 1e-3)
        line = {
            "metric": "state_validity_checks_per_sec", "value": value, "unit": "states/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64" if args.precision == "fast" else "f64",
            "data": "synthetic", "config": workload_config(n),
            "e2e": {"value": n * world * e2e_steps / e2e_s, "unit": "states/s",
                    "h2d_bytes_per_step": n * desc.nvar * 4, "d2h_bytes_per_step": n, "steps": e2e_steps,
                    "input": "page-locked fp32 rows (b200df_host_alloc-style), page-locked verdicts",
                    "verdicts_match_device_path": same,
                    "copy_ceiling_states_per_s": n * world * copy_steps / copy_s,
                    "copy_ceiling_note": "the same H2D + D2H bytes as plain cudaMemcpyAsync on every rank at once, "
                                         "no kernels: what the box's PCIe fabric allows e2e to reach"},
            "e2e_plugin": {"value": n * world * plugin_steps / plugin_s, "unit": "states/s", "steps": plugin_steps,
                           "input": "pageable fp64 rows (RobotState doubles in heap memory), pageable verdicts, "
                                    "through b200df_check_batch: rows are narrowed to fp32 while being staged into "
                                    "page-locked chunks by a few host threads, the undecided states go again as fp64",
                           "h2d_bytes_per_step": n * desc.nvar * 4, "host_bytes_read_per_step": n * desc.nvar * 8,
                           "d2h_bytes_per_step": n, "verdicts_match_pinned_fp32_call": same_plugin},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": counters["dram_bytes_per_state"] * n if counters else None,
                         "traffic_source": counters["source"] if counters else "no ncu capture file (--traffic-json)",
                         "peak_source": peak_src,
                         "issue": issue_roofline(counters, per_gpu_rate, clocks),
                         "l2": l2_roofline(counters, per_gpu_rate, clocks),
                         "kernel": ("validity_fast_kernel<float, unsigned char, 128, simple joints, no self field> twice: "
                                    "the field pass over every state (no pair code), then the pair pass over the "
                                    "states it left at 0 (+ validity_small_kernel on the undecided "
                                    "states)" if args.precision == "fast" else "validity_kernel<float, unsigned char, 64>"),
                         "passes": counters.get("passes") if counters else None,
                         "note": "37 algorithmic bytes/state (fp32 joint values in, verdict byte out) over the whole "
                                 "call; traffic / issue / l2 = ncu counters of the committed capture of ONE call "
                                 "(both passes summed) per state x this run's states. The FK+lookup and FK+pair "
                                 "kernels execute some hundred warp instructions per state and are issue/latency "
                                 "bound by construction (DESIGN.md section 4): `issue` and `l2` are the resources "
                                 "they actually load; the HBM-bound kernel of this path is the EDT, see "
                                 "edt_ms.field_b_roofline"},
            "colliding_fraction": colliding / n,
            "field_replication_bytes": replicated_bytes,
            "cross_rank_check": cross,
            "precision": args.precision,
            "fast_undecided_fraction": undecided,
            "fast_equals_exact_on_batch": True if args.precision == "fast" else None,
        }
        if not args.no_edt:
            try:
                line["edt_ms"] = edt_timings(engine, scenarios, B)
                if world == 1 and not args.no_cpu_baseline:
                    line["edt_ms"]["cpu_baseline"] = edt_cpu_baseline(scenarios)
            except Exception as ex:  # keep the headline line even if the big field does not fit
                line["edt_ms"] = {"error": str(ex)}
        if world == 1 and not args.no_extras:
            try:  # secondary workloads; their CPU figures belong to the cpu_baseline leg (oracle = the checker's port)
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import extras_bench
                line["extras"] = extras_bench.run(with_oracle=not args.no_cpu_baseline)
            except Exception as ex:
                line["extras"] = {"error": str(ex)}
        if world == 1 and not args.no_cpu_baseline:
            oracle_py, odesc, setup = build_oracle_env()
            rate, cores, ns, ts = cpu_rate(setup, odesc, oracle_py)
            line["cpu_baseline"] = {"value": rate, "unit": "states/s", "cores": cores, "kind": "port",
                                    "sample": "%d state checks (the first 4 M states of the bench batch, repeated; "
                                              "%.1f s), CPU oracle, one thread per host core; omits MoveIt's per-call "
                                              "GSR copies / string-map ACM look-ups, i.e. faster than stock MoveIt"
                                              % (ns, ts)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
