#!/usr/bin/env python
"""Turns an ncu report / launch list brought back in gpurun_out/ into the small text summaries kept in profiles/.

  python profiles/summarize.py rep  gpurun_out/prof.ncu-rep   profiles/<name>.txt
  python profiles/summarize.py list gpurun_out/launches.csv   profiles/<name>.txt
  python profiles/summarize.py traffic gpurun_out/prof.ncu-rep profiles/<name>.json [states_per_launch]
      per-launch DRAM bytes / L2 sectors / warp instructions of the validity kernels in the report, as the JSON
      bench.py --traffic-json reads (keys "fast" and "exact"); merges into an existing file
"""
import json
import os
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
]


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n" % path)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write("\nkernel: %s\n" % name[:200])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("  %-75s %s %s\n" % (k, r[i], units[i]))


def launch_list(path, out):
    agg = OrderedDict()
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0][-90:]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    total = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write("# launch list (ncu --metrics gpu__time_duration.sum --clock-control none) of %s\n" % path)
        f.write("# cold-cache, serialised: compare SHARES, not absolutes\n")
        f.write("%-92s %8s %14s %8s\n" % ("kernel", "launches", "total_us", "share"))
        for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-92s %8d %14.1f %7.2f%%\n" % (name, cnt, us, 100 * us / total))


def _num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def traffic(path, out, states=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in hdr}

    def val(r, k, to_bytes=False):
        if k not in col:
            return 0.0
        v = _num(r[col[k]])
        if to_bytes:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col[k]], 1.0)
        return v

    data = {}
    if os.path.exists(out):
        with open(out) as f:
            data = json.load(f)
    data["report"] = os.path.basename(path)
    seen = set()
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = "fast" if "validity_fast_kernel" in name else "exact" if name.startswith("validity_kernel") or \
            "::validity_kernel" in name else None
        if key is None:
            continue
        n = float(states) if states else val(r, "launch__grid_size") * val(r, "launch__block_size")
        dur = val(r, "gpu__time_duration.sum")
        dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[col["gpu__time_duration.sum"]], 1.0)
        one = {"kernel": name[:160], "states_per_launch": n,
               "dram_bytes_read": val(r, "dram__bytes_read.sum", True),
               "dram_bytes_write": val(r, "dram__bytes_write.sum", True),
               "lts_sectors_read": val(r, "lts__t_sectors_srcunit_tex_op_read.sum"),
               "lts_sectors_write": val(r, "lts__t_sectors.sum") - val(r, "lts__t_sectors_srcunit_tex_op_read.sum"),
               "lts_sector_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
               "warp_instructions": val(r, "smsp__inst_executed.sum"),
               "duration_us_under_ncu": dur_us,
               "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
               "registers": val(r, "launch__registers_per_thread"),
               "grid": val(r, "launch__grid_size")}
        if key in seen:
            # the call runs several kernels of this kind (FAST: the field pass, then the pair pass over the survivors):
            # the per-call counters are their sums, the per-pass rows are kept beside them
            tot = data[key]
            for f_ in ("dram_bytes_read", "dram_bytes_write", "lts_sectors_read", "lts_sectors_write",
                       "warp_instructions", "duration_us_under_ncu"):
                tot[f_] += one[f_]
            tot["kernel"] = tot["passes"][0]["kernel"][:70] + " ... + " + one["kernel"][:70]
            tot["issue_active_pct"] = sum(p_["issue_active_pct"] * p_["duration_us_under_ncu"]
                                          for p_ in tot["passes"] + [one]) / tot["duration_us_under_ncu"]
            tot["warps_active_pct"] = sum(p_["warps_active_pct"] * p_["duration_us_under_ncu"]
                                          for p_ in tot["passes"] + [one]) / tot["duration_us_under_ncu"]
            tot["passes"].append(one)
        else:
            seen.add(key)
            data[key] = dict(one)
            data[key]["passes"] = [one]
    with open(out, "w") as f:
        json.dump(data, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    fn = {"rep": rep, "list": launch_list, "traffic": traffic}[sys.argv[1]]
    fn(*sys.argv[2:])
