#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_full.txt
    python tools/ncu_summary.py json "<workload>" a.ncu-rep [b.ncu-rep ...] > profiles/rNN_ncu_kernels.json

The json form is what bench.py reads `roofline.traffic` from (dram__bytes_read.sum +
dram__bytes_write.sum per launch of the dominant kernel): regenerate it after every capture.
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("us", "usecond"):
            v *= 1e3
        elif r[ui] in ("ms", "msecond"):
            v *= 1e6
        tot[r[ki]] += v
        cnt[r[ki]] += 1
    T = sum(tot.values())
    print(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})")
    print("# cold-cache, serialised launches: compare SHARES, not absolutes")
    print(f"{'total_us':>12s} {'share':>7s} {'n':>4s} {'avg_us':>10s}  kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v / 1e3:12.1f} {100 * v / T:6.2f}% {cnt[k]:4d} {v / 1e3 / cnt[k]:10.1f}  {k[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# selected metrics from `ncu --set full --clock-control none --import-source on` ({path})")
    for r in rows[2:]:
        print("\n== " + r[idx["Kernel Name"]][:120])
        for k in KEYS:
            if k in idx:
                print(f"  {k:85s} {r[idx[k]]:>18s} {units[idx[k]]}")


def _num(v, unit):
    x = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "usecond": 1e-3, "ms": 1.0,
             "msecond": 1.0, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}
    return x * scale.get(unit, 1.0)


def as_json(workload, paths):
    import json
    import re
    m1, m2 = re.search(r"points=(\d+)", workload), re.search(r"nodes=(\d+)", workload)
    out = {"workload": workload, "points": int(m1.group(1)) if m1 else None,
           "nodes": int(m2.group(1)) if m2 else None, "source": paths, "how": "ncu --set full --clock-control none (one launch each; "
           "averaged over the launches captured)", "kernels": {}}
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        acc = defaultdict(list)
        for r in rows[2:]:
            acc[r[idx["Kernel Name"]]].append(r)
        for name, rs in acc.items():
            short = re.sub(r"\(.*", "", name).replace("void ", "")
            get = lambda k: sum(_num(r[idx[k]], units[idx[k]]) for r in rs) / len(rs) if k in idx else None
            out["kernels"][short] = {
                "launches_captured": len(rs),
                "time_ms": get("gpu__time_duration.sum"),
                "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
                "registers": get("launch__registers_per_thread"), "grid": get("launch__grid_size"),
                "pipe_fma_cycles_active_pct": get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                "inst_executed_pipe_fma_pct": get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "inst_executed_pipe_xu_pct": get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "sm_clock_ghz": get("sm__cycles_elapsed.avg.per_second"),
                "inst_executed": get("smsp__inst_executed.sum"),
            }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "json":
        as_json(sys.argv[2], sys.argv[3:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
