"""Summarise an ncu report for profiles/: per kernel launch the metrics DESIGN.md / bench.py quote.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex]      (needs `ncu` on PATH; reads the report only)
    python scripts/ncu_summary.py --launches gpurun_out/list.csv [skip]   (per-kernel shares of an ncu launch list)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs/thread", "launch__registers_per_thread"),
    ("dynamic smem / block", "launch__shared_mem_per_block_dynamic"),
    ("occupancy limit (smem) blocks/SM", "launch__occupancy_limit_shared_mem"),
    ("occupancy limit (regs) blocks/SM", "launch__occupancy_limit_registers"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("IPC (per SM, max 4)", "sm__inst_executed.avg.per_cycle_active"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("active lanes per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("shared-memory bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("alu pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("lsu pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "mio_throttle", "branch_resolving", "lg_throttle",
          "not_selected", "no_instruction", "math_pipe_throttle", "dispatch_stall"]


def report(path, pattern):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        name = d["Kernel Name"]
        if pattern and not re.search(pattern, name):
            continue
        u = dict(zip(head, units))
        print("kernel:", name)
        for label, key in KEYS:
            if key in d:
                print(f"  {label:42s} {d[key]} {u[key]}")
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in d:
                print(f"  stall {s + ' /issue':36s} {d[key]}")
        print()


def launches(path, skip):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1 + skip:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        ms = v / 1e6 if r[ui].startswith("n") else v / 1e3 if r[ui].startswith("u") else v
        name = re.sub(r"^void ", "", r[ki]).split("(")[0].split("<")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1]:9.3f} ms  {a[1] / tot * 100:5.1f} %  {a[0]:4d} launches  {k}")
    print(f"{tot:9.3f} ms total")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        report(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
