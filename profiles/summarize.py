#!/usr/bin/env python
"""Summarise ncu output kept under profiles/ (no GPU needed).

  python profiles/summarize.py launches <launches.csv>          -> markdown table of kernel shares
  python profiles/summarize.py full <report.ncu-rep> [kernel]   -> key metrics of one `ncu --set full` capture
"""
import csv
import collections
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        agg.setdefault(name, []).append(float(r[iv].replace(",", "")))
    ours = {k: v for k, v in agg.items() if "at::" not in k and "int_peak" not in k}
    tot = sum(sum(v) for v in ours.values())
    print("| kernel | launches | mean us | share of hot-path kernels |\n|---|---|---|---|")
    for k, v in agg.items():
        share = "%.1f%%" % (100 * sum(v) / tot) if k in ours else "-"
        print("| `%s` | %d | %.1f | %s |" % (k[:70], len(v), sum(v) / len(v) / 1e3, share))


def full(path, kernel=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        if kernel and kernel not in name:
            continue
        print("### %s\n\n| metric | unit | value |\n|---|---|---|" % name)
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print("| %s | %s | %s |" % (h, u, v))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
