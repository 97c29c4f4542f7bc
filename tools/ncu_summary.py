#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
  python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep [regex]    > profiles/rNN_kernel.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
           "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
           "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
           "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
           "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else (v * 1000.0 if r[ui] == "ms" else v)   # -> us
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.1f | %.1f%% | %.1f |" % (k[:70], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))


def kernel(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[ki]):
            continue
        print("### %s  (id %s)\n" % (r[ki].split("(")[0], r[0]))
        print("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("| %s | %s | %s |" % (m, r[i], units[i]))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
