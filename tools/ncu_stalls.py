#!/usr/bin/env python
"""Stall-reason shares + a few headline counters of one kernel in an ncu report.  python tools/ncu_stalls.py REP KERNEL_REGEX"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "-k", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]; vals = rows[2] if len(rows) > 2 else rows[1]
d = dict(zip(hdr, vals))
def f(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return None
keys = [k for k in d if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k]
tot = sum(f(k) or 0 for k in keys) or 1
print("kernel", d.get("Kernel Name"), "duration us", (f("gpu__time_duration.sum") or 0) / 1e3)
for k in sorted(keys, key=lambda k: -(f(k) or 0))[:9]:
    print("  %-62s %5.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * (f(k) or 0) / tot))
for k in ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
          "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__warps_eligible.avg.per_cycle_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_tensor_op_imma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
          "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]:
    if k in d: print("  %-82s %s" % (k, d[k]))
for k in d:
    if "tensor" in k and k not in ("",) and d[k] not in ("", "0") and "pct" in k: print("  [tensor] %-70s %s" % (k, d[k]))
