#!/usr/bin/env python
"""One line of key counters + top stall reasons per kernel launch of an ncu report (captured with --set full).

  python tools/ncu_brief.py gpurun_out/prof.ncu-rep [kernel-regex]
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum']


def main(path, regex=None):
    cmd = ["ncu", "-i", path, "--page", "raw", "--csv"] + (["-k", "regex:" + regex] if regex else [])
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
    h = rows[0]
    keys = [k for k in KEYS if k in h]
    st = [i for i, a in enumerate(h) if a.startswith('smsp__average_warps_issue_stalled') and a.endswith('_per_issue_active.ratio')]
    for r in rows[2:]:
        print(r[h.index('Kernel Name')][:36], ' '.join('%s=%s' % (k.split('.')[0].split('__')[1][:16], r[h.index(k)][:8]) for k in keys))
        s = sorted([(float(r[i]), h[i].split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')) for i in st], reverse=True)[:6]
        print('    stalls per issue:', ', '.join('%s %.2f' % (n, x) for x, n in s))


if __name__ == "__main__":
    main(*sys.argv[1:])
