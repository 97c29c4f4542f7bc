#!/usr/bin/env python
"""Per-CUDA-source-line share of stall samples / executed instructions of one kernel in an ncu report
(captured with --set full --import-source on; code built with -lineinfo).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep KERNEL_REGEX [TOP]
"""
import csv
import io
import subprocess
import sys


def main(path, regex, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + regex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    data, fname, hdr = [], "", None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = r
            si, ie = r.index("# Samples"), r.index("Instructions Executed")
        elif hdr and len(r) > 8 and r[0] != "":
            try:
                data.append((int(r[si]), int(r[ie]), fname, int(r[0]), r[1].strip()))
            except ValueError:
                pass
    ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
    print("total samples %d, warp instructions %d" % (ts, ti))
    print("| samples | instr | line | source |\n|---|---|---|---|")
    for d in sorted(data, reverse=True)[:top]:
        print("| %.1f%% | %.1f%% | %s:%d | `%s` |" % (100 * d[0] / ts, 100 * d[1] / ti, d[2], d[3], d[4][:110]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
