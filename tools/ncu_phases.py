#!/usr/bin/env python
"""Stall samples / executed warp instructions of one kernel aggregated over named source-line ranges (phases).

  python tools/ncu_phases.py REPORT.ncu-rep KERNEL_REGEX name:lo-hi[,lo-hi...] [name:...]
"""
import csv
import io
import subprocess
import sys


def load(path, regex):
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
                data.append((fname, int(r[0]), int(r[si]), int(r[ie])))
            except ValueError:
                pass
    return data


def main():
    data = load(sys.argv[1], sys.argv[2])
    phases = []
    for spec in sys.argv[3:]:
        name, rng = spec.split(":")
        phases.append((name, [tuple(int(v) for v in x.split("-")) for x in rng.split(",")]))
    ts = sum(d[2] for d in data) or 1
    ti = sum(d[3] for d in data) or 1
    acc = {name: [0, 0] for name, _ in phases}
    acc["(other files)"] = [0, 0]; acc["(unassigned)"] = [0, 0]
    for fname, line, s, i in data:
        if not fname.endswith(".cu"):
            acc["(other files)"][0] += s; acc["(other files)"][1] += i
            continue
        for name, rs in phases:
            if any(lo <= line <= hi for lo, hi in rs):
                acc[name][0] += s; acc[name][1] += i
                break
        else:
            acc["(unassigned)"][0] += s; acc["(unassigned)"][1] += i
    print("total samples %d, warp instructions %d" % (ts, ti))
    print("| phase | samples | instructions |\n|---|---|---|")
    for k, (s, i) in acc.items():
        print("| %s | %.1f%% | %.1f%% |" % (k, 100 * s / ts, 100 * i / ti))


if __name__ == "__main__":
    main()
