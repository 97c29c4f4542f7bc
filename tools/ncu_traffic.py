#!/usr/bin/env python
"""Per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full` capture -> JSON.

  python tools/ncu_traffic.py gpurun_out/prof.ncu-rep FRAMES_PER_LAUNCH > profiles/rNN_traffic.json

bench.py reads the file to fill roofline.traffic (per launch, like roofline.achieved).
"""
import collections
import csv
import io
import json
import subprocess
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TSCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path, frames):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = (hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "gpu__time_duration.sum"))
    agg = collections.OrderedDict()
    for r in rows[2:]:
        b = float(r[ri].replace(",", "")) * SCALE[units[ri]] + float(r[wi].replace(",", "")) * SCALE[units[wi]]
        t = float(r[ti].replace(",", "")) * TSCALE[units[ti]]
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0, 0.0])
        a[0] += 1; a[1] += b; a[2] += t
    res = {"frames_per_launch": frames, "source": "%s (ncu --set full --clock-control none)" % path, "kernels": {}}
    for k, (n, b, t) in agg.items():
        res["kernels"][k] = {"launches": n, "dram_bytes_per_launch": b / n, "dram_bytes_per_frame": round(b / frames),
                             "us_per_launch_under_ncu": t / n}
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
