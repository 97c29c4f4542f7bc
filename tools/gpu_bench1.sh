#!/bin/bash
# One-GPU bench line (default flags) -> gpurun_out/bench_$1.json, plus a short summary on stdout.
TAG=${1:-x}; shift
mkdir -p gpurun_out
env "$@" python bench.py $BENCH_ARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read())
print("value %.0f  e2e %.0f  ms/step %.2f  scaling %s  launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"], d["gpu_launches"]))
print({k: round(v["ms_per_step"], 2) for k, v in d["stages"].items()})
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "link", d["e2e"]["link"])
print("hamming_map", d["hamming_map"])
print("cpu", d["cpu_baseline"])
print("parity", d["config"]["parity_spot_check"], d["config"]["checksum"])
for k in ("config3_1080p", "config3_4k", "config4_sbp", "tracking_frame", "configs_error", "single_frame", "other_scaling", "variants"):
    if k in d: print(k, d[k])
PY
