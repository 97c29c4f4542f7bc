#!/bin/bash
# N-GPU bench line under torchrun -> gpurun_out/bench_$TAG_n$N.json + summary.  usage: tools/gpu_benchN.sh TAG N [extra bench args]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "rc=$?"
tail -2 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_n$N.json").read())
print("N=%d value %.0f ms/step %.2f scaling %s | e2e %.0f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["scaling"], d["e2e"]["value"]))
print("link", d["e2e"]["link"])
print("other", d.get("other_scaling"))
print("map", d["hamming_map"])
print("checksum", d["config"]["checksum"]["keypoints"], d["config"]["checksum"]["accepted_matches"], d["config"]["checksum"]["descriptor_byte_sum"], d["config"]["parity_spot_check"])
print({k: round(v["ms_per_step"], 2) for k, v in d["stages"].items()})
PY
