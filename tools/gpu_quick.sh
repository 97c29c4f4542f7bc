#!/bin/bash
# Quick GPU pass: parity tests + short bench (stage times).  usage: tools/gpu_quick.sh TAG [ENV=VAL ...]
TAG=${1:-q}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
env "$@" python bench.py --skip-map --skip-cpu --skip-single > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(round(d['value']),round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
