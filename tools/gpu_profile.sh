#!/bin/bash
# One GPU-box pass: whole GPU suite, launch list, full ncu capture of one 512-frame chunk (default kernels + the opt-in IMMA
# Hamming variant), default bench.  Outputs -> gpurun_out/   usage: tools/gpu_profile.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
SMALL="python bench.py --frames 512 --steps 1 --warmup 1 --skip-map --skip-cpu --skip-single --skip-configs --skip-variants --skip-other"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > /dev/null 2> gpurun_out/ncu_launch_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:'resize|fast|octree|blur|orient|knn2' -s 12 -c 12 -f -o gpurun_out/prof_$TAG $SMALL > /dev/null 2> gpurun_out/ncu_full_$TAG.err
tail -2 gpurun_out/ncu_full_$TAG.err
ORBX_HAMM_MMA=1 ncu --set full --clock-control none --import-source on -k regex:knn2_pairs_mma -c 1 -f -o gpurun_out/prof_${TAG}_mma $SMALL > /dev/null 2> gpurun_out/ncu_mma_$TAG.err
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print(d['value'],d['e2e']['value'],{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null; echo "reference arm rc=$?"
