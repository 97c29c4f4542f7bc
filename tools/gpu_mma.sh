#!/bin/bash
# parity + A/B + ncu stall summary of the IMMA Hamming variant
python -m pytest tests/test_gpu_hamming_mma.py -m gpu -x -q 2>&1 | tail -3
for m in 0 1; do ORBX_HAMM_MMA=$m python bench.py --skip-map --skip-cpu --skip-single --skip-configs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['config']['checksum']['accepted_matches'], d['config']['checksum']['accepted_index_sum'])"; done
ORBX_HAMM_MMA=1 ncu --set full --clock-control none --import-source on -k regex:knn2_pairs_mma -c 1 -f -o gpurun_out/prof_mma python bench.py --frames 512 --steps 1 --warmup 1 --skip-map --skip-cpu --skip-single --skip-configs --skip-other > /dev/null 2> gpurun_out/ncu_mma.err
python tools/ncu_stalls.py gpurun_out/prof_mma.ncu-rep knn2_pairs_mma | grep -v "\[tensor\]"
