#!/bin/bash
run() { env "$@" python bench.py --skip-map --skip-cpu --skip-single --skip-configs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['config']['checksum']['accepted_index_sum'])"; }
python -m pytest tests/test_gpu_hamming_mma.py -m gpu -x -q 2>&1 | tail -2
echo "== default lib, POPC"; run ORBX_HAMM_MMA=0; echo "== default lib, MMA"; run ORBX_HAMM_MMA=1; ORBX_HAMM_MMA=1 python bench.py --frames 64 --steps 1 --warmup 1 --skip-cpu --skip-single --skip-configs 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.read())[\"hamming_map\"])"
for v in vo_slam_test_b200/lib/variants/*/libvoslam_b200.so; do [ -f "$v" ] || continue; echo "== $v MMA"; run ORBX_HAMM_MMA=1 ORBX_LIB=$PWD/$v; done
