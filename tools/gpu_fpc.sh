#!/bin/bash
# parity of the FAST kernel changes + sweep of frames per CTA
python -m pytest tests/test_gpu_extract.py tests/test_gpu_batch.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -3
run() { env "$@" python bench.py --skip-map --skip-cpu --skip-single --skip-configs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"; }
for k in 1 2 4 8 16; do echo "== ORBX_FAST_FPC=$k"; run ORBX_FAST_FPC=$k; done
