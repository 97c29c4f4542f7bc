#!/bin/bash
# parity of the extractor + one resident A/B line.  usage: tools/gpu_quick2.sh [ENV=val ...]
python -m pytest tests/test_gpu_extract.py tests/test_gpu_batch.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -2
env "$@" python bench.py --skip-map --skip-cpu --skip-single --skip-configs --skip-variants 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['config']['checksum']['accepted_index_sum'])"
