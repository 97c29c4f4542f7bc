#!/bin/bash
# e2e (host C ABI) sweep over the host pipeline's knobs.  usage: tools/gpu_e2e_sweep.sh "ENV=val ENV=val" "ENV=val" ...
python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -1
for cfg in "$@"; do
  env $cfg python bench.py --skip-map --skip-cpu --skip-single --skip-configs --skip-variants 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$cfg', 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'copies-only', round(d['e2e']['link']['copies_only_frames_per_s']), d['config']['checksum']['accepted_index_sum'])"
done
