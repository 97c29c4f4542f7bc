#!/bin/bash
# resident step over lane counts / chunk sizes.  usage: tools/gpu_lanes.sh "ENV=val ..." ...
python -m pytest tests/test_gpu_batch.py tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -1
for cfg in "$@"; do
  env $cfg python bench.py --skip-map --skip-cpu --skip-single --skip-configs --skip-variants 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$cfg', 'resident', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['config']['checksum']['accepted_index_sum'])"
done
