#!/bin/bash
# A/B of library builds on the resident step: default lib, then every variant under vo_slam_test_b200/lib/variants/ (parity first).
mkdir -p gpurun_out
run() { env "$@" python bench.py --skip-map --skip-cpu --skip-single --skip-configs 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"; }
echo "== default"; run X=1
for v in vo_slam_test_b200/lib/variants/*/libvoslam_b200.so; do
  [ -f "$v" ] || continue
  echo "== $v"
  ORBX_LIB=$PWD/$v python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -1
  run ORBX_LIB=$PWD/$v
done
