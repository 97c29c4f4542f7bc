#!/bin/bash
# A/B of library builds: default lib, then every vo_slam_test_b200/lib/variants/*/libvoslam_b200.so (ORBX_LIB).  Extra env via args.
mkdir -p gpurun_out
run() { env "$@" python bench.py --skip-map --skip-cpu --skip-single 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['e2e']['value']),{k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"; }
echo "== default"; run "$@" X=1
for v in vo_slam_test_b200/lib/variants/*/libvoslam_b200.so; do [ -f "$v" ] || continue; echo "== $v"; run "$@" ORBX_LIB=$PWD/$v; done
