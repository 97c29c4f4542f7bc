#!/bin/bash
# First GPU call of round 2 (about 3 GPU-minutes).  Before the call, on the CPU side:
#     tools/build_variants.sh fast_a2 "-DORBX_FAST_A2=1"
# 1. the whole GPU suite including the tests written without GPU minutes left (opt-in switch) and the adapter program against the
#    reference's own compiled matcher.cpp;  2. parity + A/B of every built library variant;  3. a per-source-line profile of the
#    FAST kernel (where do the issue slots go: stage a, stage b, arc score, NMS?).  Outputs -> gpurun_out/
mkdir -p gpurun_out
ORBX_EXTRA_GPU_TESTS=1 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_first.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2_first.log
for v in vo_slam_test_b200/lib/variants/*/libvoslam_b200.so; do
  [ -f "$v" ] || continue
  ORBX_LIB=$PWD/$v python -m pytest tests/test_gpu_extract.py tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/pytest_$(basename $(dirname $v)).log 2>&1
  echo "variant $v parity rc=$?"
done
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2_first.txt
SMALL="python bench.py --frames 512 --steps 1 --warmup 1 --skip-map --skip-cpu --skip-single"
ncu --set full --clock-control none --import-source on -k regex:fast_warp -c 1 -f -o gpurun_out/prof_fast_r2 $SMALL > /dev/null 2> gpurun_out/ncu_fast_r2.err
python tools/ncu_lines.py gpurun_out/prof_fast_r2.ncu-rep fast_warp 60 > gpurun_out/fast_lines_r2.txt 2>&1; head -30 gpurun_out/fast_lines_r2.txt
