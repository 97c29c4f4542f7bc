#!/bin/bash
# Round 2, GPU call 1: whole GPU suite (no opt-in tests left), fast_a2 variant parity + A/B, per-source-line ncu profiles of the four
# extractor kernels that hold the step (FAST, quadtree, blur, orientation/descriptor).  Outputs -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_c1.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2_c1.log
for v in vo_slam_test_b200/lib/variants/*/libvoslam_b200.so; do
  [ -f "$v" ] || continue
  ORBX_LIB=$PWD/$v python -m pytest tests/test_gpu_extract.py tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/pytest_$(basename $(dirname $v)).log 2>&1
  echo "variant $v parity rc=$?"
done
bash tools/gpu_ab.sh 2>&1 | tee gpurun_out/ab_r2_c1.txt
SMALL="python bench.py --frames 512 --steps 1 --warmup 1 --skip-map --skip-cpu --skip-single"
for k in fast_warp octree_kernel blur_tma orient_desc_tma; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/prof_r2c1_$k $SMALL > /dev/null 2> gpurun_out/ncu_r2c1_$k.err
  python tools/ncu_lines.py gpurun_out/prof_r2c1_$k.ncu-rep $k 60 > gpurun_out/lines_r2c1_$k.txt 2>&1
done
head -20 gpurun_out/lines_r2c1_fast_warp.txt
ls -la gpurun_out/*.ncu-rep
