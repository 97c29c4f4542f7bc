#!/bin/bash
# Build A/B variants of the library into vo_slam_test_b200/lib/variants/NAME/ (picked up by tools/gpu_ab.sh via ORBX_LIB).
# usage: tools/build_variants.sh NAME "-DFLAG=..." [NAME "-D..."]...
set -e
cd "$(dirname "$0")/../vo_slam_test_b200/csrc"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  out=../lib/variants/$name; mkdir -p $out
  for f in *.cu; do
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags -c $f -o $out/${f%.cu}.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o $out/libvoslam_b200.so $out/*.o
  rm -f $out/*.o
  echo "built $out"
done
