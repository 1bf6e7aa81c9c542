#!/bin/bash
# Launch-bound sweep: builds libaprilgpu variants with other __launch_bounds__ macros (EDGE_MINB, QF_MINB2, DEC_MINB) next
# to the product library and runs tools/prof_run.py on each (AGPU_LIB selects the library).  Build here (no GPU needed):
#   tools/macro_sweep.sh build
# then on the GPU box (gpurun ships the in-tree .so files):   tools/macro_sweep.sh run > gpurun_out/macro.txt
# and remove the variants afterwards:                          tools/macro_sweep.sh clean
cd "$(dirname "$0")/.."
VARIANTS="EDGE_MINB=20 EDGE_MINB=28 QF_MINB2=12 QF_MINB2=20 DEC_MINB=3"
name() { echo "aprilslam_b200/libv_$(echo $1 | tr '=' '_').so"; }
case "$1" in
build)
  for v in $VARIANTS; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --shared \
        -Xcompiler -fPIC -D$v -o $(name $v) aprilslam_b200/csrc/aprilgpu.cu &
  done; wait ;;
clean) rm -f aprilslam_b200/libv_*.so ;;
run)
  run() { echo "== $1 $2"; env $2 AGPU_LIB=$PWD/$1 python tools/prof_run.py 128 1 128 1 2>&1 | tail -2; }
  run aprilslam_b200/libaprilgpu.so ""
  run $(name EDGE_MINB=20) ""
  run $(name EDGE_MINB=28) ""
  run $(name QF_MINB2=12) "AGPU_TIER_CTAS=3,12,8,3,1"
  run $(name QF_MINB2=20) "AGPU_TIER_CTAS=3,20,8,3,1"
  run $(name DEC_MINB=3) "AGPU_DECODE_CTAS=3" ;;
*) echo "usage: $0 build|run|clean" ;;
esac
