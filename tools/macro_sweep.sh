#!/bin/bash
cd $GRAFT_REPO_ROOT
run() { echo "== $1 $2"; env $2 AGPU_LIB=$PWD/aprilslam_b200/$1 python tools/prof_run.py 128 1 128 1 2>&1 | tail -2; env $2 AGPU_LIB=$PWD/aprilslam_b200/$1 python tools/prof_run.py 1024 1 0 3 2>&1 | tail -2 | head -1; }
run libaprilgpu.so ""
run libv_EDGE_MINB_20.so ""
run libv_EDGE_MINB_28.so ""
run libv_QF_MINB2_12.so "AGPU_TIER_CTAS=3,12,8,3,1"
run libv_QF_MINB2_20.so "AGPU_TIER_CTAS=3,20,8,3,1"
run libv_DEC_MINB_3.so "AGPU_DECODE_CTAS=3"
