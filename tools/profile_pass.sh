#!/bin/bash
# One GPU-box pass that produces everything under profiles/ for a tag:  tools/profile_pass.sh r4a
# (bench lines of the five named shapes, the reference arm, the ncu launch list of bench.py, one ncu full-set capture of a
# 128-frame chunk, achieved tolerances, the blur front end, single-frame latency).  Every ncu command runs only after
# the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
tag=${1:-rX}
o=gpurun_out
set -x
python bench.py --steps 10 --warmup 3 > $o/${tag}_bench_C3.json 2> $o/${tag}_bench_C3.err || exit 1
for c in C1 C2 C4 C5; do
  python bench.py --config $c --steps 5 --warmup 3 > $o/${tag}_bench_$c.json 2> $o/${tag}_bench_$c.err || echo "bench $c failed"
done
python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_reference_arm.json 2> $o/${tag}_reference_arm.err
python bench.py --steps 2 --warmup 1 --no-bgr > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches_bench_py.csv \
    python bench.py --steps 2 --warmup 1 --no-bgr > $o/${tag}_ncu_launches.log 2>&1
python tools/prof_run.py 128 1 128 1 > $o/${tag}_prof_run.log 2>&1 && \
ncu --set full --clock-control none --import-source on --launch-skip 48 --launch-count 16 -f -o $o/${tag}_chunk128 \
    python tools/prof_run.py 128 1 128 1 > $o/${tag}_ncu_full.log 2>&1
python tools/gpu_tolerances.py > $o/${tag}_tolerances.json 2> $o/${tag}_tolerances.err
python tools/blur_bench.py > $o/${tag}_blur_front_end.txt 2>&1
python tools/latency_c1.py > $o/${tag}_latency_c1.txt 2>&1
AGPU_GRAPH=0 python tools/latency_c1.py > $o/${tag}_latency_c1_nograph.txt 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $o/${tag}_gpu.txt
