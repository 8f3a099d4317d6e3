#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for cfg in "2 1 23 1" "2 1 23 0"; do
set -- $cfg
echo "== microbench bps=$1 tma=$2 shift=$3 hint=$4"; timeout 300 python scripts/engine_microbench.py 20 $1 $2 $3 $4 > gpurun_out/mb_$1_$2_$3_$4.json 2> gpurun_out/mb_$1_$2_$3_$4.err; echo "rc=$?"; cat gpurun_out/mb_$1_$2_$3_$4.json; tail -3 gpurun_out/mb_$1_$2_$3_$4.err
done
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 --breakdown > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
