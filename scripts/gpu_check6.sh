#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
for cfg in "2 1 23 1 none" "2 1 23 1 polyfit"; do
set -- $cfg
echo "== microbench bps=$1 tma=$2 shift=$3 hint=$4 value=$5"; timeout 300 python scripts/engine_microbench.py 20 $1 $2 $3 $4 $5 > gpurun_out/mb_$5.json 2> gpurun_out/mb_$5.err; echo "rc=$?"; cat gpurun_out/mb_$5.json; tail -3 gpurun_out/mb_$5.err
done
echo "== bench both"; timeout 900 python bench.py --steps 10 --warmup 3 --config both --no-e2e --breakdown > gpurun_out/bench_both.json 2> gpurun_out/bench_both.err; echo "rc=$?"; cat gpurun_out/bench_both.json; tail -5 gpurun_out/bench_both.err
