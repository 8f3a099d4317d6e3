#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (engine tests)"; timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -q --maxfail=30 -p no:cacheprovider -k "engine or trainer" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for cfg in "2 1 23" "2 0 23" "2 1 22" "1 1 23"; do
set -- $cfg
echo "== microbench bps=$1 tma=$2 shift=$3"; timeout 300 python scripts/engine_microbench.py 20 $1 $2 $3 > gpurun_out/mb_$1_$2_$3.json 2> gpurun_out/mb_$1_$2_$3.err; echo "rc=$?"; cat gpurun_out/mb_$1_$2_$3.json; tail -3 gpurun_out/mb_$1_$2_$3.err
done
