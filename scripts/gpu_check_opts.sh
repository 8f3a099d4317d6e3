#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/pytest_gpu_opts.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_opts.log
for o in "0 0" "1 0" "0 1" "1 1"; do set -- $o
DR_OWN_FLAGS=$1 DR_EMIT_COUNTS=$2 timeout 60 python scripts/engine_microbench.py 20 2 > gpurun_out/microbench_v10_own$1_cnt$2.json 2>/dev/null; echo "own=$1 counts=$2 $(grep -o '"fused_ms_median": [0-9.]*' gpurun_out/microbench_v10_own$1_cnt$2.json) $(grep -o '"emit": [0-9.]*' gpurun_out/microbench_v10_own$1_cnt$2.json) $(grep -o '"query": [0-9.]*' gpurun_out/microbench_v10_own$1_cnt$2.json) $(grep -o '"decode": [0-9.]*' gpurun_out/microbench_v10_own$1_cnt$2.json)"
done
