#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for v in none polyfit; do
echo "== microbench value=$v"; timeout 300 python scripts/engine_microbench.py 20 2 1 23 1 $v > gpurun_out/mb7_$v.json 2> gpurun_out/mb7_$v.err; echo "rc=$?"; cat gpurun_out/mb7_$v.json; tail -3 gpurun_out/mb7_$v.err
done
bash scripts/gpu_sanitize.sh
