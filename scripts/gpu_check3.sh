#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for bps in 1 2; do
echo "== microbench bps=$bps"; timeout 300 python scripts/engine_microbench.py 20 $bps > gpurun_out/microbench_bps$bps.json 2> gpurun_out/microbench_bps$bps.err; echo "rc=$?"; cat gpurun_out/microbench_bps$bps.json; tail -3 gpurun_out/microbench_bps$bps.err
done
bash scripts/gpu_profile_phases.sh 2
