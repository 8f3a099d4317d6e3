#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_gpu_rle.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_rle.log
for cfg in rle topk dense; do
echo "== bench ncf $cfg"
timeout 600 python bench.py --model ncf --config $cfg --steps 20 --warmup 5 --breakdown > gpurun_out/bench_ncf_$cfg.json 2> gpurun_out/bench_ncf_$cfg.err; echo "rc=$?"; grep "^{" gpurun_out/bench_ncf_$cfg.json | cut -c1-260; grep -o '"exchange_ms_per_step": [0-9.]*\|"wire_bytes[a-z_]*": [0-9.]*' gpurun_out/bench_ncf_$cfg.json; tail -2 gpurun_out/bench_ncf_$cfg.err
done
