#!/bin/bash
mkdir -p gpurun_out
for mb in 32 64 128; do
echo "== bench bloom bucket=$mb"; timeout 600 python bench.py --steps 15 --warmup 4 --no-e2e --breakdown --bucket-mb $mb > gpurun_out/bench_b$mb.json 2> gpurun_out/bench_b$mb.err; grep -o '"ms_per_step": [0-9.]*\|"exchange_ms_per_step": [0-9.]*\|"value": [0-9.]*' gpurun_out/bench_b$mb.json | tr '\n' ' '; echo; tail -2 gpurun_out/bench_b$mb.err
done
echo "== bench bloom bucket=128 no-overlap"; timeout 600 python bench.py --steps 15 --warmup 4 --no-e2e --bucket-mb 128 --no-overlap > gpurun_out/bench_b128_noov.json 2> gpurun_out/bench_b128_noov.err; grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*' gpurun_out/bench_b128_noov.json | tr '\n' ' '; echo
echo "== bench dense"; timeout 600 python bench.py --steps 15 --warmup 4 --no-e2e --config dense > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*' gpurun_out/bench_dense.json | tr '\n' ' '; echo
