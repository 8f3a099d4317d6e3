#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_*.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -12 > gpurun_out/final_gpu_tests.log; tail -8 gpurun_out/final_gpu_tests.log
timeout 300 python bench.py --model bert_large --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_bert_r2_n1.json 2> gpurun_out/bench_bert_r2_n1.err; echo "bert rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_bert_r2_n1.json'));print({k:d.get(k) for k in ('value','ms_per_step','exchange_ms_per_step','relative_volume')}, d['harness']['buckets'], d.get('dense_allreduce_context'))"; tail -2 gpurun_out/bench_bert_r2_n1.err | cut -c1-300
timeout 300 python bench.py --model ncf --config rle --steps 12 --warmup 4 --no-e2e > gpurun_out/bench_ncf_rle_r2_n1.json 2> gpurun_out/bench_ncf_rle_r2_n1.err; echo "ncf rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_ncf_rle_r2_n1.json'));print({k:d.get(k) for k in ('value','ms_per_step','exchange_ms_per_step','relative_volume')}, d.get('dense_allreduce_context'))"; tail -2 gpurun_out/bench_ncf_rle_r2_n1.err | cut -c1-300
for f in gpurun_out/diag_*.txt; do [ -f "$f" ] && { echo "== $f"; head -12 "$f"; }; done 2>/dev/null | head -40
