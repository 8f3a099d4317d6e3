#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_*.txt
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "single_rank or resnet50_shapes or state_roundtrip or fused_recipe or topk_select_exact" --timeout 300 -x 2>&1 | tail -15 > gpurun_out/v13_tests.log
tail -8 gpurun_out/v13_tests.log
timeout 120 python scripts/engine_microbench.py 20 2 1 22 > gpurun_out/microbench_v13_hs22.json 2> gpurun_out/microbench_v13_hs22.err; echo "mb22 rc=$?"; cat gpurun_out/microbench_v13_hs22.json; tail -3 gpurun_out/microbench_v13_hs22.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v13_n1.json 2> gpurun_out/bench_v13_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_v13_n1.json | cut -c1-2500; tail -5 gpurun_out/bench_v13_n1.err
for f in gpurun_out/diag_*.txt; do [ -f "$f" ] && { echo "== $f"; head -12 "$f"; }; done 2>/dev/null | head -60
