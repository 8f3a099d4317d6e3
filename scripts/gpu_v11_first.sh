#!/bin/bash
# first hardware run of the v11 engine: oracle tests (all failures, not -x), then the microbenchmark
mkdir -p gpurun_out
rm -f gpurun_out/diag_*.txt
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "single_rank or resnet50_shapes or state_roundtrip" --timeout 300 2>&1 | tail -60 > gpurun_out/v11_tests.log
tail -40 gpurun_out/v11_tests.log
timeout 120 python scripts/engine_microbench.py 20 2 1 22 > gpurun_out/microbench_v11_hs22.json 2> gpurun_out/microbench_v11_hs22.err; echo "mb22 rc=$?"; cat gpurun_out/microbench_v11_hs22.json; tail -5 gpurun_out/microbench_v11_hs22.err
timeout 120 python scripts/engine_microbench.py 20 2 1 23 > gpurun_out/microbench_v11_hs23.json 2> gpurun_out/microbench_v11_hs23.err; echo "mb23 rc=$?"; cat gpurun_out/microbench_v11_hs23.json
ls gpurun_out/diag_* 2>/dev/null | head; for f in gpurun_out/diag_*.txt; do [ -f "$f" ] && { echo "== $f"; head -12 "$f"; }; done 2>/dev/null | head -150
