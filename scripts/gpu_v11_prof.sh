#!/bin/bash
# rest of the GPU test-suite + per-phase ncu capture of the v11 engine
mkdir -p gpurun_out
rm -f gpurun_out/diag_*.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_gpu_engine.py::test_engine_vs_oracle_single_rank 2>&1 | tail -40 > gpurun_out/v11_tests2.log
tail -30 gpurun_out/v11_tests2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dr_engine_kernel -s 8 -c 6 -o gpurun_out/prof_phases_v11 -f python scripts/engine_microbench.py 3 2 1 22 > gpurun_out/ncu_phases_v11.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_phases_v11.log
for f in gpurun_out/diag_*.txt; do [ -f "$f" ] && { echo "== $f"; head -12 "$f"; }; done 2>/dev/null | head -100
