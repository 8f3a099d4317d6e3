#!/bin/bash
# final single-GPU evidence: bench line, microbench (index / both), per-phase ncu --set full, launch list of a bench run
mkdir -p gpurun_out
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_final_n1.json)"; cut -c1-330 gpurun_out/bench_final_n1.json
timeout 120 python scripts/engine_microbench.py 20 2 > gpurun_out/microbench_v8_index.json 2>/dev/null; cut -c1-900 gpurun_out/microbench_v8_index.json
timeout 120 python scripts/engine_microbench.py 20 2 1 23 1 polyfit > gpurun_out/microbench_v8_both.json 2>/dev/null; cut -c1-900 gpurun_out/microbench_v8_both.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dr_engine_kernel -s 8 -c 16 -o gpurun_out/prof_phases_v8 -f python scripts/engine_microbench.py 3 2 > gpurun_out/ncu_phases_v8.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_phases_v8.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/launches_bench.out 2>&1; echo "launch list rc=$? rows=$(wc -l < gpurun_out/launches_bench.csv)"
