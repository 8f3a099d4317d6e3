#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dr_engine_kernel -s 8 -c 6 -o gpurun_out/prof_phases_v17 -f python scripts/engine_microbench.py 3 2 1 22 > gpurun_out/ncu_phases_v17.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_phases_v17.log | cut -c1-200
