#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "tf_compat or trainer_cuda or state_roundtrip" --timeout 200 -x 2>&1 | tail -8
timeout 120 python scripts/cta_timeline.py 22 2>&1 | tail -40 | tee gpurun_out/cta_timeline_v16.txt
for hs in 21 22; do timeout 120 python scripts/engine_microbench.py 20 2 1 $hs > gpurun_out/microbench_v16_hs$hs.json 2>/dev/null; echo "hs=$hs $(python -c "import json;d=json.load(open('gpurun_out/microbench_v16_hs$hs.json'));print(round(d['fused_ms_median'],4), round(d['fused_ms_min'],4), {k:round(v,4) for k,v in list(d['phase_ms_unfused'].items())[:6]})")"; done
