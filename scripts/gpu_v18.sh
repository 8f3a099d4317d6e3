#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "single_rank or resnet50_shapes" --timeout 300 -x 2>&1 | tail -4
DR_SEG_COST=6.0,2.0 timeout 120 python scripts/cta_timeline.py 22 2>&1 | grep -v "slow CTA" | tail -8
for hs in 21 22; do DR_SEG_COST=6.0,2.0 timeout 120 python scripts/engine_microbench.py 20 2 1 $hs > gpurun_out/microbench_v18_hs$hs.json 2>/dev/null; echo "hs=$hs $(python -c "import json;d=json.load(open('gpurun_out/microbench_v18_hs$hs.json'));print(round(d['fused_ms_median'],4), round(d['fused_ms_min'],4), {k:round(v,4) for k,v in list(d['phase_ms_unfused'].items())[:6]})")"; done
