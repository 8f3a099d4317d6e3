#!/bin/bash
mkdir -p gpurun_out
for sc in "2.0,1.0" "3.0,1.0" "4.0,1.5" "6.0,2.0"; do
echo "=== DR_SEG_COST=$sc"
DR_SEG_COST=$sc timeout 120 python scripts/cta_timeline.py 22 2>&1 | grep -v "slow CTA" | tail -8
DR_SEG_COST=$sc timeout 120 python scripts/engine_microbench.py 20 2 1 22 2>/dev/null | python -c "import json,sys;d=json.load(sys.stdin);print('fused', round(d['fused_ms_median'],4), round(d['fused_ms_min'],4), {k:round(v,4) for k,v in list(d['phase_ms_unfused'].items())[:6]})"
done
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "single_rank or resnet50_shapes" --timeout 300 -x 2>&1 | tail -4
