#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/diag_*.txt
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "single_rank or resnet50_shapes" --timeout 300 -x 2>&1 | tail -15 > gpurun_out/v15_tests.log
tail -6 gpurun_out/v15_tests.log
for tma in 1 0; do
timeout 120 python scripts/engine_microbench.py 20 2 $tma 22 > gpurun_out/microbench_v15_tma$tma.json 2> gpurun_out/microbench_v15_tma$tma.err; echo "tma=$tma rc=$?"; cat gpurun_out/microbench_v15_tma$tma.json | cut -c1-700; tail -3 gpurun_out/microbench_v15_tma$tma.err
done
# raw HBM capability for the same traffic pattern with library kernels: r += g ; g = 0  (409 MB)
python - <<'PY'
import torch
n = 25557032
g = torch.randn(n, device='cuda'); r = torch.randn(n, device='cuda')
flush = torch.empty(64*1024*1024, device='cuda')
def run():
    r.add_(g); g.zero_()
for _ in range(5): run()
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("torch r.add_(g); g.zero_() median ms", ts[10], "min", ts[0], "-> GB/s", 4*4*n/ts[10]/1e6)
c = torch.empty(n, device='cuda')
def cp(): c.copy_(g)
for _ in range(5): cp()
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cp(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); print("torch copy 102MB median ms", ts[10], "-> GB/s", 2*4*n/ts[10]/1e6)
PY
for f in gpurun_out/diag_*.txt; do [ -f "$f" ] && { echo "== $f"; head -12 "$f"; }; done 2>/dev/null | head -40
