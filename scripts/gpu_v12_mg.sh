#!/bin/bash
# multi-GPU oracle run (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29541 tests/run_multigpu.py > gpurun_out/mg_v12_n$N.log 2>&1; echo "rc=$?"
grep -a "MULTIGPU_OK\|MISMATCH\|rror\|fault\|engine index" gpurun_out/mg_v12_n$N.log | head -60
tail -5 gpurun_out/mg_v12_n$N.log
