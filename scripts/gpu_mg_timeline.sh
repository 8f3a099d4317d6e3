#!/bin/bash
N=$(nvidia-smi -L | wc -l)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 scripts/cta_timeline.py 22 2>&1 | grep -v "slow CTA\|OMP_NUM\|\*\*\*\*" | tail -20
