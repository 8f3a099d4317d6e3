#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 tests/run_multigpu.py > gpurun_out/mg_v20_n$N.log 2>&1; echo "oracle rc=$?"
grep -a "MULTIGPU_OK\|MISMATCH\|rror\|fault" gpurun_out/mg_v20_n$N.log | head -30
timeout 200 $TR --master-port 29561 scripts/cta_timeline.py 22 2>&1 | grep -v "slow CTA\|OMP_NUM\|\*\*\*\*" | tail -14
for cfg in "bloom none" "bloom polyfit"; do
set -- $cfg
timeout 200 $TR --master-port 29551 scripts/engine_microbench_mg.py 20 $1 $2 2> gpurun_out/mbmg_v20_$1_$2_n$N.err | grep '^{' > gpurun_out/mbmg_v20_$1_$2_n$N.json; echo "microbench $cfg rc=$? $(cut -c1-330 gpurun_out/mbmg_v20_$1_$2_n$N.json)"
done
