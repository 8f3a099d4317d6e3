#!/bin/bash
# sharded-decode A/B on N GPUs: oracle test, then bench breakdown with DR_SHARD=1/0
N=${1:-2}
mkdir -p gpurun_out
echo "== run_multigpu"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/run_multigpu.py > gpurun_out/mg_shard_n$N.log 2>&1; echo "rc=$?"; grep -a "MULTIGPU_OK\|MISMATCH\|Error\|error" gpurun_out/mg_shard_n$N.log | head -20
for sh in 1 0; do
for cfg in bloom both; do
echo "== bench ours $cfg N=$N DR_SHARD=$sh"
DR_SHARD=$sh timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --breakdown --no-e2e --config $cfg > gpurun_out/bench_${cfg}_shard${sh}_n$N.json 2> gpurun_out/bench_${cfg}_shard${sh}_n$N.err; echo "rc=$?"; grep "^{" gpurun_out/bench_${cfg}_shard${sh}_n$N.json | cut -c1-200; grep -o '"exchange_ms_per_step": [0-9.]*' gpurun_out/bench_${cfg}_shard${sh}_n$N.json; tail -2 gpurun_out/bench_${cfg}_shard${sh}_n$N.err
done
done
