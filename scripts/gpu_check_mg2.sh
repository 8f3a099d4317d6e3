#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
echo "== pytest gpu (incl. multi-GPU engine test)"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu_mg.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_mg.log
for cfg in bloom both; do
echo "== bench ours $cfg N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --breakdown --config $cfg > gpurun_out/bench_${cfg}_n$N.json 2> gpurun_out/bench_${cfg}_n$N.err; echo "rc=$?"; grep "^{" gpurun_out/bench_${cfg}_n$N.json | cut -c1-400; grep -o '"exchange_ms_per_step": [0-9.]*' gpurun_out/bench_${cfg}_n$N.json; tail -3 gpurun_out/bench_${cfg}_n$N.err
done
