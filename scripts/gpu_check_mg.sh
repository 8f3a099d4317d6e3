#!/bin/bash
# multi-GPU correctness + short scaling point.  usage: gpu_check_mg.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== multigpu engine vs oracle (N=$N)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/run_multigpu.py > gpurun_out/mg_engine.log 2>&1; echo "rc=$?"; grep -v "^W0\|^\*\*\*" gpurun_out/mg_engine.log | tail -15
echo "== bench ours N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --breakdown > gpurun_out/bench_ours_n$N.json 2> gpurun_out/bench_ours_n$N.err; echo "rc=$?"; cat gpurun_out/bench_ours_n$N.json; tail -5 gpurun_out/bench_ours_n$N.err
echo "== bench dense N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --config dense --no-e2e > gpurun_out/bench_dense_n$N.json 2> gpurun_out/bench_dense_n$N.err; echo "rc=$?"; cat gpurun_out/bench_dense_n$N.json; tail -5 gpurun_out/bench_dense_n$N.err
echo "== bench reference N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 2 --no-e2e > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; cat gpurun_out/bench_ref_n$N.json; tail -5 gpurun_out/bench_ref_n$N.err
