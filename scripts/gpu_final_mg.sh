#!/bin/bash
# final multi-GPU measurement: oracle check, A/B of the exchange variants, ours / dense / reference
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "== $name rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/$name.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/$name.json | head -1) $(grep -o '"exchange_ms_per_step": [0-9.]*' gpurun_out/$name.json)"
  grep -a "Error\|error" gpurun_out/$name.err | head -3
}
echo "== multigpu engine vs oracle (N=$N, NVLS on)"
DR_NVLS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/run_multigpu.py > gpurun_out/mg_engine_n$N.log 2>&1; echo "rc=$?"; grep -a "MULTIGPU_OK\|MISMATCH\|rror" gpurun_out/mg_engine_n$N.log | head
run bench_ours_n${N} DR_NVLS=0 -- --steps 20 --warmup 5 --breakdown
run bench_ours_nvls_n${N} DR_NVLS=1 -- --steps 20 --warmup 5 --breakdown --no-e2e
run bench_ours_noshard_n${N} DR_SHARD=0 -- --steps 10 --warmup 3 --breakdown --no-e2e
run bench_dense_n${N} A=1 -- --steps 20 --warmup 5 --config dense --no-e2e
run bench_ncf_rle_n${N} A=1 -- --steps 20 --warmup 5 --model ncf --config rle --breakdown --no-e2e
run bench_ref_n${N} A=1 -- --impl reference --steps 3 --warmup 2 --no-e2e

