#!/bin/bash
# Hardware runs left pending at the end of round 1.  Usage (from the repo root):
#   gpurun --timeout 900 -- 'bash scripts/round2_first_runs.sh 1'          # single GPU: sanitizer over every fused mode
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/round2_first_runs.sh 2'  # 2 GPUs: nccl transport, own-flags at W>1, allgather microbench
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  bash scripts/gpu_sanitize.sh
  for o in "1 0" "0 1"; do set -- $o
    echo "== memcheck with DR_OWN_FLAGS=$1 DR_EMIT_COUNTS=$2"
    DR_OWN_FLAGS=$1 DR_EMIT_COUNTS=$2 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_engine.py > gpurun_out/sanitizer_memcheck_own$1_cnt$2.log 2>&1
    grep -E "ERROR SUMMARY|SANITIZE_RUN_DONE|matches_oracle=False" gpurun_out/sanitizer_memcheck_own$1_cnt$2.log | head -5
  done
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  echo "== oracle test incl. the NCCL transport"
  DR_TEST_NCCL_TRANSPORT=1 timeout 600 $TR --master-port 29541 tests/run_multigpu.py > gpurun_out/mg_nccl_transport_n$N.log 2>&1; echo "rc=$?"
  grep -a "MULTIGPU_OK\|MISMATCH\|rror" gpurun_out/mg_nccl_transport_n$N.log | head
  echo "== oracle test with DR_OWN_FLAGS=1 at W=$N"
  DR_OWN_FLAGS=1 timeout 600 $TR --master-port 29542 tests/run_multigpu.py > gpurun_out/mg_own_flags_n$N.log 2>&1; echo "rc=$?"
  grep -a "MULTIGPU_OK\|MISMATCH\|rror" gpurun_out/mg_own_flags_n$N.log | head
  echo "== compressed-allgather microbench"
  timeout 300 $TR --master-port 29543 scripts/allgather_microbench.py 50 > gpurun_out/allgather_microbench_n$N.json 2> gpurun_out/allgather_microbench_n$N.err; echo "rc=$?"
  grep "^{" gpurun_out/allgather_microbench_n$N.json | cut -c1-600
  echo "== bench with the NCCL transport (for the multi-host path's cost on one box)"
  DR_TRANSPORT=nccl timeout 300 $TR --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --breakdown --no-e2e 2> gpurun_out/bench_nccl_transport_n$N.err | grep "^{" > gpurun_out/bench_nccl_transport_n$N.json
  grep -o '"value": [0-9.]*\|"exchange_ms_per_step": [0-9.]*' gpurun_out/bench_nccl_transport_n$N.json
fi
