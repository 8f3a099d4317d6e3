#!/bin/bash
# multi-GPU: exchange microbenchmark (bloom / both / plain) + bench.py with the self-check
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
for cfg in "bloom none" "none none" "bloom polyfit"; do
set -- $cfg
timeout 200 $TR --master-port 29551 scripts/engine_microbench_mg.py 20 $1 $2 2> gpurun_out/mbmg_$1_$2_n$N.err | grep '^{' > gpurun_out/mbmg_$1_$2_n$N.json; echo "microbench $cfg rc=$? $(cut -c1-600 gpurun_out/mbmg_$1_$2_n$N.json)"; tail -2 gpurun_out/mbmg_$1_$2_n$N.err | cut -c1-300
done
timeout 400 $TR --master-port 29552 bench.py --gpus $N --steps 12 --warmup 4 > gpurun_out/bench_mg_n$N.json 2> gpurun_out/bench_mg_n$N.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_mg_n$N.json'))
print({k:d.get(k) for k in ('value','ms_per_step','exchange_ms_per_step','multi_gpu_check','compressed_allgather_bus_gbs','stage2_bytes_per_step_per_rank','gpu_launches')})
print('e2e',d.get('e2e')); print('dense',d.get('dense_allreduce_context')); print('roofline',d.get('roofline')); print('check detail', d.get('multi_gpu_check_detail'))
PY
tail -4 gpurun_out/bench_mg_n$N.err | cut -c1-400
