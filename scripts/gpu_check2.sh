#!/bin/bash
mkdir -p gpurun_out
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for bps in 1 2; do
echo "== microbench bps=$bps"; timeout 300 python scripts/engine_microbench.py 20 $bps > gpurun_out/microbench_bps$bps.json 2> gpurun_out/microbench_bps$bps.err; echo "rc=$?"; cat gpurun_out/microbench_bps$bps.json; tail -3 gpurun_out/microbench_bps$bps.err
done
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "== bench ours no-overlap"; timeout 900 python bench.py --steps 10 --warmup 3 --no-overlap --no-e2e > gpurun_out/bench_ours_noov.json 2> gpurun_out/bench_ours_noov.err; echo "rc=$?"; cat gpurun_out/bench_ours_noov.json; tail -5 gpurun_out/bench_ours_noov.err
echo "== bench dense"; timeout 900 python bench.py --steps 10 --warmup 3 --config dense --no-e2e > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; echo "rc=$?"; cat gpurun_out/bench_dense.json; tail -5 gpurun_out/bench_dense.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
echo "== ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:dr_engine_kernel -s 3 -c 1 -o gpurun_out/prof_engine2 -f python scripts/engine_microbench.py 3 1 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
ls -la gpurun_out
