#!/bin/bash
# gpu_check.sh — one GPU call: sorter A/B, the GPU test suite, a small and a full bench run (logs under gpurun_out/).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
T=${1:-chk}
mkdir -p gpurun_out
{
echo "== sorter, small then full size, persistent kernel on / off"
for p in 1 0; do
  GCZ_SORT_PERSISTENT=$p timeout -k 5 60 python tools/sortbench.py 3000000 48 || echo "sortbench small persistent=$p rc=$?"
done
for p in 1 0; do
  GCZ_SORT_PERSISTENT=$p timeout -k 5 120 python tools/sortbench.py 248956423 48 || echo "sortbench persistent=$p rc=$?"
done
echo "== GPU tests"
timeout -k 10 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${T}_pytest.log
echo "== bench, scaled down"
timeout -k 10 400 python bench.py --steps 2 --warmup 1 --scale 0.02 --length 5000000 --count-patterns 400000 --locate-patterns 100000 --no-cpu-baseline \
   > gpurun_out/${T}_bench_small.json 2> gpurun_out/${T}_bench_small.err; echo "bench small rc=$?"; tail -5 gpurun_out/${T}_bench_small.err; tail -c 1500 gpurun_out/${T}_bench_small.json
echo "== bench, full"
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -30 gpurun_out/${T}_bench.err; tail -c 3000 gpurun_out/${T}_bench.json
} > gpurun_out/${T}.log 2>&1
tail -120 gpurun_out/${T}.log
