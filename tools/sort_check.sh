#!/bin/bash
# sort_check.sh — sorter A/B/C + parity tests of everything that sorts (+ optional ncu capture: NCU=1).
mkdir -p gpurun_out
T=${1:-s}
{
for m in 2 1 0; do
  echo "GCZ_SORT_MODE=$m"; GCZ_SORT_MODE=$m timeout -k 5 120 python tools/sortbench.py 248956423 48 || echo "rc=$?"
done
GCZ_SORT_MODE=2 timeout -k 5 60 python tools/sortbench.py 100000 64 || echo "rc=$?"
GCZ_SORT_MODE=2 timeout -k 5 60 python tools/sortbench.py 30000000 64 || echo "rc=$?"
for m in 2; do
  echo "== tests with GCZ_SORT_MODE=$m"
  GCZ_SORT_MODE=$m timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 -k "sort or suffix or build_block or find or count" 2>&1 | tail -3
  GCZ_SORT_MODE=$m timeout -k 10 300 python bench.py --steps 5 --warmup 3 --block-only --no-cpu-baseline > gpurun_out/${T}_bench_block.json 2> gpurun_out/${T}_bench_block.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_block.json')); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['phases_ms'], d['parity']['ok'])"
done
if [ -n "$NCU" ]; then
GCZ_SORT_MODE=2 timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 7 -c 1 -f -o gpurun_out/${T}_sort_scan python tools/sortbench.py 248956423 48 2>&1 | tail -2
fi
} > gpurun_out/${T}.log 2>&1
tail -40 gpurun_out/${T}.log
