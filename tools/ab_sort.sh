#!/bin/bash
# ab_sort.sh — digit-pass variants side by side: sortbench (random 48-bit keys) and the real build.  ab_sort.sh "0 2 8" [tag-unused] "variants to test" [tag]
T=${4:-ab}
mkdir -p gpurun_out
{
for v in $1; do
  echo "GCZ_SORT_VARIANT=$v"; GCZ_SORT_VARIANT=$v timeout -k 5 120 python tools/sortbench.py 248956423 48 || echo "sortbench rc=$?"
  GCZ_SORT_VARIANT=$v timeout -k 5 120 python tools/build_once.py 4 || echo "build_once rc=$?"
done
for v in $3; do
  echo "== sort / suffix tests with GCZ_SORT_VARIANT=$v"
  GCZ_SORT_VARIANT=$v timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 -k "sort or suffix" 2>&1 | tail -3
done
echo "wide status (lean kernel, 64-bit words)"
GCZ_SORT_WIDE_STATUS=1 timeout -k 5 120 python tools/sortbench.py 248956423 48
GCZ_SORT_WIDE_STATUS=1 timeout -k 5 120 python tools/build_once.py 4
} > gpurun_out/${T}.log 2>&1
cat gpurun_out/${T}.log
