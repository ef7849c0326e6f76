#!/bin/bash
# profile_r02.sh — ncu captures of the digit pass (both kernels) and of the query kernels; reports under gpurun_out/.
mkdir -p gpurun_out
{
for p in 1 0; do
  K=onesweep_pairs; [ $p = 0 ] && K=onesweep_kernel
  GCZ_SORT_PERSISTENT=$p timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 7 -c 1 -f -o gpurun_out/r02_sort_p$p \
      python tools/sortbench.py 248956423 48 2>&1 | tail -3
done
for m in small large; do
  timeout -k 5 200 python tools/query_once.py $m
  timeout -k 5 400 ncu --set full --clock-control none --import-source on -k 'regex:::count_kernel<|locate_occurrences_kernel|split_by_string_kernel' -s 2 -c 6 -f -o gpurun_out/r02_query_$m \
      python tools/query_once.py $m 2>&1 | tail -3
done
} > gpurun_out/profile_r02.log 2>&1
tail -30 gpurun_out/profile_r02.log
