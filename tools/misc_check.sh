#!/bin/bash
# misc_check.sh — sorter on non-i.i.d. texts, the reference arm at N > 1 (one short step), the file-level native path (cold and warm)
mkdir -p gpurun_out
T=${1:-m}
{
echo "== sorter generality (64 Mbp each)"
timeout -k 10 600 python tools/sorter_generality.py 64000000
echo "== reference arm as the driver runs it at N = 2 (rank 0 only), one step"
WORLD_SIZE=2 RANK=0 timeout -k 10 900 python bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --ref-budget 60 > gpurun_out/${T}_ref_n2.json 2> gpurun_out/${T}_ref_n2.err; echo "rc=$?"; cut -c1-900 gpurun_out/${T}_ref_n2.json; tail -3 gpurun_out/${T}_ref_n2.err
echo "== native whole-genome path, three runs in one process"
GCZ_HOST_TRACE=0 timeout -k 10 900 python tools/hostbench.py --engine cuda --scale 1.0 --out /dev/shm --repeat 3 2> gpurun_out/${T}_hostbench.err | tail -12
rm -f /dev/shm/hostbench_*.fa
} > gpurun_out/${T}.log 2>&1
tail -40 gpurun_out/${T}.log
