#!/bin/bash
# query_check.sh — L2 persisting-window experiment for the count kernel, an ncu capture of it on an index beyond L2, e2e phase times.
mkdir -p gpurun_out
T=${1:-q}
{
for w in 0 24 48 96; do echo "GCZ_L2_WINDOW_MB=$w"; GCZ_L2_WINDOW_MB=$w timeout -k 5 200 python tools/query_once.py large; done
for w in 0 24; do echo "small GCZ_L2_WINDOW_MB=$w"; GCZ_L2_WINDOW_MB=$w timeout -k 5 200 python tools/query_once.py small; done
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k 'regex:^count_kernel' -s 6 -c 6 -f -o gpurun_out/${T}_count_large python tools/query_once.py large 2>&1 | tail -2
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --block-only --no-cpu-baseline > gpurun_out/${T}_bench_block.json 2> gpurun_out/${T}_bench_block.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_block.json')); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'serial', d['e2e']['serial']['ms_per_step'], d['e2e']['build_call_phases_ms'], 'device phases', d['phases_ms'])"
} > gpurun_out/${T}.log 2>&1
tail -40 gpurun_out/${T}.log
