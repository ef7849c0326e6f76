#!/bin/bash
# r2b_check.sh — one GPU call: sorter timing, the GPU suite, the block bench, the launch list of a build step.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/r2b_check.sh [tag] [1 = also three small ncu --set full reports]'
T=${1:-r2b}
mkdir -p gpurun_out
{
echo "== sorter"
timeout -k 5 60 python tools/sortbench.py 3000000 48 || echo "sortbench small rc=$?"
timeout -k 5 120 python tools/sortbench.py 248956423 48 || echo "sortbench rc=$?"
echo "== GPU tests"
timeout -k 10 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest.log
echo "== block bench"
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --block-only --no-cpu-baseline > gpurun_out/${T}_bench_block.json 2> gpurun_out/${T}_bench_block.err; echo "bench rc=$?"; tail -5 gpurun_out/${T}_bench_block.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_block.json')); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['phases_ms'], d['parity']['ok'])"
echo "== launch list"
python tools/build_once.py 2
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python tools/build_once.py 3 > gpurun_out/${T}_launches.log 2>&1; tail -2 gpurun_out/${T}_launches.log
python tools/launch_shares.py gpurun_out/${T}_launches.csv 1
if [ -n "$2" ]; then
echo "== ncu --set full (second build of build_once.py): first two digit passes; text histogram + first grouping; refinement keys + IWT passes"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout -k 10 300 $NCU -k regex:onesweep_kernel -s 26 -c 2 -o gpurun_out/${T}_passes python tools/build_once.py 2 2>&1 | tail -1
timeout -k 10 300 $NCU -k 'regex:text_hist_kernel|group_flags_kernel<1>|group_apply_kernel<1>' -s 3 -c 3 -o gpurun_out/${T}_group python tools/build_once.py 2 2>&1 | tail -1
timeout -k 10 300 $NCU -k 'regex:refine_keys_kernel|iwt_top_emit_kernel' -s 4 -c 4 -o gpurun_out/${T}_refine python tools/build_once.py 2 2>&1 | tail -1
fi
} > gpurun_out/${T}.log 2>&1
tail -100 gpurun_out/${T}.log
