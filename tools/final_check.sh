#!/bin/bash
# final_check.sh [tag] [gpus] — bench.py (both arms when gpus = 1: the reference arm with 2 steps) + the launch list of bench.py's block leg
mkdir -p gpurun_out
T=${1:-f}
N=${2:-1}
{
if [ "$N" = 1 ]; then
  GCZ_BUILD_TRACE=1 timeout -k 10 300 python bench.py --steps 6 --warmup 3 --block-only --no-cpu-baseline > gpurun_out/${T}_bench_block.json 2> gpurun_out/${T}_bench_block.err; echo "bench block rc=$?"
  grep "gcz build" gpurun_out/${T}_bench_block.err | tail -4
  timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
  timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --block-only --no-cpu-baseline > gpurun_out/${T}_launches_bench.log 2>&1; tail -c 300 gpurun_out/${T}_launches_bench.log
else
  timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
     > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
  grep -v "^\[bench rank [1-9]" gpurun_out/${T}_bench.err | tail -20
fi
python - <<P
import json
for l in open('gpurun_out/${T}_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'scaling', d['scaling'], 'parity', d['parity'].get('ok'))
        if 'build_call_phases_ms' in d['e2e']: print(d['e2e']['build_call_phases_ms'])
        g=d.get('genome') or {}
        print('genome', {k: (v if not isinstance(v, dict) else v.get('value')) for k, v in g.items() if k in ('value','ms_per_step','e2e','blocks_per_rank','lpt_efficiency_bound')})
        if d.get('count'): print('count', {k:v for k,v in d['count'].items() if k in ('value','ms_per_batch','found_somewhere','sharding')}, 'e2e', d['count']['e2e']['value'], 'roofline', d['count']['roofline']['frac'])
        if d.get('locate'): print('locate', {k:v for k,v in d['locate'].items() if k in ('value','ms','occurrences')})
        print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'], 'clocks', d['clocks'])
P
} > gpurun_out/${T}.log 2>&1
tail -40 gpurun_out/${T}.log
