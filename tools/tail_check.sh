#!/bin/bash
mkdir -p gpurun_out
T=${1:-t}
{
timeout -k 10 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -4
GCZ_BUILD_TRACE=1 timeout -k 10 300 python bench.py --steps 10 --warmup 3 --block-only --no-cpu-baseline > gpurun_out/${T}_bench_block.json 2> gpurun_out/${T}_bench_block.err; echo "bench block rc=$?"
grep "gcz build" gpurun_out/${T}_bench_block.err | tail -4
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_block.json')); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'serial', d['e2e']['serial']['ms_per_step'], d['e2e']['build_call_phases_ms'], d['parity']['ok'])"
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<P
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity'].get('ok'))
g=d['genome']; print('genome', g['value'], g['ms_per_step'], 'e2e', g['e2e']['value'], g['e2e']['ms_per_step'], g['parity'])
print('count', d['count']['value'], d['count']['e2e']['value'], d['count']['roofline']['frac'], 'locate', d['locate']['value'])
P
} > gpurun_out/${T}.log 2>&1
tail -30 gpurun_out/${T}.log
