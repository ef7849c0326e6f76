#!/bin/bash
# e2e_check.sh — where the e2e milliseconds go: staging piece sizes, host wall clock per stage; count kernel occupancy A/B; find after the change
mkdir -p gpurun_out
T=${1:-e}
{
for c in 4 1 16 64 512; do
  echo "== GCZ_STAGE_CHUNK_MB=$c"
  GCZ_STAGE_CHUNK_MB=$c GCZ_BUILD_TRACE=1 timeout -k 10 300 python bench.py --steps 6 --warmup 3 --block-only --no-cpu-baseline > gpurun_out/${T}_bench_block.json 2> gpurun_out/${T}_bench_block_$c.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_block.json')); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'serial', d['e2e']['serial']['ms_per_step'], d['e2e']['build_call_phases_ms'])"
  grep "gcz build" gpurun_out/${T}_bench_block_$c.err | tail -4
done
echo "== count kernel: 5 vs 6 CTAs per SM"
for o in 0 1; do GCZ_COUNT_OCC6=$o timeout -k 5 200 python tools/query_once.py large; done
for o in 0 1; do GCZ_COUNT_OCC6=$o timeout -k 5 200 python tools/query_once.py small; done
echo "== query tests"
timeout -k 10 600 python -m pytest tests -q -m gpu -x --timeout 300 -k "find or count or cfg4 or match or gff" 2>&1 | tail -3
timeout -k 10 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print('locate', d['locate']['value'], d['locate']['ms'], 'count', d['count']['value'], d['count']['e2e']['value'])"
} > gpurun_out/${T}.log 2>&1
tail -70 gpurun_out/${T}.log
