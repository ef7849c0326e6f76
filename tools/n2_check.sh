#!/bin/bash
# n2_check.sh — two GPUs: the NCCL parity tests and bench.py at N=2 (both arms), stderr kept.
mkdir -p gpurun_out
T=${1:-n2}
N=${2:-2}
{
echo "== NCCL tests"
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 300 -k "two_gpus" 2>&1 | tail -5
echo "== bench N=$N"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
   > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
grep -v "^\[bench rank [1-9]" gpurun_out/${T}_bench.err | tail -25
python - <<P
import json
for l in open('gpurun_out/${T}_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'scaling', d['scaling'], 'parity', d['parity'], 'genome', d.get('genome'))
        print('count', {k:v for k,v in d['count'].items() if k in ('value','ms_per_batch','found_somewhere','e2e','sharding')})
        print('locate', {k:v for k,v in d['locate'].items() if k in ('value','ms','occurrences')})
        print('roofline', d['roofline']['frac'], d['roofline']['avg_launch_ms'])
P
} > gpurun_out/${T}.log 2>&1
tail -60 gpurun_out/${T}.log
