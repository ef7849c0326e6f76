#!/bin/bash
# round2_first_call.sh — what the first GPU call of the next round should run (ONE GPU, ~15 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Everything written after the GPU budget of round 1 ran out gets its first hardware run here, with stderr kept.
# Each step goes on when the previous one fails: the logs in gpurun_out/ say which did.
mkdir -p gpurun_out
L=gpurun_out/r2_first
{
echo "== 1. the GPU test suite (default kernels; includes the rewritten writers and the native host layer)"
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu.log 2>&1; tail -40 gpurun_out/r2_pytest_gpu.log
echo "== 2. the test that was waiting for hardware"
GCZ_TEST_PENDING=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k native_callers > gpurun_out/r2_pytest_pending.log 2>&1; tail -30 gpurun_out/r2_pytest_pending.log
echo "== 3. bench, default kernels"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 600 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
echo "== 4. every prepared variant: parity against the default kernels, then time"
timeout 1500 python tools/variants.py --steps 5 --out gpurun_out/variants_r02.jsonl 2>&1 | tail -25
echo "== 5. e2e with the marker vector copied out early"
GCZ_EARLY_MARKER=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --patterns 100000 > gpurun_out/r2_bench_early_marker.json 2> gpurun_out/r2_bench_early_marker.err; tail -c 400 gpurun_out/r2_bench_early_marker.json
echo "== 6. the native whole-genome path (FASTA written here: ~40 s of host work), per-block times on stderr"
GCZ_HOST_TRACE=1 timeout 900 python tools/hostbench.py --engine cuda --scale 1.0 --out /dev/shm 2> gpurun_out/r2_hostbench_trace.log | tail -8
rm -f /dev/shm/hostbench_*.fa
} > $L.log 2>&1
tail -60 $L.log
