#!/bin/bash
# evidence_r02.sh — one GPU call: the whole GPU suite, both bench arms at N = 1, the launch list of a build step, and
# ncu --set full captures of the digit pass (real suffix keys) and of the query kernels.  Logs and reports under gpurun_out/.
mkdir -p gpurun_out
T=${1:-ev}
{
echo "== GPU tests"
timeout -k 10 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
echo "== bench"
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
echo "== reference arm (2 steps)"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/${T}_bench_ref.json
echo "== launch list of build steps"
python tools/build_once.py 2
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python tools/build_once.py 3 > gpurun_out/${T}_launches.log 2>&1; tail -2 gpurun_out/${T}_launches.log
echo "== ncu --set full: a full-size digit pass of a real build (second build, second array pass), text pass, text histogram"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k 'regex:onesweep_kernel|text_hist_kernel' -s 30 -c 4 -f -o gpurun_out/${T}_digit python tools/build_once.py 2 > gpurun_out/${T}_digit.log 2>&1; tail -2 gpurun_out/${T}_digit.log
echo "== query kernels"
for m in small large; do
  timeout -k 5 200 python tools/query_once.py $m
  timeout -k 5 400 ncu --set full --clock-control none --import-source on -k 'regex:::count_kernel<|locate_occurrences_kernel|split_by_string_kernel' -s 2 -c 6 -f -o gpurun_out/${T}_query_$m \
      python tools/query_once.py $m 2>&1 | tail -2
done
} > gpurun_out/${T}.log 2>&1
tail -60 gpurun_out/${T}.log
