#!/bin/bash
# profile_small.sh — one `ncu --set full` capture of every kernel of a build step that is not a digit pass, in ONE GPU call.
#   gpurun --timeout 900 -- 'bash tools/profile_small.sh'      (one GPU; reports land in gpurun_out/)
# The second build's launches are captured (the first one warms up: allocations, attributes).  Read the reports here with
#   python tools/ncu_summary.py gpurun_out/small_r02.ncu-rep       and       ncu -i ... --page source --csv
# Environment variables select experimental variants as usual (GCZ_EMIT_VARIANT=1 bash tools/profile_small.sh).
set -e
mkdir -p gpurun_out
OUT=${OUT:-gpurun_out/small_r02}
python tools/build_once.py 2 > gpurun_out/build_once_plain.log 2>&1          # must pass without the profiler first
cat gpurun_out/build_once_plain.log
# matching launches per build: bwt_count 1, hswt_emit 1, text_hist 1, group_flags 4, group_apply 4, group_finish 3,
# refine_keys 2, sample 1, run_keys 1, iwt_low_levels 1 = 19; skip the first build's
ncu --set full --clock-control none --import-source on \
    -k 'regex:bwt_count|hswt_emit|text_hist|group_flags|group_apply|group_finish|refine_keys|sample_kernel|run_keys|iwt_low_levels' \
    -s 19 -c 19 -o "$OUT" python tools/build_once.py 2 > gpurun_out/build_once_ncu.log 2>&1 || { tail -20 gpurun_out/build_once_ncu.log; exit 1; }
# and the launch list of one whole step for the shares
# (94 launches per build with the default kernels: any window of 94 consecutive launches after the first build is one whole step)
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 94 --csv --log-file gpurun_out/launches_small_r02.csv \
    python tools/build_once.py 3 > gpurun_out/build_once_list.log 2>&1 || true
ls -la gpurun_out | tail -8
