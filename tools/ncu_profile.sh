#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU).  Every ncu pass is preceded by the
# same command without ncu, as B200_PROFILING.md requires.
set -u
mkdir -p gpurun_out
CMD="python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_quick.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_quick.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_quick2.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:"k_small_persistent|k_ccl_local|k_assign|k_single_decide|k_band_sums" -s 10 -c 10 \
    -o gpurun_out/prof_quick -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
