#!/bin/bash
# ncu launch list of the bench command itself (run under gpurun, one GPU); the plain run first,
# as B200_PROFILING.md requires.  Per-launch times under ncu are cold-cache and serialised:
# compare SHARES with bench.py's own event timing, not absolutes.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/bench_launches_$TAG.csv $CMD > gpurun_out/bench_ncu_$TAG.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | grep bench_ | tail -4
