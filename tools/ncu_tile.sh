#!/bin/bash
# ncu evidence on one full-size tile (run under gpurun, one GPU): plain run first, then the
# launch list, then one --set full capture of the second pass's kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
SIZE=${2:-4096}
SKIP=${3:-0}       # k_ kernels to skip / capture in the full pass (two passes of about 40 each)
COUNT=${4:-90}
CMD="python tools/prof_tile.py $SIZE $SIZE 4 1"
$CMD > gpurun_out/tile_plain_$TAG.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/tile_plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/tile_launches_$TAG.csv $CMD > gpurun_out/tile_ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none \
    -k regex:"^k_" -s $SKIP -c $COUNT \
    -o gpurun_out/tile_full_$TAG -f $CMD > gpurun_out/tile_ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/tile_full_$TAG.ncu-rep --page raw --csv > gpurun_out/tile_full_raw_$TAG.csv 2>/dev/null
# the report itself is too large to travel (gpurun_out is capped at 64 MiB): keep the csv export
if [ $(stat -c %s gpurun_out/tile_full_$TAG.ncu-rep 2>/dev/null || echo 0) -gt 30000000 ]; then rm -f gpurun_out/tile_full_$TAG.ncu-rep; fi
ls -la gpurun_out | tail -6
