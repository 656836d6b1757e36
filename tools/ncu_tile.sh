#!/bin/bash
# ncu evidence on one full-size tile (run under gpurun, one GPU): plain run first, then the
# launch list, then one --set full capture of the second pass's kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
SIZE=${2:-4096}
CMD="python tools/prof_tile.py $SIZE $SIZE 4 1"
$CMD > gpurun_out/tile_plain_$TAG.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/tile_plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/tile_launches_$TAG.csv $CMD > gpurun_out/tile_ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:"^k_" -s 29 -c 28 \
    -o gpurun_out/tile_full_$TAG -f $CMD > gpurun_out/tile_ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/tile_full_$TAG.ncu-rep --page raw --csv > gpurun_out/tile_full_raw_$TAG.csv 2>/dev/null
ls -la gpurun_out | tail -6
