#!/bin/bash
# ncu --set full on selected kernels of one tile: tools/ncu_kernels.sh <tag> <regex> [size] [skip] [count]
set -u
mkdir -p gpurun_out
TAG=$1; RE=$2; SIZE=${3:-4096}; SKIP=${4:-1}; COUNT=${5:-2}
CMD="python tools/prof_tile.py $SIZE $SIZE 4 1"
$CMD > gpurun_out/k_plain_$TAG.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/k_plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $COUNT \
    -o gpurun_out/k_$TAG -f $CMD > gpurun_out/k_ncu_$TAG.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/k_$TAG.ncu-rep --page raw --csv > gpurun_out/k_raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/k_$TAG.ncu-rep --page details --csv > gpurun_out/k_details_$TAG.csv 2>/dev/null
ls -la gpurun_out | grep "k_.*$TAG"
