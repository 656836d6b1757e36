#!/bin/bash
# ncu --set full on the two kernels the north star names (k_assign, k_ccl_local), one 4096 tile
set -u
TAG=${1:-r1}
CMD="python tools/prof_tile.py 4096 4096 4 1"
$CMD > gpurun_out/two_plain_$TAG.log 2>&1 || { echo plain run failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"k_assign|k_ccl_local|k_gather_ids|k_band_sums|k_tile_extents" -s 4 -c 5 \
    -o gpurun_out/two_$TAG -f $CMD > gpurun_out/two_ncu_$TAG.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/two_$TAG.ncu-rep --page raw --csv > gpurun_out/two_raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/two_$TAG.ncu-rep --page details --csv > gpurun_out/two_details_$TAG.csv 2>/dev/null
ls -la gpurun_out | grep two_
