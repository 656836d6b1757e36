#!/bin/bash
# sweep of the persistent merge kernel's grid shape (run under gpurun)
for th in 128 256 512; do for b in 1 2 3 4; do
  echo "threads=$th blocks/SM=$b"
  SSG_SMALL_THREADS=$th SSG_SMALL_BLOCKS_PER_SM=$b python tools/prof_tile.py ${1:-4096} ${1:-4096} 4 3 2>&1 | tail -1
done; done
