#!/bin/bash
for rw in 0 1 2 3; do for ew in 2 3; do
  echo "resident workers=$rw e2e workers=$ew"
  BENCH_RESIDENT_WORKERS=$rw BENCH_E2E_WORKERS=$ew python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('   resident %.1f ms (%.0f Mpix/s)  e2e %.1f ms  (%.0f Mpix/s)' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value']))"
done; done
