#!/bin/bash
# e2e throughput against the number of workers and the size of the persistent merge kernel
for th in ${THREADS:-512 256}; do for w in ${WORKERS:-2 3}; do
  echo "small threads=$th workers=$w"
  SSG_SMALL_THREADS=$th BENCH_E2E_WORKERS=$w python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('   resident %.1f ms  e2e %.1f ms  (%.0f Mpix/s)' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']))"
done; done
