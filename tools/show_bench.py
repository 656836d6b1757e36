"""Print the scaling-relevant fields of a bench.py JSON line (per-rank host timers included)."""
import json
import sys

line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value %.0f  ms/step %.2f  e2e %.0f (%.1f ms)  parity_vs_one_rank %s  parity_vs_oracle %s' % (
    line['value'], line['ms_per_step'], line['e2e']['value'], line['e2e']['ms_per_step'],
    line.get('parity_vs_one_rank'), line.get('parity_vs_oracle')))
if line.get('one_rank'):
    print(line['one_rank'])
for (i, r) in enumerate(line.get('per_rank') or []):
    print('rank', i, r['tiles'], 'tiles', r['tile_mpix'], 'Mpix', r.get('step_wall_ms'))
    print('   resident', r['resident'])
    print('   e2e     ', r['e2e'])
