"""Turn ncu csv exports (launch list + --page raw) into the markdown summaries kept in profiles/.

    python tools/summarize_ncu.py launches <launches.csv> [first-kernel-regex]   > profiles/xxx_launches.md
    python tools/summarize_ncu.py raw <raw.csv> [last]                           > profiles/xxx_full.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = re.sub(r'cub::(?:CUB_\w+::)?(?:detail::\w+::)?', 'cub::', name)
    return name[:70]


def launches(path, firstRe=None):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    if firstRe:   # keep the last repetition: from the last launch matching firstRe
        idx = [i for (i, r) in enumerate(rows) if re.search(firstRe, r[4])]
        if idx:
            rows = rows[idx[-1]:]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        ns = float(r[-1].replace(',', ''))
        ent = agg.setdefault(short(r[4]), [0, 0.0, r[8], r[7]])
        ent[0] += 1
        ent[1] += ns
        total += ns
    print('| kernel | launches | total us | share | grid | block |')
    print('|---|---|---|---|---|---|')
    for (k, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.1f %% | %s | %s |' % (k, v[0], v[1] / 1e3, 100 * v[1] / total, v[2], v[3]))
    print('\n%d launches, %.1f us of kernel time (ncu per-launch times are cold-cache and serialised: '
        'compare shares, not absolutes)' % (len(rows), total / 1e3))


WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
]


def raw(path, last=False):
    rows = list(csv.reader(open(path)))
    (hdr, units) = (rows[0], rows[1])
    ci = dict((h, i) for (i, h) in enumerate(hdr))
    cols = [(h, n) for (h, n) in WANT if h in ci]
    print('| kernel | ' + ' | '.join(n for (_, n) in cols) + ' |')
    print('|---|' + '---|' * len(cols))
    body = rows[2:]
    if last:       # several passes captured: keep the last launch of every kernel (warm caches, sized scratch)
        seen = OrderedDict()
        for r in body:
            seen[(r[ci['Kernel Name']], r[ci['launch__grid_size']] if 'launch__grid_size' in ci else '')] = r
        body = list(seen.values())
    for r in body:
        vals = []
        for (h, _) in cols:
            v = r[ci[h]]
            u = units[ci[h]]
            try:
                f = float(v.replace(',', ''))
                v = ('%.3f' % f).rstrip('0').rstrip('.')
            except ValueError:
                pass
            vals.append('%s %s' % (v, u) if u and u not in ('%',) else v)
        print('| `%s` | ' % short(r[ci['Kernel Name']]) + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        raw(sys.argv[2], last=len(sys.argv) > 3 and sys.argv[3] == 'last')
