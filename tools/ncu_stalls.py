"""Top stall reasons (warps stalled per issue-active cycle) of every kernel in an ncu report's raw csv.
    ncu -i x.ncu-rep --page raw --csv | python tools/ncu_stalls.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
ci = dict((h, i) for (i, h) in enumerate(hdr))
cols = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    kn = r[ci['Kernel Name']].split('(')[0]
    vals = []
    for c in cols:
        try:
            vals.append((float(r[ci[c]].replace(',', '')), c[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        except ValueError:
            pass
    vals.sort(reverse=True)
    print('%-36s %s' % (kn[:36], '  '.join('%s %.1f' % (n, v) for (v, n) in vals[:4])))
