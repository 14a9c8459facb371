"""Launch list (ncu --metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares, as a markdown table.

    python tools/launch_shares.py profiles/r2_launches.csv "title" > profiles/r2_launch_shares.md
"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
tot = OrderedDict()
for r in rows:
    if r and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(d['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(d.get('Metric Unit', 'ns'), 1e-6)
        name = d['Kernel Name'].split('(')[0].replace('void ', '')[:92]
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
total = sum(v for _, v in tot.values())
print('# %s\n' % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print('Source: `%s` (`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`; the timed steps are' % sys.argv[1])
print('bracketed by cudaProfilerStart/Stop).  Per-launch times under ncu are cold-cache and serialised: compare SHARES with the live CUDA-event')
print('times of the bench line (`entry_points_ms_per_step`), not absolute values.\n')
print('| kernel | launches | total ms | share |\n|---|---|---|---|')
for name, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('| `%s` | %d | %.3f | %.1f %% |' % (name, c, v, 100.0 * v / total))
print('\nTotal %.2f ms.' % total)
