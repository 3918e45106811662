"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`)
as the markdown committed under profiles/: per-kernel launches, total time, share, grid."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head, rows = rows[0], rows[1:]
col = {n: i for i, n in enumerate(head)}


def short(name):
    name = re.sub(r'\bfewbit::', '', name)
    return re.sub(r'\(.*', '', name)[:72]


per = collections.OrderedDict()
launches = []
for r in rows:
    if r[col['Metric Name']] != 'gpu__time_duration.sum':
        continue
    t = float(r[col['Metric Value']].replace(',', ''))
    unit = r[col['Metric Unit']]
    t *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}.get(unit, 1.0)
    name = short(r[col['Kernel Name']])
    launches.append((int(r[col['ID']]), name, t, r[col['Grid Size']]))
    e = per.setdefault(name, [0, 0.0, r[col['Grid Size']]])
    e[0] += 1
    e[1] += t
total = sum(e[1] for e in per.values())
print('| kernel | launches | total µs | share | grid |\n|---|---|---|---|---|')
for name, (n, t, grid) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{name}` | {n} | {t:.1f} | {100 * t / total:.1f}% | {grid} |')
tiles = [l for l in launches if 'tiles_kernel' in l[1]][:12]
print('\nFirst 12 tile-kernel launches, µs:\n\n```')
for i, name, t, grid in tiles:
    print(f'{i:4d} {name:72s} {t:8.1f} grid {grid}')
print('```')
