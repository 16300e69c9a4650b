"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py <csv> [n_forwards] [min_us] [end-of-forward kernel]
Prints the per-kernel totals of the LAST forward (the list covers n_forwards identical forwards) and every launch above min_us."""
import collections
import csv
import re
import sys

path = sys.argv[1]
nfw = int(sys.argv[2]) if len(sys.argv) > 2 else 3
min_us = float(sys.argv[3]) if len(sys.argv) > 3 else 60.0
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith('==')]))
per = len(rows) // nfw
last = rows[(nfw - 1) * per:]
marker = sys.argv[4] if len(sys.argv) > 4 else None   # a kernel that runs once, at the end of every forward
if marker:
    idx = [i for i, r in enumerate(rows) if marker in r['Kernel Name']]
    if len(idx) >= 2:
        last = rows[idx[-2] + 1:]
        per = len(last)


def us(row):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    return v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)


agg, tot = collections.OrderedDict(), 0.0
for row in last:
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'^void ', '', name)[:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us(row)
    tot += us(row)
print(f"{len(rows)} launches in the list, {per} per forward; last forward: {tot:.1f} us summed (cold-cache, serialised)")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f"{t:10.1f} us {c:5d}  {k}")
print(f"--- launches above {min_us:.0f} us, in order")
for row in last:
    if us(row) > min_us:
        print(f"{us(row):9.1f}  {re.sub(r'^void ', '', row['Kernel Name'])[:100]}  grid={row.get('Grid Size')}")
