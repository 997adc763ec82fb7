"""Summarise an ncu gpu__time_duration launch list: python tools_summarize.py <csv> [marker-kernel-substring] [step-index]"""
import collections
import csv
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else None
which = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lines = open(path).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[i:]))
names = [r['Kernel Name'].split('(')[0] for r in rows]
a, b = 0, len(rows)
if marker:
    idx = [k for k, n in enumerate(names) if marker in n]
    a, b = idx[which], idx[which + 1]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[a:b]:
    k = r['Kernel Name'].split('(')[0].replace('void ', '').replace('rtsds::', '')[:70]
    agg[k][0] += 1
    agg[k][1] += float(r['Metric Value']) / 1000
tot = sum(v[1] for v in agg.values())
print(f"launches {b - a}, total {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]:10.1f} us {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")
if '-v' in sys.argv:
    for r in rows[a:b]:
        t = float(r['Metric Value']) / 1000
        if t > 40:
            print(f"{t:8.1f} {r['Kernel Name'].split('(')[0][:60]:60s} grid={r['Grid Size']}")
