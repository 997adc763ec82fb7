"""Trim an ncu gpu__time_duration launch list to ONE step: rows between two consecutive occurrences of a marker kernel.
usage: python tools/trim_launches.py <in.csv> <marker-substring> <occurrence-index> <out.csv>"""
import sys

src, marker, which, dst = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
lines = open(src).read().splitlines()
h = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = lines[h + 1:]
idx = [k for k, l in enumerate(rows) if marker in l]
a, b = idx[which], idx[which + 1]
with open(dst, "w") as f:
    f.write(lines[h] + "\n")
    f.write("\n".join(rows[a:b]) + "\n")
print(f"{dst}: {b - a} launches")
