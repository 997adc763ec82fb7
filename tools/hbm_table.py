"""Merge tools/hbm_kernels.py's CUDA-event table with the ncu DRAM-traffic pass of the same script into one markdown table:
   python tools/hbm_table.py gpurun_out/r02_hbm_kernels.json gpurun_out/r02_hbm_ncu.csv > profiles/r02_hbm_kernels.md"""
import collections
import csv
import json
import sys

tab = json.load(open(sys.argv[1]))
traffic = collections.defaultdict(lambda: collections.defaultdict(list))
if len(sys.argv) > 2:
    lines = open(sys.argv[2]).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')]
    for r in (csv.DictReader(lines[i[0]:]) if i else []):
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("rtsds::", "")
        traffic[name][r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
print(f"| kernel | replaces | algorithmic MB | cold us | achieved GB/s | of {tab['peak_hbm_gbs']:.0f} ({tab['peak_source']}) | warm GB/s | ncu DRAM MB (read+write) | ncu us |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
seen = collections.Counter()
for r in tab["rows"]:
    base = r["kernel"].split("<")[0]
    cands = [k for k in traffic if k.split("<")[0] == base]
    dram = t_ncu = ""
    if cands:
        k = cands[0] if len(cands) == 1 else next((c for c in cands if c.replace(" ", "") == r["kernel"].replace(" ", "")), cands[0])
        # the script launches every case twice, in table order: take this case's second launch
        idx = seen[k] * 2 + 1
        seen[k] += 1
        rd, wr, tt = traffic[k].get("dram__bytes_read.sum", []), traffic[k].get("dram__bytes_write.sum", []), traffic[k].get("gpu__time_duration.sum", [])
        if idx < len(rd):
            dram = f"{(rd[idx] + wr[idx]) / 1e6:.1f}"
            t_ncu = f"{tt[idx] / 1e3:.1f}"
    print(f"| `{r['name']}` | {r['what']} | {r['algorithmic_bytes'] / 1e6:.1f} | {r['cold_us']} | {r.get('achieved_gbs', '')} | "
          f"{r.get('frac_of_hbm_peak', '')} | {r.get('warm_gbs', '')} | {dram} | {t_ncu} |")
