"""Development aid: summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list.
  python tests/ncu_launches.py gpurun_out/launches.csv [--seq]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
seq = []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        seq.append((d["Kernel Name"].split("(")[0], d["Grid Size"], float(d["Metric Value"].replace(",", "")) / 1e3))
if "--seq" in sys.argv:
    for k, g, v in seq:
        print(f"{k:28s} grid={g:>16s} {v:9.1f} us")
agg = collections.OrderedDict()
for k, g, v in seq:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {k} | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0] / 1e3:.3f} |")
print(f"| **total** | {len(seq)} | {tot / 1e3:.3f} | | |")
