"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel/grid count, mean, total."""
import collections
import csv
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    k = row["Kernel Name"][:70] + " grid=" + row["Grid Size"]
    agg.setdefault(k, []).append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:100s} n={len(v):3d} avg={sum(v)/len(v)/1000:9.1f}us total={sum(v)/1000:9.1f}us {100*sum(v)/tot:5.1f}%")
print(f"total {tot/1000:.1f} us over {n - skip} launches")
