"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/summarize_launches.py file.csv"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.3f} |")
print(f"| total | | {tot:.2f} | | |")
