"""Kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv …`):
   python tools/launch_summary.py profiles/launches_r02c.csv > table.md"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(rows)
tot, cnt = defaultdict(float), defaultdict(int)
n = 0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void\s+", "", r["Kernel Name"])
    name = name.replace("<unnamed>::", "")
    name = re.sub(r"[<(].*$", "", name)
    ms = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0}.get(r["Metric Unit"], 1e-6)
    tot[name] += ms
    cnt[name] += 1
    n += 1
total = sum(tot.values())
print(f"total: {n} launches, {total:.1f} ms\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 24]:
    print(f"| `{k}` | {cnt[k]} | {v:.2f} | {100 * v / total:.1f} % |")
