"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name, share of the step."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
print(f"total {total/1e3:.3f} ms over {sum(v[0] for v in tot.values())} launches")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us/1e3:9.3f} ms {100*us/total:5.1f}%  n={n:4d}  avg {us/n:8.1f} us  {name[:110]}")
