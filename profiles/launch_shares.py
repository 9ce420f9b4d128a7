"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys

lines = [ln for ln in open(sys.argv[1]) if not ln.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("mpbp::", "")
    t = float(row["Metric Value"].replace(",", ""))
    t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[row["Metric Unit"]]
    agg[(name, row["Grid Size"])][0] += 1
    agg[(name, row["Grid Size"])][1] += t
    tot += t
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg={v[1] / v[0]:8.1f} us  {k[0]} grid={k[1]}")
