"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: count, total us, share."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = {c: i for i, c in enumerate(rows[hdr])}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= cols["Metric Value"] or r[cols["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[cols["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    grid = r[cols["Grid Size"]] if "Grid Size" in cols else ""
    val = float(r[cols["Metric Value"]].replace(",", ""))
    unit = r[cols["Metric Unit"]]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    key = name if len(sys.argv) < 3 else f"{name} grid={grid}"
    agg[key][0] += 1
    agg[key][1] += us
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1000:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us / 1000:9.3f} ms {100 * us / tot:5.1f}%  n={n:5d}  avg={us / n:8.1f} us  {k}")
