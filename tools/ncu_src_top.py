"""Summarise an ncu source-page CSV export: python tools/ncu_src_top.py report.ncu-rep [N]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:160])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
agg = Counter()
for r in data:
    for h in stalls:
        agg[h] += int(r[ix[h]] or 0)
print("samples", tot, "| instructions", len(data), "|", ", ".join(f"{k[6:]}={v}" for k, v in agg.most_common(8)))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
    st = sorted(((h[6:], int(r[ix[h]] or 0)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print(r[ix["Address"]][-5:], r[ix["# Samples"]].rjust(5), r[ix["Instructions Executed"]].rjust(8), r[ix["Source"]][:64].ljust(64), st)
