"""Turn an `ncu --set full` report into a small JSON / text summary (run HERE, no GPU needed):

  python tools/ncu_summary.py gpurun_out/r02_targets.ncu-rep profiles/r02_ncu_kernels.json

Per profiled launch: kernel name, duration, DRAM bytes read / written, DRAM throughput %, tensor-pipe activity %, achieved
occupancy, registers -- the figures DESIGN.md section 5 and bench.py's roofline.traffic quote.
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    ("duration_us", r"^gpu__time_duration\.sum$"),
    ("dram_read_bytes", r"^dram__bytes_read\.sum$"),
    ("dram_write_bytes", r"^dram__bytes_write\.sum$"),
    ("dram_throughput_pct", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$"),
    ("tensor_pipe_pct", r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active$"),
    ("tensor_pipe_pct_alt", r"^sm__inst_executed_pipe_tensor.*pct_of_peak_sustained_active$"),
    ("sm_throughput_pct", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$"),
    ("warps_active_pct", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$"),
    ("registers_per_thread", r"^launch__registers_per_thread$"),
    ("grid", r"^launch__grid_size$"),
    ("block", r"^launch__block_size$"),
    ("smem_dynamic", r"^launch__shared_mem_per_block_dynamic$"),
    ("l2_hit_pct", r"^lts__t_sector_hit_rate\.pct$"),
]
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "second": 1e6, "s": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0}


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    kn = hdr.index("Kernel Name")
    cols = {}
    for key, pat in KEEP:
        for j, h in enumerate(hdr):
            if re.search(pat, h):
                cols.setdefault(key, j)
    tensor_like = [h for h in hdr if "tensor" in h]
    res = []
    for r in rows[hdr_i + 2:]:
        if len(r) <= kn:
            continue
        d = {"kernel": re.sub(r"\(.*", "", r[kn])[:120]}
        for key, j in cols.items():
            try:
                v = float(r[j].replace(",", ""))
            except ValueError:
                continue
            u = units[j]
            if key == "duration_us":
                v *= UNIT_SCALE.get(u, 1.0)
            elif key.endswith("_bytes"):
                v *= UNIT_SCALE.get(u, 1.0)
            d[key] = v
        if "duration_us" in d and "dram_read_bytes" in d:
            d["dram_GBps"] = (d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)) / d["duration_us"] / 1e3
        res.append(d)
    txt = json.dumps({"report": rep, "tensor_metric_columns": tensor_like[:12], "launches": res}, indent=1)
    if out:
        open(out, "w").write(txt)
    for d in res:
        print(f"{d['kernel'][:70]:70s} {d.get('duration_us', 0):9.1f} us  dram {d.get('dram_read_bytes', 0) / 1e6:8.1f}+{d.get('dram_write_bytes', 0) / 1e6:8.1f} MB"
              f" ({d.get('dram_GBps', 0):6.0f} GB/s, {d.get('dram_throughput_pct', 0):5.1f}%)  tensor {d.get('tensor_pipe_pct', d.get('tensor_pipe_pct_alt', float('nan'))):5.1f}%"
              f"  regs {int(d.get('registers_per_thread', 0))}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
