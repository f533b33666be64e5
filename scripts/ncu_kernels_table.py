#!/usr/bin/env python3
"""Per-kernel table out of an `ncu --csv --metrics ...` log: for every kernel name the number of launches, total time,
and for its LONGEST launch the duration, DRAM bytes moved and achieved DRAM GB/s (against MEASURED_PEAKS.json).
usage: ncu_kernels_table.py kernels.csv"""
import csv, json, os, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
iid, ik, im, iu, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
L = defaultdict(dict)
name = {}
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0, "": 1.0}
for r in rows[1:]:
    if len(r) <= iv:
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    L[r[iid]][r[im]] = v * scale.get(r[iu], 1.0)
    name[r[iid]] = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("mpcr::", "")
peak = 6542.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
K = defaultdict(list)
for i, m in L.items():
    K[name[i]].append(m)
print(f"# per kernel: launches, total ms; longest launch: ms, DRAM read+write MB, achieved DRAM GB/s, fraction of the measured HBM peak ({peak:.0f} GB/s), sm throughput %, grid x block")
print(f"{'kernel':34s} {'n':>5s} {'total ms':>9s} | {'ms':>8s} {'DRAM MB':>9s} {'GB/s':>7s} {'frac':>5s} {'sm%':>5s}  grid x block")
for k, ms in sorted(K.items(), key=lambda kv: -sum(m.get("gpu__time_duration.sum", 0) for m in kv[1])):
    tot = sum(m.get("gpu__time_duration.sum", 0) for m in ms)
    b = max(ms, key=lambda m: m.get("gpu__time_duration.sum", 0))
    t = b.get("gpu__time_duration.sum", 0)
    by = b.get("dram__bytes_read.sum", 0) + b.get("dram__bytes_write.sum", 0)
    gbs = by / (t * 1e-3) / 1e9 if t else 0
    print(f"{k:34s} {len(ms):5d} {tot:9.3f} | {t:8.4f} {by/1e6:9.2f} {gbs:7.0f} {gbs/peak:5.2f} {b.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}  "
          f"{int(b.get('launch__grid_size', 0))} x {int(b.get('launch__block_size', 0))}")
