#!/usr/bin/env python3
"""Launch list (`ncu --metrics gpu__time_duration.sum --csv`) -> per-kernel totals and the launches of ONE step
(from one scan_kernel launch to the next), with the dominant kernel's share.   usage: ncu_launches_summary.py launches.csv"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ik, im, iu, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
L = []
for r in rows[1:]:
    if len(r) > iv and r[im] == "gpu__time_duration.sum":
        name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("mpcr::", "")
        L.append((name, float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)))
tot = sum(t for _, t in L)
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised) over `python bench.py --steps 2 "
      f"--warmup 1 --no-cpu-baseline --no-e2e`; {len(L)} launches of our kernels, {tot:.3f} ms total")
agg = OrderedDict()
for n, t in L:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t:9.3f} ms {100 * t / tot:5.1f}%  x {c:3d}  {n}")
scans = [i for i, (n, _) in enumerate(L) if n.startswith("scan_kernel")]
if len(scans) >= 2:
    a, b = scans[-2], scans[-1]
    step = L[a:b]
    st = sum(t for _, t in step)
    print(f"# one step (scan_kernel .. next scan_kernel): {len(step)} launches, {st:.3f} ms; scan_kernel share {100 * step[0][1] / st:.1f}%")
    for n, t in step:
        print(f"    {t:9.3f} ms x  1 {n}")
