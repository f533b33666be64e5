#!/usr/bin/env python3
"""Group the SASS of the profiled kernel into regions of equal execution count: share of instructions and of
warp-stall samples per region.  usage: ncu_regions.py report.ncu-rep [units]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
E = lambda r: int(r[ix['Instructions Executed']]); S = lambda r: int(r[ix['# Samples']])
tot = sum(E(r) for r in data); ts = sum(S(r) for r in data)
print('total warp instr', tot, ('per unit %.1f' % (tot / units)) if units else '')
regions = []; prev = None; start = 0; acc = 0; samp = 0
for k, r in enumerate(data):
    e = E(r)
    if prev is None or abs(e - prev) > 0.02 * max(e, prev, 1):
        if prev is not None: regions.append((start, k - 1, prev, acc, samp))
        start = k; acc = 0; samp = 0
    acc += e; samp += S(r); prev = e
regions.append((start, len(data) - 1, prev, acc, samp))
for st, en, e, acc, samp in regions:
    if acc > 0.01 * tot or samp > 0.015 * ts:
        print(f"sass[{st:4d}-{en:4d}] n={en-st+1:4d} exec={e:>10d} instr%={100*acc/tot:5.1f} samples%={100*samp/ts:5.1f}  {data[st][1].strip()[:44]} .. {data[en][1].strip()[:44]}")
