#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep: duration, dram bytes, L2 hit, issue utilisation, stall breakdown, hottest lines.
usage: ncu_summary.py report.ncu-rep [n_source_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
def g(k):
    return d.get(k, ("?", ""))
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
for k in keys:
    v, u = g(k)
    print(f"{k:75s} {v} {u}")
st = []
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            st.append((float(d[h][0].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
print("stall cycles per issued instruction:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:10]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if n:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[0]
    try:
        i_src = h.index("Source"); i_smp = h.index("# Samples") if "# Samples" in h else h.index("Sampling Data (All)")
    except ValueError:
        print(h); sys.exit(0)
    out = []
    for r in rows[1:]:
        try:
            out.append((int(r[i_smp].replace(",", "") or 0), r[0], r[i_src].strip()[:110]))
        except (ValueError, IndexError):
            pass
    tot = sum(x[0] for x in out) or 1
    for smp, ln, text in sorted(out, reverse=True)[:n]:
        print(f"{100.0*smp/tot:5.1f}%  L{ln:>5s}  {text}")
