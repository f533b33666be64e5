"""Where does the end-to-end step spend its time?  (upload = H2D + pack, scan, D2H)"""
import os, sys, time
sys.path[:0] = ['.', 'tests']
import numpy as np, torch
import synth
from merpcr_b200 import MerPCR
dev = torch.device('cuda', 0)
lengths = synth.GRCH38_LENGTHS
sts = synth.make_sts_set(8, 100000)
open('/tmp/x.sts', 'wb').write(synth.sts_lines(sts))
eng = MerPCR(mismatches=1, device=0)
eng.load_sts_file('/tmp/x.sts')
lay = eng.make_layout(lengths)
host = []
for ci, L in enumerate(lengths):
    t = synth.dna_torch(1000 + ci, 0, L, dev)
    h = torch.empty(L, dtype=torch.uint8).pin_memory(); h.copy_(t); host.append(h)
torch.cuda.synchronize()
print('pinned?', host[0].is_pinned(), host[0][1000:2000].is_pinned())
sh = None
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sh = eng.upload(lay, host, shard=sh)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = eng.scan(lay, sh)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):.1f} ms ({sum(lengths)/(t1-t0)/1e9:.1f} GB/s), scan+sort+d2h {1e3*(t2-t1):.2f} ms, hits {len(out)}", flush=True)
# raw: one .to() per contig, no pack
torch.cuda.synchronize(); t0 = time.perf_counter()
keep = [h.to(dev, non_blocking=True) for h in host]
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"raw .to() per contig: {1e3*(t1-t0):.1f} ms ({sum(lengths)/(t1-t0)/1e9:.1f} GB/s)")
del keep
buf = torch.empty(1 << 26, dtype=torch.uint8, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for h in host:
    for a in range(0, h.numel(), 1 << 26):
        b = min(h.numel(), a + (1 << 26))
        buf[: b - a].copy_(h[a:b], non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"copy_ into a fixed 64 MiB device buffer: {1e3*(t1-t0):.1f} ms ({sum(lengths)/(t1-t0)/1e9:.1f} GB/s)")
torch.cuda.synchronize(); t0 = time.perf_counter()
sh.plane2.zero_(); sh.plane4.zero_(); sh.valid.zero_()
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"zeroing planes: {1e3*(t1-t0):.2f} ms")

# the pipeline (upload_and_scan): copy stream | pack | per-contig range scans, at several scan-group sizes
import merpcr_b200.engine as E
for grp in (1 << 25, 1 << 27, 1 << 40):
    E.STREAM_SCAN_BASES = grp
    shp = None
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        shp, hits_t, n = eng.upload_and_scan(lay, host, shard=shp)
        out = eng._hits_to_host(hits_t, n)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if it:
            print(f"upload_and_scan, scan groups >= {grp} bases, iter {it}: {1e3*(t1-t0):.1f} ms ({sum(lengths)/(t1-t0)/1e9:.1f} Gbp/s), hits {len(out)}", flush=True)
