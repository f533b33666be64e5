"""End-to-end step from pinned host ASCII (upload_and_scan): sweep of the hybrid wire scheduling -- the backlog factor
that decides whether a piece is packed on the host or goes up as ASCII, and the packer's thread count."""
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
    del t
torch.cuda.synchronize()
ncpu = len(os.sched_getaffinity(0))
shp = None
for threads in sorted({ncpu, max(1, ncpu * 3 // 4), max(1, ncpu // 2)}, reverse=True):
    for factor in (0.7, 1.0, 1.4, 2.0, 3.0):
        eng._pack_threads, eng.hybrid_backlog = threads, factor
        ts = []
        for it in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            shp, hits_t, n = eng.upload_and_scan(lay, host, shard=shp)
            out = eng._hits_to_host(hits_t, n, copy=False)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            if it:
                ts.append(1e3 * (t1 - t0))
        print(f"threads {threads:2d} backlog factor {factor:.1f}: {min(ts):.1f} ms (median {sorted(ts)[1]:.1f}), "
              f"h2d {eng.last_h2d_bytes / 1e9:.2f} GB, pack {1e3 * eng.last_timing['host_pack_s']:.1f} ms, hits {n}", flush=True)
