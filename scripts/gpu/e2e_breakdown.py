"""End-to-end step on the bench workload (planted cfg3): where do the milliseconds above the PCIe copy go?"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import numpy as np, torch
import bench, synth
import merpcr_b200.engine as E
from merpcr_b200 import MerPCR
dev = torch.device('cuda', 0)
lengths, n_sts, sts = bench.workload(1.0)
with tempfile.NamedTemporaryFile('wb', suffix='.sts', delete=False) as f:
    f.write(synth.sts_lines(sts))
eng = MerPCR(**bench.PARAMS, device=0)
assert eng.load_sts_file(f.name)
lay = eng.make_layout(lengths)
expected, writes = bench.plan_writes(lengths, sts, 0)
by = {}
for ci, off, b in writes:
    by.setdefault(ci, []).append((off, b))
host = []
for ci, L in enumerate(lengths):
    t = synth.dna_torch(bench.contig_seed(0, ci), 0, L, dev)
    w = by.get(ci)
    if w:
        idx = np.concatenate([np.arange(off, off + len(b), dtype=np.int64) for off, b in w])
        t[torch.from_numpy(idx).to(dev)] = torch.from_numpy(np.concatenate([b for _, b in w])).to(dev)
    h = torch.empty(L, dtype=torch.uint8).pin_memory(); h.copy_(t); host.append(h); del t
torch.cuda.synchronize()
for grp in (1 << 27, 1 << 40, 1 << 26):
    E.STREAM_SCAN_BASES = grp
    sh = None
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sh, hits_t, n = eng.upload_and_scan(lay, host, shard=sh)
        t1 = time.perf_counter()
        out = eng._hits_to_host(hits_t, n)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if it:
            print(f"groups >= {grp}: upload_and_scan {1e3*(t1-t0):.2f} ms + hits to host {1e3*(t2-t1):.2f} ms = {1e3*(t2-t0):.2f} ms, hits {len(out)}", flush=True)
buf = torch.empty(1 << 26, dtype=torch.uint8, device=dev)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for h in host:
        for a in range(0, h.numel(), 1 << 26):
            b = min(h.numel(), a + (1 << 26))
            buf[: b - a].copy_(h[a:b], non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"copy only: {1e3*(t1-t0):.2f} ms")
