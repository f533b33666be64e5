"""Exercise every CUDA kernel of the path once on a representative workload, for the per-kernel ncu table
(scripts/gpu/prof_kernels.sh): FASTA text ingest -> pack -> table build -> sparse scan -> verify -> hit ordering, then a
candidate-heavy W=8 search over the first record (dense scanner, bucket walks, a larger hit list).

    python scripts/gpu/all_kernels.py [Mbp]          (default 512 Mbp in 4 records, 60-column FASTA lines, 100k STS planted)
"""
import os, sys, time
sys.path[:0] = ['.', 'tests']
import numpy as np, torch
import synth
from merpcr_b200 import MerPCR

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 512
per = mbp * 1_000_000 // 4 // 60 * 60 + 17            # not a multiple of 16: records start misaligned in HBM
sts = synth.make_sts_set(8, 100000)
expected, writes = synth.plant_amplicons(9, [per] * 4, sts, 50, sub_mode="cfg3")
with open('/tmp/k.fa', 'wb') as f:
    for c in range(4):
        seq = synth.dna_chunked(177 + c, per)
        for ci, off, b in writes:
            if ci == c:
                seq[off: off + len(b)] = b
        body, tail = seq[: per // 60 * 60].reshape(-1, 60), seq[per // 60 * 60:]
        buf = np.empty((body.shape[0], 61), dtype=np.uint8); buf[:, :60] = body; buf[:, 60] = 10
        f.write(b'>chr%d synthetic\n' % c); f.write(buf.tobytes()); f.write(tail.tobytes() + b'\n')
open('/tmp/k.sts', 'wb').write(synth.sts_lines(sts))
for label, kw in (("sparse W=11 N=1", dict(wordsize=11, mismatches=1)),
                  ("dense W=8 N=1 M=500", dict(wordsize=8, mismatches=1, margin=500))):
    eng = MerPCR(device=0, **kw)
    assert eng.load_sts_file('/tmp/k.sts')
    torch.cuda.synchronize(); t0 = time.time()
    recs = eng.load_fasta_file('/tmp/k.fa')
    if label.startswith("dense"):
        recs = recs[:1]
    nh = eng.search(recs, '/tmp/k.out')
    torch.cuda.synchronize()
    print(f"{label}: {sum(len(r) for r in recs)/1e6:.0f} Mbp, hits {nh} (planted findable: "
          f"{sum(1 for e in expected if e[0] < len(recs))}), {time.time()-t0:.2f}s, scan {eng.last_scan_ms:.3f} ms", flush=True)
