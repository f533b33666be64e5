"""File -> hits wall clock with the device-side FASTA ingest vs the host parser (1 Gbp FASTA, 60-column lines)."""
import os, sys, time
sys.path[:0] = ['.', 'tests']
import numpy as np, torch
import synth
from merpcr_b200 import MerPCR
from merpcr_b200.fasta import FASTALoader
n = 1_000_000_020
rng = synth.Rng(5)
t0 = time.time()
with open('/tmp/big.fa', 'wb') as f:
    for c in range(4):
        seq = synth.dna_chunked(77 + c, n // 4)
        lines = seq[: len(seq) // 60 * 60].reshape(-1, 60)
        buf = np.empty((lines.shape[0], 61), dtype=np.uint8); buf[:, :60] = lines; buf[:, 60] = 10
        f.write(b'>chr%d synthetic\n' % c); f.write(buf.tobytes())
print(f"wrote {os.path.getsize('/tmp/big.fa')/1e9:.2f} GB in {time.time()-t0:.1f}s", flush=True)
sts = synth.make_sts_set(8, 100000)
open('/tmp/x.sts', 'wb').write(synth.sts_lines(sts))
eng = MerPCR(mismatches=1, device=0)
t0 = time.time(); eng.load_sts_file('/tmp/x.sts'); print(f"load_sts_file (100k lines): {time.time()-t0:.2f}s", flush=True)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    recs = eng.load_fasta_file('/tmp/big.fa')
    torch.cuda.synchronize(); t1 = time.time()
    nh = eng.search(recs, '/tmp/out.txt')
    torch.cuda.synchronize(); t2 = time.time()
    print(f"device ingest: load_fasta_file {t1-t0:.2f}s ({sum(len(r) for r in recs)/(t1-t0)/1e9:.2f} Gbp/s), search {t2-t1:.3f}s, hits {nh}", flush=True)
t0 = time.time(); recs_h = FASTALoader.load_file('/tmp/big.fa'); t1 = time.time()
print(f"host parser: load_file {t1-t0:.2f}s ({sum(len(r) for r in recs_h)/(t1-t0)/1e9:.3f} Gbp/s)", flush=True)
t1 = time.time(); nh = eng.search(recs_h, '/tmp/out2.txt'); t2 = time.time()
print(f"search from host records {t2-t1:.3f}s hits {nh}; outputs equal: {open('/tmp/out.txt').read() == open('/tmp/out2.txt').read()}")
