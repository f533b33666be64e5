"""Candidate-heavy searches that allow mismatches: the dense scanner (bucket walks at every position) against the block
tables (mpcr_ctx_set_seed_blocks: N + 1 sparse passes over split keys).  Same genome, same STS, same hits.

    python scripts/gpu/blocks_probe.py [Mbp] [n_sts]      (default 256 Mbp, 100k STS, -W 8 -N 1 -M 500 and -M 50)
"""
import os, sys, time
sys.path[:0] = ['.', 'tests']
import numpy as np, torch
import synth
from merpcr_b200 import MerPCR, FASTARecord

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_sts = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
L = mbp * 1_000_000
sts = synth.make_sts_set(8, n_sts)
expected, writes = synth.plant_amplicons(9, [L], sts, 50, sub_mode="cfg3")
seq = synth.dna_chunked(177, L)
for ci, off, b in writes:
    seq[off: off + len(b)] = b
open('/tmp/b.sts', 'wb').write(synth.sts_lines(sts))
recs = [FASTARecord(">chr1 synthetic", seq)]
results = {}
for margin in (50, 500):
    for flag in ("0", "1"):
        os.environ["MPCR_SEED_BLOCKS"] = flag
        eng = MerPCR(device=0, wordsize=8, mismatches=1, margin=margin)
        assert eng.load_sts_file('/tmp/b.sts')
        layout = eng.make_layout([L])
        shard = eng.upload(layout, [torch.from_numpy(seq).cuda()])
        torch.cuda.synchronize()
        eng.scan_device(layout, shard)
        ms = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            hits_t, n = eng.scan_device(layout, shard)
            torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
        h = hits_t[: n * 24].cpu().numpy().copy()
        results[(margin, flag)] = h
        print(f"-W 8 -N 1 -M {margin}, {n_sts} STS, {mbp} Mbp, block tables {'on ' if flag == '1' else 'off'}: "
              f"{len(eng._all_ctxs())} tables, step {min(ms):.3f} ms = {L / min(ms) / 1e6:.1f} Gbp/s, "
              f"scan kernels {eng.last_scan_ms:.3f} ms, hits {n}", flush=True)
        eng.close()
        del shard
    print("   hit lists identical:", np.array_equal(results[(margin, "0")], results[(margin, "1")]),
          " planted found:", sum(1 for e in expected), flush=True)
