"""Phase timings of the scanner on the bench workload (cfg3): scan_kernel alone with MPCR_DEBUG = 1 (stage 1 only:
rolling keys + filter), 2 (stage 1 + 2: slot gathers + tags, survivors dropped) and 0 (everything).

    python scripts/gpu/phases.py [--scale 1.0] [--reps 5]          (GPU box; honours $MPCR_B200_LIB)
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
import synth  # noqa: E402
from merpcr_b200 import MerPCR  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lengths, n_sts, sts = bench.workload(args.scale)
    with tempfile.NamedTemporaryFile("wb", suffix=".sts", delete=False) as f:
        f.write(synth.sts_lines(sts))
    eng = MerPCR(**bench.PARAMS, device=0)
    assert eng.load_sts_file(f.name)
    os.unlink(f.name)
    layout = eng.make_layout(lengths)
    expected, writes = bench.plan_writes(lengths, sts, 0)
    by_contig = {}
    for ci, off, b in writes:
        by_contig.setdefault(ci, []).append((off, b))
    seqs = []
    for ci, L in enumerate(lengths):
        t = synth.dna_torch(bench.contig_seed(0, ci), 0, L, dev)
        w = by_contig.get(ci)
        if w:
            idx = np.concatenate([np.arange(off, off + len(b), dtype=np.int64) for off, b in w])
            val = np.concatenate([b for _, b in w])
            t[torch.from_numpy(idx).to(dev)] = torch.from_numpy(val).to(dev)
        seqs.append(t)
    sh = eng.upload(layout, seqs)
    torch.cuda.synchronize()
    lib, ctx = eng._be.lib, eng._ctx
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    contigs = layout["contigs"]
    bp = float(sum(lengths))
    for dbg in (1, 2, 0, 1, 2, 0):
        os.environ["MPCR_DEBUG"] = str(dbg)
        ms = []
        for _ in range(args.reps):
            eng._be.check(lib.mpcr_scan(ctx, contigs.ctypes.data, len(contigs), sh.plane2.data_ptr(), sh.plane4.data_ptr(),
                                        sh.valid.data_ptr(), sh.origin, sh.alloc, sh.begin, sh.end, 0, 0, count.data_ptr(),
                                        eng._stream()))
            k = int(count.item())
            ms.append(float(lib.mpcr_last_scan_ms(ctx)))
        print(f"MPCR_DEBUG={dbg}: scan_kernel {np.median(ms):.3f} ms (min {min(ms):.3f}), counter {k} "
              f"({k / bp:.5f} per bp)", flush=True)


if __name__ == "__main__":
    main()
