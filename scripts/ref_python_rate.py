"""Context number for profiles/: the rate of the UNMODIFIED Python reference (/root/reference, build container only --
it cannot travel to the GPU box) on a cfg3-shaped sample: 100 000 planted STS (-W 11 -N 1 -X 1 -M 50) against the first
few Mbp of the synthetic chr1, with -T 1 and -T <cores>.  Writes profiles/r2_python_reference_rate.json.

    python scripts/ref_python_rate.py [Mbp]
"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), "/root/reference/src"]
import numpy as np
import synth
import bench
from merpcr.core.engine import MerPCR as RefMerPCR   # the reference itself

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
lengths, n_sts, sts = bench.workload(1.0)
sts_text = synth.sts_lines(sts)
nb = int(mbp * 1e6)
arr = synth.dna_chunked(bench.contig_seed(0, 0), nb)
_, writes = bench.plan_writes(lengths, sts, 0)
bench.apply_writes_numpy(arr, 0, writes)
cores = os.cpu_count() or 1
out = dict(what="unmodified Python reference (merpcr.core.engine.MerPCR) on the first %d bp of the synthetic cfg3 chr1 x %d STS, "
                "-W 11 -N 1 -X 1 -M 50; search() only, FASTA/STS load excluded" % (nb, n_sts), cores=cores, runs=[])
with tempfile.TemporaryDirectory() as d:
    sp, fp = os.path.join(d, "s.sts"), os.path.join(d, "s.fa")
    open(sp, "wb").write(sts_text)
    with open(fp, "wb") as f:
        f.write(b">chr1 sample\n")
        body = arr[: nb // 60 * 60].reshape(-1, 60)
        f.write(b"\n".join(r.tobytes() for r in body) + b"\n" + arr[nb // 60 * 60:].tobytes() + b"\n")
    for threads in (1, cores):
        eng = RefMerPCR(wordsize=11, margin=50, mismatches=1, three_prime_match=1, threads=threads)
        t0 = time.time(); assert eng.load_sts_file(sp); t_sts = time.time() - t0
        t0 = time.time(); recs = eng.load_fasta_file(fp); t_fa = time.time() - t0
        t0 = time.time(); hits = eng.search(recs, os.path.join(d, "o.txt")); t_search = time.time() - t0
        out["runs"].append(dict(threads=threads, hits=int(hits), load_sts_s=round(t_sts, 2), load_fasta_s=round(t_fa, 2),
                                search_s=round(t_search, 2), gbp_per_s=round(nb / t_search / 1e9, 6)))
        print(out["runs"][-1], flush=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_python_reference_rate.json"), "w"), indent=1)
