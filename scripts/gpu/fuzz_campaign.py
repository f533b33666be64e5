"""Randomised GPU-vs-oracle parity campaign beyond the committed goldens (the oracle is pinned to the reference by
tests/test_oracle_golden.py; this script never reads /root/reference).

  part 1: fresh seeds of tests/fuzzcases.py (files in, text out) -- `oracle.run_files` vs the product CLI path;
  part 2: medium random workloads (1-3 Mbp, decorations, random W/N/X/M/I, planted + mutated amplicons) -- ordered hit
          arrays vs the oracle, once through one table and once sharded 3 ways and merged.

    python scripts/gpu/fuzz_campaign.py [--small 2000] [--medium 40] [--seed0 100000] [--out gpurun_out/fuzz_campaign.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import fuzzcases  # noqa: E402
import parity  # noqa: E402
import synth  # noqa: E402
from merpcr_b200 import FASTARecord, MerPCR  # noqa: E402
from merpcr_b200 import multi  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def small_case(seed, tmp):
    c = fuzzcases.make_case(seed)
    sp, fp, op = (os.path.join(tmp, x) for x in ("in.sts", "in.fa", "out.txt"))
    with open(sp, "w", newline="") as f:
        f.write(c["sts_text"])
    with open(fp, "w", newline="") as f:
        f.write(c["fasta_text"])
    p = c["params"]
    o = Oracle(**p)
    try:
        st, n_o, text_o = o.run_files(sp, fp)
    except Exception as e:  # noqa: BLE001
        st, n_o, text_o = -1, 0, f"oracle raised {type(e).__name__}"
    eng = MerPCR(**p)
    try:
        if not eng.load_sts_file(sp):
            got = (1, 0, "")
        else:
            try:
                recs = eng.load_fasta_file(fp)
            except IndexError:
                return True, c                  # bare '>' header: the reference raises too (golden-tested)
            if not recs:
                got = (1, 0, "")
            else:
                n = eng.search(recs, op)
                got = (0, n, open(op, newline="").read())
    finally:
        eng.close()
    want = (st, n_o, text_o) if st == 0 else (1, 0, "")
    return got == want, c


def medium_case(seed):
    r = synth.Rng(seed)
    W = r.choice([6, 8, 9, 10, 11, 11, 12, 13, 14, 16])
    params = dict(wordsize=W, margin=r.choice([0, 20, 50, 200]), mismatches=r.choice([0, 1, 1, 2, 3]),
                  three_prime_match=r.choice([0, 1, 1, 3]), iupac_mode=r.choice([0, 0, 1]))
    lengths = [r.randint(200_000, 1_500_000) for _ in range(r.randint(1, 3))] + [r.randint(0, 40)]
    n_sts = r.randint(100, 600 if W <= 8 else 3000)
    contigs = [r.dna(n) for n in lengths]
    sts = synth.make_sts_set(seed + 1, n_sts, max(W, 12), 28, 60, 900)
    big = [c for c in contigs if len(c) > 100000]
    synth.plant_amplicons(seed + 2, big, sts, params["margin"], sub_mode=r.choice(["none", "cfg3"]),
                          plant_count=min(n_sts, sum(len(c) for c in big) // 2500))
    if r.chance(0.6):
        for c in big:
            for _ in range(30):
                a = r.randint(0, len(c) - 1)
                c[a: a + r.choice([1, 3, 10, 200, 5000])] = ord("N")
            pos = r.ints(0, len(c) - 1, len(c) // 4000)
            c[pos] = np.frombuffer(b"RYKMSWBDHVNX", dtype=np.uint8)[r.ints(0, 11, len(pos))]
        for i in range(0, n_sts, 4):
            if int(sts["l1"][i]) > W + 3:
                j = r.randint(W + 1, int(sts["l1"][i]) - 2)
                sts["p1"][i, j] = ord(r.choice("RYMKSWBDHVN"))
            j = r.randint(1, int(sts["l2"][i]) - 2)
            sts["p2"][i, j] = ord(r.choice("RYMKSWBDHVN"))
    sts_text = synth.sts_lines(sts, ranged=r.chance(0.3))
    want = parity.oracle_hits(params, sts_text.decode(), [c.tobytes() for c in contigs])
    with tempfile.NamedTemporaryFile("wb", suffix=".sts", delete=False) as f:
        f.write(sts_text)
    try:
        def recs():
            out = []
            for i, c in enumerate(contigs):
                x = FASTARecord(f">c{i}", c)
                x._from_loader = True
                out.append(x)
            return out
        eng = MerPCR(**params)
        assert eng.load_sts_file(f.name)
        got = parity.engine_hits(eng, recs())
        ok = got.shape == want.shape and np.array_equal(got, want)
        eng.close()
        # sharded 3 ways, merged by the order key
        parts = []
        for k in range(3):
            e = MerPCR(**params, shard=(k, 3))
            assert e.load_sts_file(f.name)
            parts.append(e.search_hits(recs()))
            e.close()
        merged = multi.merge_hits(parts)
        ok_sh = len(merged) == len(want) and np.array_equal(merged["pos1"], want[:, 1]) and \
            np.array_equal(merged["pos2"], want[:, 2]) and np.array_equal(merged["contig"], want[:, 0])
    finally:
        os.unlink(f.name)
    return ok, ok_sh, params, len(want), sum(lengths), n_sts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", type=int, default=2000)
    ap.add_argument("--medium", type=int, default=40)
    ap.add_argument("--seed0", type=int, default=100000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fuzz_campaign.json"))
    ap.add_argument("--emul", action="store_true", help="dry run of this script on a CPU box through tests/host_emul")
    a = ap.parse_args()
    if a.emul:
        import emul
        emul.inject()
    t0 = time.time()
    bad_small, bad_medium, hits_total = [], [], 0
    with tempfile.TemporaryDirectory() as tmp:
        for s in range(a.seed0, a.seed0 + a.small):
            ok, c = small_case(s, tmp)
            if not ok:
                bad_small.append(dict(seed=s, params=c["params"]))
                print("SMALL MISMATCH", s, c["params"], flush=True)
    t1 = time.time()
    print(f"part 1: {a.small} cases, {len(bad_small)} mismatches, {t1 - t0:.0f}s", flush=True)
    med = []
    for s in range(a.seed0, a.seed0 + a.medium):
        ok, ok_sh, params, n, bp, n_sts = medium_case(7 * s + 1)
        hits_total += n
        med.append(dict(seed=7 * s + 1, params=params, hits=n, bp=bp, n_sts=n_sts, ok=bool(ok), ok_sharded=bool(ok_sh)))
        if not (ok and ok_sh):
            bad_medium.append(med[-1])
            print("MEDIUM MISMATCH", med[-1], flush=True)
    t2 = time.time()
    print(f"part 2: {a.medium} workloads, {hits_total} hits compared, {len(bad_medium)} mismatches, {t2 - t1:.0f}s", flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(dict(small=a.small, medium=a.medium, seed0=a.seed0, small_mismatches=bad_small, medium_mismatches=bad_medium,
                   medium_cases=med, hits_compared=hits_total, seconds=t2 - t0), open(a.out, "w"), indent=1)
    return 1 if (bad_small or bad_medium) else 0


if __name__ == "__main__":
    sys.exit(main())
