"""BASELINE.json configs at (or near) full size on one B200 (test / probe infrastructure): build the synthetic workload
in HBM, scan it through the public engine, and check the size-independent properties -- every planted amplicon found,
output sorted, rescan idempotent, the head of contig 0 bit-exact against the oracle.  Used by
tests/test_gpu_fullsize.py and scripts/gpu/configs_probe.py."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402
from merpcr_b200 import MerPCR, _capi  # noqa: E402

DEGEN = {ord("A"): b"RMWN", ord("C"): b"YMSN", ord("G"): b"RKSN", ord("T"): b"YKWN"}   # codes that contain the base


def build_genome(seed, lengths, dev, n_runs_frac=0.0, iupac_frac=0.0):
    contigs = []
    rng = synth.Rng(seed + 17)
    for ci, L in enumerate(lengths):
        t = synth.dna_torch(seed * 1000003 + ci, 0, L, dev)
        if n_runs_frac > 0:
            covered, target = 0, int(L * n_runs_frac)
            while covered < target:
                ln = int(10 ** (1 + 5 * float(rng.u64(1)[0] >> np.uint64(11)) / float(1 << 53)))   # log-uniform 10 .. 1e6
                ln = min(ln, target - covered + 10, L // 4)
                a = rng.randint(0, L - ln)
                t[a: a + ln] = ord("N")
                covered += ln
        if iupac_frac > 0:
            k = int(L * iupac_frac)
            pos = torch.from_numpy(rng.ints(0, L - 1, k)).to(dev)
            letters = torch.tensor(list(b"RYKMSWBDHVN"), dtype=torch.uint8, device=dev)
            t[pos] = letters[torch.from_numpy(rng.ints(0, 10, k)).to(dev)]
        contigs.append(t)
    return contigs


def plant(contigs, lengths, sts, margin, seed, sub_mode):
    expected, writes = synth.plant_amplicons(seed, list(lengths), sts, margin, sub_mode=sub_mode)
    by = {}
    for ci, off, b in writes:
        by.setdefault(ci, []).append((off, b))
    dev = contigs[0].device
    for ci, w in by.items():
        idx = np.concatenate([np.arange(off, off + len(b), dtype=np.int64) for off, b in w])
        val = np.concatenate([b for _, b in w])
        contigs[ci][torch.from_numpy(idx).to(dev)] = torch.from_numpy(val).to(dev)
    return expected


def degenerate_primers(sts, seed, frac=0.2):
    """1-3 degenerate codes in `frac` of the STS, each containing the original base, primer1's first 11-mer kept clean."""
    rng = synth.Rng(seed)
    n = len(sts["l1"])
    pick = rng.ints(0, 99, n) < int(frac * 100)
    for i in np.flatnonzero(pick).tolist():
        for _ in range(rng.randint(1, 3)):
            which = "p1" if rng.chance(0.5) else "p2"
            ln = int(sts["l1" if which == "p1" else "l2"][i])
            j = rng.randint(11, ln - 2)   # both primers keep their first 11-mer clean (each is hashed for one strand)
            base = int(sts[which][i, j])
            if base in DEGEN:
                sts[which][i, j] = DEGEN[base][rng.randint(0, 3)]


def run(name, lengths, n_sts, params, sub_mode, ranged, seed, decorate, dev, oracle_bp=1_000_000, verbose=True):
    t0 = time.time()
    sts = synth.make_sts_set(seed + 1, n_sts, 18, 25, 100, 1000)
    contigs = build_genome(seed, lengths, dev, 0.05 if decorate else 0.0, 1e-4 if decorate else 0.0)
    expected = plant(contigs, lengths, sts, params["margin"], seed + 2, sub_mode)
    if decorate:
        degenerate_primers(sts, seed + 3)
    sts_text = synth.sts_lines(sts, ranged=ranged)
    with tempfile.NamedTemporaryFile("wb", suffix=".sts", delete=False) as f:
        f.write(sts_text)
        path = f.name
    t_gen = time.time() - t0
    eng = MerPCR(**params, device=dev.index)
    t0 = time.time()
    assert eng.load_sts_file(path)
    t_sts = time.time() - t0
    os.unlink(path)
    layout = eng.make_layout(lengths)
    shard = eng.upload(layout, contigs)
    torch.cuda.synchronize()
    times, n = [], 0
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        hits_t, n = eng.scan_device(layout, shard, sort="--timing-only" not in sys.argv)
        torch.cuda.synchronize(); times.append(time.time() - t0)
    scan_ms = float(eng.last_scan_ms)      # scan kernels of every table (exact dense searches use two)
    ver_ms = sum(float(eng._be.lib.mpcr_last_verify_ms(c)) for c in eng._all_ctxs() if c)
    if "--timing-only" in sys.argv:   # used with MPCR_DEBUG phase switches, where the "hits" are only a counter
        print(json.dumps(dict(config=name, count=int(n), scan_kernel_ms=round(scan_ms, 3), verify_kernel_ms=round(ver_ms, 3),
                              debug=os.environ.get("MPCR_DEBUG", "0"))), flush=True)
        eng.close()
        return
    hits = hits_t[: n * _capi.HIT_DTYPE.itemsize].cpu().numpy().view(_capi.HIT_DTYPE).copy()
    hits2_t, n2 = eng.scan_device(layout, shard)
    hits2 = hits2_t[: n2 * _capi.HIT_DTYPE.itemsize].cpu().numpy().view(_capi.HIT_DTYPE)
    idem = n == n2 and bool(np.array_equal(hits, hits2))
    found = set(zip(hits["contig"].tolist(), hits["pos1"].tolist(), hits["pos2"].tolist()))
    planted_ok = all((c, a, b) in found for c, a, b, _, _ in expected)
    key = np.stack([hits["contig"], hits["pos1"]], axis=1).astype(np.int64)
    sorted_ok = bool(np.all((key[1:, 0] > key[:-1, 0]) | ((key[1:, 0] == key[:-1, 0]) & (key[1:, 1] >= key[:-1, 1]))))
    # oracle on the head of contig 0 (single thread == reference -T 1)
    from oracle.oracle import Oracle
    sub = min(oracle_bp, lengths[0])
    head = contigs[0][:sub].cpu().numpy()
    o = Oracle(**params)
    assert o.load_sts_text(sts_text)
    t0 = time.time()
    oh = o.search_hits(head.tobytes(), threads=1)
    t_or = time.time() - t0
    safe = sub - (int(sts["size"].max()) + 40 + params["margin"] + 64)
    oh = oh[oh[:, 1] < safe]
    g = hits[(hits["contig"] == 0) & (hits["pos2"] < safe)]
    recs = o.records()
    line = np.array([r["offset"] for r in recs], dtype=np.int64)
    minus = np.array([r["direct"] == "-" for r in recs], dtype=np.int64)
    gl = np.array([eng.sts_records[i].offset for i in eng._rec_to_idx[g["rec"]].tolist()], dtype=np.int64)
    parity = (len(oh) == len(g) and bool(np.array_equal(oh[:, 0], g["pos1"]) and np.array_equal(oh[:, 1], g["pos2"]) and
                                         np.array_equal(line[oh[:, 2]], gl) and np.array_equal(minus[oh[:, 2]], g["rec"] & 1)))
    bp = int(sum(lengths))
    out = dict(config=name, bp=bp, n_sts=n_sts, params=params, hits=int(n), planted=len(expected), planted_found=planted_ok,
               sorted=sorted_ok, idempotent=idem, oracle_head_bp=sub, oracle_head_hits=int(len(oh)), oracle_head_bit_exact=parity,
               oracle_head_seconds=round(t_or, 2), scan_kernel_ms=round(scan_ms, 3), verify_kernel_ms=round(ver_ms, 3),
               step_ms=round(1e3 * min(times), 3), gbp_per_s=round(bp / min(times) / 1e9, 2), load_sts_s=round(t_sts, 2),
               gen_s=round(t_gen, 1))
    if verbose:
        print(json.dumps(out), flush=True)
    eng.close()
    del contigs, shard
    torch.cuda.empty_cache()
    return out


def configs(scale: float = 1.0):
    """name -> run() arguments for BASELINE.json configs 2..5 at the given scale."""
    g38 = [max(20000, int(L * scale)) for L in synth.GRCH38_LENGTHS]
    return {
        "cfg2": ("cfg2: 100 Mbp single contig x 10k STS, -W 11 -N 0 -M 50", [int(100_000_000 * scale)],
                 max(100, int(10000 * scale)), dict(wordsize=11, margin=50, mismatches=0), "none", False, 1001, False),
        "cfg3": ("cfg3: 3.1 Gbp x 100k STS, -W 11 -N 1 -X 1 -M 50", g38, max(100, int(100000 * scale)),
                 dict(wordsize=11, margin=50, mismatches=1, three_prime_match=1), "cfg3", False, 1003, False),
        "cfg4": ("cfg4: 3.1 Gbp with 5% N-runs + IUPAC, 100k STS (20% degenerate), -I 1 -N 2 -W 11 -M 50 -X 1", g38,
                 max(100, int(100000 * scale)),
                 dict(wordsize=11, margin=50, mismatches=2, three_prime_match=1, iupac_mode=1), "cfg3", False, 1004, True),
        "cfg5": ("cfg5: 3.1 Gbp x 1M STS, ranged sizes, -W 8 -M 500 -N 0", g38, max(1000, int(1000000 * scale)),
                 dict(wordsize=8, margin=500, mismatches=0), "none", True, 1005, False),
    }


