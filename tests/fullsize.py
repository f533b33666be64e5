"""BASELINE.json configs at (or near) full size on one B200 (test / probe infrastructure): build the synthetic workload
in HBM, scan it through the public engine, and check it against the oracle -- the COMPLETE ordered hit list of every
contig (oracle single-threaded per contig == reference -T 1, contigs spread over the host cores), or slices spread over
the contigs for the candidate-heavy config -- plus the size-independent properties: every planted amplicon found, output
sorted, rescan idempotent.  Used by tests/test_gpu_fullsize.py and scripts/gpu/configs_probe.py."""
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


def gpu_rows(eng, hits):
    """(contig, pos1, pos2, source line of the STS, strand) rows of a hit array, in its order."""
    line_nos = np.asarray(eng._sts_lines.line_nos, dtype=np.int64)
    return np.stack([hits["contig"].astype(np.int64), hits["pos1"].astype(np.int64), hits["pos2"].astype(np.int64),
                     line_nos[hits["rec"] >> 1], (hits["rec"] & 1).astype(np.int64)], axis=1)


def oracle_rows(params, sts_text, jobs, fetch, workers=None):
    """Oracle hit rows for slices of contigs.  jobs = [(contig, start, stop, cut)]: the bases [start, stop) of the
    contig are searched as one sequence by ONE thread (== the reference with -T 1; never its threaded path)
    and the hits with slice-local pos2 < cut are kept (cut = stop - start when the slice ends where the contig ends,
    otherwise far enough from the slice's end that its artificial end clamp cannot matter).  fetch(contig, start, stop)
    returns the bases as a contiguous uint8 array.  Jobs run concurrently on `workers` host threads (the C oracle
    releases the GIL; the engine is read-only while searching).  Returns a list of row arrays, one per job."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import Oracle
    o = Oracle(**params)
    assert o.load_sts_text(sts_text)
    line, minus = o.record_lines()

    def one(job):
        ci, a, b, cut = job
        h = o.search_hits_array(fetch(ci, a, b), threads=1)
        h = h[h[:, 1] < cut]
        return np.stack([np.full(len(h), ci, dtype=np.int64), h[:, 0] + a, h[:, 1] + a, line[h[:, 2]], minus[h[:, 2]]], axis=1)

    workers = workers or max(1, min(len(jobs), os.cpu_count() or 1))
    order = sorted(range(len(jobs)), key=lambda i: jobs[i][1] - jobs[i][2])     # longest first
    out = [None] * len(jobs)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        for i, rows in zip(order, ex.map(one, [jobs[i] for i in order])):
            out[i] = rows
    return out


def slice_jobs(lengths, n_slices, size, safe):
    """n_slices slices of `size` bases spread over the contigs, the first at the very start of contig 0 and the last
    ending exactly where the last contig ends (a genuine end clamp); the others are cut `safe` bases short."""
    jobs = []
    nc = len(lengths)
    for k in range(n_slices):
        ci = (k * nc) // n_slices if k + 1 < n_slices else nc - 1
        L = lengths[ci]
        sz = min(size, L)
        if k == 0:
            a = 0
        elif k + 1 == n_slices:
            a = L - sz
        else:
            a = ((k * 2654435761) % max(1, L - sz)) // 64 * 64
        b = a + sz
        jobs.append((ci, a, b, sz if b == L else sz - safe))
    return jobs


def compare_with_oracle(eng, hits, params, sts_text, contigs, lengths, safe, mode):
    """mode = "whole" (every contig, the complete ordered list) or ("slices", n, size).  Returns a dict of facts."""
    def fetch(ci, a, b):
        return contigs[ci][a:b].cpu().numpy()

    rows = gpu_rows(eng, hits)
    t0 = time.time()
    if mode == "whole":
        jobs = [(ci, 0, L, L) for ci, L in enumerate(lengths)]
        want = oracle_rows(params, sts_text, jobs, fetch)
        want = np.concatenate(want) if want else np.zeros((0, 5), dtype=np.int64)
        ok = want.shape == rows.shape and bool(np.array_equal(want, rows))
        first_diff = None
        if not ok:
            m = min(len(want), len(rows))
            d = np.flatnonzero((want[:m] != rows[:m]).any(axis=1))
            first_diff = dict(index=int(d[0]) if d.size else m, want=want[d[0]].tolist() if d.size else None,
                              got=rows[d[0]].tolist() if d.size else None, n_want=len(want), n_got=len(rows))
        return dict(oracle_mode="whole genome, every contig, complete ordered hit list", oracle_bp=int(sum(lengths)),
                    oracle_hits=int(len(want)), oracle_bit_exact=ok, oracle_first_diff=first_diff,
                    oracle_seconds=round(time.time() - t0, 2))
    _, n_slices, size = mode
    jobs = slice_jobs(lengths, n_slices, size, safe)
    want = oracle_rows(params, sts_text, jobs, fetch)
    ok, total, bad = True, 0, None
    for (ci, a, b, cut), w in zip(jobs, want):
        g = rows[(rows[:, 0] == ci) & (rows[:, 1] >= a) & (rows[:, 2] < a + cut)]
        total += len(w)
        if w.shape != g.shape or not np.array_equal(w, g):
            ok = False
            bad = bad or dict(job=[ci, a, b, cut], n_want=len(w), n_got=len(g))
    return dict(oracle_mode=f"{n_slices} slices of {size} bp spread over the contigs (both genome ends included)",
                oracle_bp=int(sum(b - a for _, a, b, _ in jobs)), oracle_hits=int(total), oracle_bit_exact=ok,
                oracle_first_diff=bad, oracle_seconds=round(time.time() - t0, 2))


def run(name, lengths, n_sts, params, sub_mode, ranged, seed, decorate, dev, oracle="whole", verbose=True):
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
    # the oracle (single thread per contig == reference -T 1) against the complete result
    safe = int(sts["size"].max()) + 40 + params["margin"] + 64
    facts = compare_with_oracle(eng, hits, params, sts_text, contigs, lengths, safe, oracle)
    bp = int(sum(lengths))
    out = dict(config=name, bp=bp, n_sts=n_sts, params=params, hits=int(n), planted=len(expected), planted_found=planted_ok,
               sorted=sorted_ok, idempotent=idem, **facts, scan_kernel_ms=round(scan_ms, 3), verify_kernel_ms=round(ver_ms, 3),
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
        # not a BASELINE config: config 5 with one mismatch allowed -- the candidate-heavy search that can use neither the
        # seed extension nor position sampling and runs on block tables (mpcr_ctx_set_seed_blocks)
        "cfg5n1": ("cfg5n1: 3.1 Gbp x 1M STS, ranged sizes, -W 8 -M 500 -N 1", g38, max(1000, int(1000000 * scale)),
                   dict(wordsize=8, margin=500, mismatches=1), "cfg3", True, 1006, False),
    }


