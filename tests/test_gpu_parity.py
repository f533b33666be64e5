"""GPU tier (-m gpu): the CUDA path, called through the C ABI by the Python host, against
  * the reference's fixture golden line and the reference-generated fuzz goldens (bit-exact text),
  * the CPU oracle on seeded synthetic workloads shaped like BASELINE.json's configs (bit-exact hit lists in order),
  * planted-amplicon truth and size-independent properties at larger sizes.
Nothing here reads /root/reference; nothing uses the host emulation."""
import ctypes as C
import os

import numpy as np
import pytest

import goldens
import parity
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _real_backend():
    import torch
    from merpcr_b200 import _capi
    _capi._inject_backend_for_tests("", "cpu")          # drop any emulation a CPU-tier module left behind
    be = _capi.backend()
    # (MPCR_TEST_ALLOW_VARIANT: development runs of a tuning build through this tier; the driver never sets it)
    assert be.device_kind == "cuda" and (os.path.basename(_capi.LIB_PATH) == "libmerpcr_b200.so" or
                                         os.environ.get("MPCR_TEST_ALLOW_VARIANT") == "1")
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    yield


def _records(contigs, names=None):
    from merpcr_b200 import FASTARecord
    out = []
    for i, c in enumerate(contigs):
        r = FASTARecord(f">{names[i] if names else 'c%d' % i}", c)
        r._from_loader = True
        out.append(r)
    return out


def _write(tmp_path, name, data):
    p = tmp_path / name
    p.write_bytes(data)
    return str(p)


def test_fixture_golden_line(tmp_path):
    from merpcr_b200 import MerPCR
    eng = MerPCR()
    assert eng.load_sts_file(goldens.FIXTURE_STS)
    assert [r.hash_offset for r in eng.sts_records] == [0] * 6
    assert sorted(eng.sts_table) == sorted([3638181, 3526114, 2555953, 3737721, 2062650, 476488])
    out = tmp_path / "o.txt"
    assert eng.search(eng.load_fasta_file(goldens.FIXTURE_FA), str(out)) == 1
    assert out.read_text() == goldens.FIXTURE_LINE
    assert eng.gpu_launches >= 4


@pytest.mark.parametrize("flags,expect_hit", [
    (dict(mismatches=1), True), (dict(mismatches=2), True), (dict(wordsize=3), True),
    (dict(wordsize=8, mismatches=2, margin=500, three_prime_match=0), True), (dict(iupac_mode=1), True),
    (dict(margin=8), True), (dict(margin=7), False), (dict(wordsize=16), True), (dict(wordsize=12), True),
])
def test_fixture_flag_sweep(tmp_path, flags, expect_hit):
    from merpcr_b200 import MerPCR
    eng = MerPCR(**flags)
    assert eng.load_sts_file(goldens.FIXTURE_STS)
    out = tmp_path / "o.txt"
    eng.search(eng.load_fasta_file(goldens.FIXTURE_FA), str(out))
    assert out.read_text() == (goldens.FIXTURE_LINE if expect_hit else "")


def test_fuzz_goldens_bit_exact():
    from merpcr_b200 import MerPCR
    cases = goldens.fuzz_cases()
    for c in cases:
        parity.check_fuzz_case(c, MerPCR)


def test_pack_planes_match_numpy():
    """mpcr_pack_sequence against a numpy restatement of the plane layout (include/merpcr_b200.h)."""
    import torch
    from merpcr_b200 import MerPCR
    from merpcr_b200.alphabet import genome_lut
    rng = synth.Rng(5)
    n = 100000 + 37
    seq = rng.dna(n)
    letters = np.frombuffer(b"NRYKMSWBDHVXnacgtu", dtype=np.uint8)
    pos = rng.ints(0, n - 1, 4000)
    seq[pos] = letters[rng.ints(0, len(letters) - 1, 4000)]
    for iupac in (0, 1):
        eng = MerPCR(iupac_mode=iupac)
        eng.load_sts_file(goldens.FIXTURE_STS)
        lay = eng.make_layout([n])
        sh = eng.upload(lay, [seq])
        lut = genome_lut(iupac)[seq].astype(np.uint64)
        pad = (-n) % 64
        e = np.concatenate([lut, np.zeros(pad, dtype=np.uint64)])
        nib = (e & 15).reshape(-1, 16)
        p4 = (nib << (np.arange(16, dtype=np.uint64) * 4)).sum(axis=1).astype(np.uint64)
        c2 = ((e >> 4) & 3).reshape(-1, 32)
        p2 = (c2 << (np.arange(32, dtype=np.uint64) * 2)).sum(axis=1).astype(np.uint64)
        vb = ((e >> 6) & 1).reshape(-1, 64)
        v = (vb << np.arange(64, dtype=np.uint64)).sum(axis=1).astype(np.uint64)
        got4 = sh.plane4.cpu().numpy().view(np.uint64)[: len(p4)]
        got2 = sh.plane2.cpu().numpy().view(np.uint64)[: len(p2)]
        gotv = sh.valid.cpu().numpy().view(np.uint64)[: len(v)]
        assert np.array_equal(got4, p4) and np.array_equal(got2, p2) and np.array_equal(gotv, v)
        assert not sh.plane4.cpu().numpy().view(np.uint64)[len(p4):].any()     # padding stays zero / invalid


def test_pack_from_unaligned_device_source():
    """Sequences the device-side ingest leaves in HBM start at arbitrary byte offsets of one buffer: pack_kernel's
    aligned-chunk path (five 16-byte loads + funnel shifts) must give the same planes as the aligned path, for every
    misalignment and for lengths around the strip size."""
    import torch
    from merpcr_b200 import MerPCR
    rng = synth.Rng(17)
    big = rng.dna(70000)
    letters = np.frombuffer(b"NRYKMSWBDHVXnacgt", dtype=np.uint8)
    pos = rng.ints(0, len(big) - 1, 3000)
    big[pos] = letters[rng.ints(0, len(letters) - 1, 3000)]
    eng = MerPCR()
    eng.load_sts_file(goldens.FIXTURE_STS)
    dev = torch.from_numpy(big).to(eng._tdev)
    for n in (1, 15, 63, 64, 65, 127, 128, 1000, 65536 + 21):
        ref = None
        for mis in range(0, 17):
            src = dev[mis: mis + n]
            lay = eng.make_layout([n])
            sh = eng.upload(lay, [big[mis: mis + n]])            # host bytes -> aligned staging buffer
            want = (sh.plane2.cpu().numpy().copy(), sh.plane4.cpu().numpy().copy(), sh.valid.cpu().numpy().copy())
            sh2 = eng.upload(lay, [src])                          # device slice, misaligned by `mis` bytes
            got = (sh2.plane2.cpu().numpy(), sh2.plane4.cpu().numpy(), sh2.valid.cpu().numpy())
            assert all(np.array_equal(a, b) for a, b in zip(want, got)), (n, mis)


def test_device_table_matches_oracle(tmp_path):
    """Device-built records (hash offsets, hash values, reverse complements) vs the oracle's load (engine.py:253-281)."""
    from merpcr_b200 import MerPCR
    from oracle.oracle import Oracle
    rng = synth.Rng(11)
    lines = []
    for i in range(3000):
        a = rng.dna(rng.randint(11, 40)).tobytes().decode()
        b = rng.dna(rng.randint(11, 40)).tobytes().decode()
        if i % 7 == 0:
            j = rng.randint(0, len(a) - 1)
            a = a[:j] + rng.choice("NRYKMSWBDHVXU") + a[j + 1:]
        if i % 11 == 0:
            j = rng.randint(0, len(b) - 1)
            b = b[:j] + rng.choice("NRYKMSWBDHVXZ") + b[j + 1:]
        lines.append(f"ID{i}\t{a}\t{b}\t{rng.randint(30, 900)}\tal {i}\n")
    text = "".join(lines)
    path = _write(tmp_path, "t.sts", text.encode())
    for W in (8, 11, 16):
        o = Oracle(wordsize=W)
        assert o.load_sts_text(text)
        eng = MerPCR(wordsize=W)
        assert eng.load_sts_file(path)
        want = o.records()
        assert len(eng.sts_records) == len(want)
        for g, w in zip(eng.sts_records, want):
            assert (g.id, g.direct, g.hash_offset, g.primer1, g.primer2, g.pcr_size) == \
                (w["id"], w["direct"], w["hash_offset"], w["primer1"], w["primer2"], w["pcr_size"])
        table = eng.sts_table
        assert sorted(table) == sorted({w["hash"] for w in want})
        # the device's encoded reverse-complement primer of every '-' record decodes back to the host string
        from merpcr_b200.alphabet import IUPAC_MASK
        inv = {v: k for k, v in IUPAC_MASK.items()}
        lib, words, nw = eng._be.lib, (C.c_uint64 * 16)(), C.c_uint32(0)
        slots = np.flatnonzero(eng._rec_to_idx >= 0)
        for slot in slots[:400].tolist():
            rec = eng.sts_records[int(eng._rec_to_idx[slot])]
            eng._be.check(lib.mpcr_table_primer_words(eng._ctx, slot, 2, words, 16, C.byref(nw)))
            half = nw.value // 2
            dec = ""
            for i, ch in enumerate(rec.primer2):
                nib = (words[i // 16] >> (4 * (i % 16))) & 15
                aux = (words[half + i // 16] >> (4 * (i % 16))) & 3
                dec += inv.get(nib, "X" if aux == 2 else "?") if not (aux & 1) else "?"
            expect = "".join(c if (c in IUPAC_MASK or c == "X") else "?" for c in rec.primer2)
            assert dec == expect, (slot, rec.primer2, dec)


CASES = [
    # (name, contig lengths, n_sts, params, sub_mode, decorate, ranged)
    ("cfg2-like", [3_000_000], 1500, dict(wordsize=11, margin=50, mismatches=0), "none", False, False),
    ("cfg3-like", [1_300_000, 900_001, 450_000, 77, 11, 250_000], 2500,
     dict(wordsize=11, margin=50, mismatches=1, three_prime_match=1), "cfg3", False, False),
    ("cfg4-like", [1_500_000, 800_000], 2000,
     dict(wordsize=11, margin=50, mismatches=2, three_prime_match=1, iupac_mode=1), "cfg3", True, False),
    ("cfg5-like", [400_000, 150_000], 20000, dict(wordsize=8, margin=500, mismatches=0), "none", False, True),
    ("w16", [600_000], 800, dict(wordsize=16, margin=20, mismatches=3, three_prime_match=0), "cfg3", False, False),
    # -W 11 with a mismatch budget that has no compile-time instantiation (scan_kernel<.., 11, -1>: linear filter map, N at
    # run time), and the neighbouring word size (general direct table, multiplicative filter map)
    ("w11-n3", [900_000, 300_000], 1500, dict(wordsize=11, margin=40, mismatches=3, three_prime_match=1), "cfg3", False, False),
    ("w10-n1", [700_000, 200_000], 1200, dict(wordsize=10, margin=50, mismatches=1, three_prime_match=1), "cfg3", False, False),
    # several records per seed on average -> the bucket-parallel dense scanner
    ("dense-w8", [300_000, 120_000], 150000, dict(wordsize=8, margin=100, mismatches=0), "none", False, True),
    ("dense-w6-iupac", [250_000, 101_000], 20000,
     dict(wordsize=6, margin=30, mismatches=1, three_prime_match=2, iupac_mode=1), "cfg3", True, False),
]


def _make_workload(seed, lengths, n_sts, params, sub_mode, decorate, ranged):
    rng = synth.Rng(seed)
    contigs = [rng.dna(n) for n in lengths]
    sts = synth.make_sts_set(seed + 1, n_sts, 18, 25, 100, 1000)
    big = [c for c in contigs if len(c) > 100000]
    expected = synth.plant_amplicons(seed + 2, big, sts, params["margin"], sub_mode=sub_mode,
                                     plant_count=min(n_sts, sum(len(c) for c in big) // 2200))
    if decorate:
        # N runs (some abutting / inside amplicons), scattered IUPAC letters, degenerate primers
        for c in big:
            for _ in range(40):
                a = rng.randint(0, len(c) - 1)
                c[a: a + rng.choice([1, 3, 10, 200, 5000])] = ord("N")
            pos = rng.ints(0, len(c) - 1, len(c) // 5000)
            c[pos] = np.frombuffer(b"RYKMSWBDHVNX", dtype=np.uint8)[rng.ints(0, 11, len(pos))]
        for i in range(0, n_sts, 5):
            j = rng.randint(12, int(sts["l1"][i]) - 2)
            sts["p1"][i, j] = ord(rng.choice("RYMKSWBDHVN"))
            j = rng.randint(1, int(sts["l2"][i]) - 2)
            sts["p2"][i, j] = ord(rng.choice("RYMKSWBDHVN"))
    return contigs, synth.sts_lines(sts, ranged=ranged), expected


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_synthetic_configs_bit_exact_vs_oracle(tmp_path, case):
    from merpcr_b200 import MerPCR
    name, lengths, n_sts, params, sub_mode, decorate, ranged = case
    contigs, sts_text, expected = _make_workload(hash(name) % 1000 + 3000 if False else len(name) * 131 + n_sts,
                                                 lengths, n_sts, params, sub_mode, decorate, ranged)
    path = _write(tmp_path, "w.sts", sts_text)
    eng = MerPCR(**params)
    assert eng.load_sts_file(path)
    got = parity.engine_hits(eng, _records(contigs))
    want = parity.oracle_hits(params, sts_text.decode(), [c.tobytes() for c in contigs])
    assert got.shape == want.shape and np.array_equal(got, want), name
    if not decorate and name != "w16":     # (w16 substitutions may land inside the 16-base seed word)
        found = {(r[0], r[1], r[2]) for r in got.tolist()}
        big_index = [i for i, c in enumerate(contigs) if len(c) > 100000]
        assert all((big_index[ci], a, b) in found for ci, a, b, _, _ in expected)
    assert len(got) > 100 and len(expected) > 100


def test_long_primers_bit_exact_vs_oracle(tmp_path):
    """Primers of 30..120 bases: several nibble words per primer, the multi-word compare of the verifier (primers past
    32 bases have no hoisted view), tags and seeds at the front of long primers, amplicons barely longer than them."""
    from merpcr_b200 import MerPCR
    for seed, params in ((41, dict(wordsize=11, margin=50, mismatches=2, three_prime_match=1)),
                         (43, dict(wordsize=14, margin=10, mismatches=0, three_prime_match=3)),
                         (47, dict(wordsize=9, margin=30, mismatches=1, three_prime_match=0, iupac_mode=1))):
        rng = synth.Rng(seed)
        contigs = [rng.dna(n) for n in (700_000, 300_000, 150)]
        sts = synth.make_sts_set(seed + 1, 600, 30, 120, 260, 1000)
        expected = synth.plant_amplicons(seed + 2, contigs[:2], sts, params["margin"], sub_mode="cfg3", plant_count=300)
        sts["p1"][::9, 40] = ord("N")            # a degenerate letter deep inside some long primers
        sts["p2"][::7, 3] = ord("R")
        text = synth.sts_lines(sts)
        eng = MerPCR(**params)
        assert eng.load_sts_file(_write(tmp_path, f"s{seed}.sts", text))
        got = parity.engine_hits(eng, _records(contigs))
        want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
        assert got.shape == want.shape and np.array_equal(got, want), params
        assert len(want) > 50, (len(want), len(expected))
        eng.close()


def test_hit_buffer_regrows_and_sort_is_total(tmp_path):
    """More hits than the initial 65 536-entry buffer: nothing is truncated, order key is exact
    (repeat-rich sequence, duplicate STS lines, multi-delta hits)."""
    from merpcr_b200 import MerPCR
    unit = "ACGGTCATTGCAGT" + "TTGACCGGTATCAG" + "CATGCATGAACC"      # 40-mer tandem repeat
    seq = np.frombuffer((unit * 9000).encode(), dtype=np.uint8).copy()
    sts_text = "".join(f"R{i}\tACGGTCATTGCAGT\tTTGACCGGTATCAG\t{68 + (i % 3)}\trep\n" for i in range(4)).encode()
    path = _write(tmp_path, "r.sts", sts_text)
    params = dict(wordsize=8, margin=45, mismatches=0)
    eng = MerPCR(**params)
    assert eng.load_sts_file(path)
    got = parity.engine_hits(eng, _records([seq]))
    want = parity.oracle_hits(params, sts_text.decode(), [seq.tobytes()])
    assert len(want) > 65536 and np.array_equal(got, want)


def test_long_runs_of_equal_pos1_keep_discovery_order(tmp_path):
    """Hundreds of identical STS lines (and a few with other sizes) meeting the same positions of a tandem repeat: runs
    of equal (contig, pos1) far longer than 32 hits, which `order_ties` settles with its heap sort -- the order inside
    a run is the reference's discovery order (file order of the lines, then the probing order of the offsets)."""
    from merpcr_b200 import MerPCR
    unit = "ACGGTCATTGCAGT" + "TTGACCGGTATCAG" + "CATGCATGAACC"      # 40-mer tandem repeat
    seq = np.frombuffer(("G" * 500 + unit * 60 + "C" * 500).encode(), dtype=np.uint8).copy()
    lines = [f"D{i}\tACGGTCATTGCAGT\tTTGACCGGTATCAG\t{68 + 40 * (i % 3)}\tdup {i}\n" for i in range(300)]
    sts_text = "".join(lines).encode()
    path = _write(tmp_path, "d.sts", sts_text)
    for params in (dict(wordsize=8, margin=45, mismatches=0), dict(wordsize=11, margin=45, mismatches=1)):
        eng = MerPCR(**params)
        assert eng.load_sts_file(path)
        got = parity.engine_hits(eng, _records([seq]))
        want = parity.oracle_hits(params, sts_text.decode(), [seq.tobytes()])
        assert np.array_equal(got, want) and len(want) > 20000
        pos1, counts = np.unique(want[:, 1], return_counts=True)
        assert counts.max() > 300                                    # runs well past the insertion-sort limit
        eng.close()


def test_sharded_equals_whole_on_device(tmp_path):
    from merpcr_b200 import MerPCR
    contigs, sts_text, _ = _make_workload(901, [700_000, 300_000, 64, 500_001], 1200,
                                          dict(margin=50), "cfg3", False, False)
    path = _write(tmp_path, "s.sts", sts_text)
    params = dict(wordsize=11, margin=50, mismatches=1)
    recs = _records(contigs)
    whole = MerPCR(**params)
    assert whole.load_sts_file(path)
    ref = whole.search_hits(recs)
    for world in (2, 8):
        parts = []
        for rank in range(world):
            e = MerPCR(**params, shard=(rank, world))
            assert e.load_sts_file(path)
            parts.append(e.search_hits(recs))
            e.close()
        merged = np.concatenate(parts)
        order = np.lexsort((merged["rank"], merged["rec"], merged["hash_off"], merged["pos1"], merged["contig"]))
        assert np.array_equal(merged[order], ref), world


def test_append_mode_range_scans_equal_one_scan_on_device(tmp_path):
    """mpcr_ctx_set_append + mpcr_scan over consecutive ranges (cuts on any multiple of 128, inside contigs too) == one
    scan; then the pipelined `upload_and_scan` from pinned host memory == upload + scan."""
    import torch
    from merpcr_b200 import MerPCR, _capi
    rng = synth.Rng(901)
    contigs = [rng.dna(n) for n in (900_000, 12, 330_000, 2047, 2049, 70_000_000 // 64, 410_000)]
    sts = synth.make_sts_set(902, 1500, 18, 25, 100, 600)
    synth.plant_amplicons(903, [c for c in contigs if len(c) > 100_000], sts, 50, sub_mode="cfg3")
    stsf = _write(tmp_path, "s.sts", synth.sts_lines(sts))
    eng = MerPCR(wordsize=11, margin=50, mismatches=1)
    assert eng.load_sts_file(stsf)
    recs = _records(contigs)
    layout = eng.make_layout([len(c) for c in contigs])
    sh = eng.upload(layout, contigs)
    want = eng.scan(layout, sh)
    assert len(want) > 1000
    lib, cg, dev = eng._be.lib, layout["contigs"], eng._tdev
    isz = _capi.HIT_DTYPE.itemsize
    for cuts in ([int(c["gstart"]) for c in cg[1:]],
                 [128 * k for k in (7, 100, 101, 3950, 3960, 6400, 9000, 12000, 20000)],
                 [2048 * 5 + 128, 2048 * 5 + 256]):
        bounds = [layout["begin"]] + [c for c in sorted(cuts) if layout["begin"] < c < layout["end"]] + [layout["end"]]
        hits = torch.zeros(4 * len(want) * isz, dtype=torch.uint8, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        eng._be.check(lib.mpcr_ctx_set_append(eng._ctx, 1))
        try:
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                eng._be.check(lib.mpcr_scan(eng._ctx, cg.ctypes.data, len(cg), sh.plane2.data_ptr(), sh.plane4.data_ptr(),
                                            sh.valid.data_ptr(), sh.origin, sh.alloc, lo, hi, hits.data_ptr(),
                                            4 * len(want), count.data_ptr(), eng._stream()))
        finally:
            eng._be.check(lib.mpcr_ctx_set_append(eng._ctx, 0))
        n = int(count.item())
        assert n == len(want), cuts
        eng._be.check(lib.mpcr_sort_hits(eng._ctx, hits.data_ptr(), n, eng._stream()))
        got = hits[: n * isz].cpu().numpy().view(_capi.HIT_DTYPE)
        assert np.array_equal(got, want), cuts
    # the pipeline: pinned host contigs, small scan groups so that several ranges are scanned while the copy runs
    import merpcr_b200.engine as E
    pinned = [torch.from_numpy(c).pin_memory() for c in contigs]
    old = E.STREAM_SCAN_BASES
    E.STREAM_SCAN_BASES = 200_000
    try:
        sh2 = None
        for _ in range(3):                                   # steady state re-uses the buffers
            sh2, hits_t, n = eng.upload_and_scan(layout, pinned, shard=sh2)
            assert np.array_equal(eng._hits_to_host(hits_t, n), want)
        sh3, hits_t, n = eng.upload_and_scan(layout, contigs)   # pageable numpy sources
        assert np.array_equal(eng._hits_to_host(hits_t, n), want)
        # ... and as ranks of a sharded run: every rank pipelines its own range (cuts inside contigs), merged == whole
        from merpcr_b200 import multi
        for world in (2, 3):
            parts = []
            for rank in range(world):
                e = MerPCR(wordsize=11, margin=50, mismatches=1, shard=(rank, world))
                assert e.load_sts_file(stsf)
                lay = e.make_layout([len(c) for c in contigs])
                _, hits_t, n = e.upload_and_scan(lay, pinned)
                parts.append(e._hits_to_host(hits_t, n))
                e.close()
            assert np.array_equal(multi.merge_hits(parts), want), world
    finally:
        E.STREAM_SCAN_BASES = old
    assert np.array_equal(eng.search_hits(recs), want)


def test_idempotent_rescan_on_resident_planes(tmp_path):
    """Scanning the same resident planes twice gives the same sorted list (no state leaks between launches)."""
    from merpcr_b200 import MerPCR
    contigs, sts_text, _ = _make_workload(902, [2_000_000], 1000, dict(margin=50), "none", False, False)
    path = _write(tmp_path, "i.sts", sts_text)
    eng = MerPCR()
    assert eng.load_sts_file(path)
    lay = eng.make_layout([len(c) for c in contigs])
    sh = eng.upload(lay, contigs)
    a = eng.scan(lay, sh)
    b = eng.scan(lay, sh)
    assert len(a) > 500 and np.array_equal(a, b)
    assert eng.last_scan_ms > 0


def _random_fasta(seed: int, n_records: int, approx_len: int) -> bytes:
    """FASTA text with everything the loader rules care about: mixed newlines, blank lines, junk characters, '>' inside
    lines, indented headers, lower case, IUPAC letters, U (dropped), data before the first header."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKMSWBDHVXUu-*1 >\t", dtype=np.uint8)
    weights = np.array([20, 20, 20, 20, 3, 3, 3, 3, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0.3, 0.3])
    weights = weights / weights.sum()
    nl = [b"\n", b"\r\n", b"\r"]
    parts = [b"leading junk ACGT\n"]
    for r in range(n_records):
        parts.append(b" " * int(rng.integers(0, 3)) + b">rec%d some text > more" % r + nl[int(rng.integers(0, 3))])
        total = int(rng.integers(approx_len // 2, approx_len * 2))
        while total > 0:
            k = int(min(total, rng.integers(1, 120)))
            line = rng.choice(alphabet, size=k, p=weights).tobytes()
            if line.lstrip(b" \t").startswith(b">"):
                line = b"A" + line
            parts.append(line + nl[int(rng.integers(0, 3))])
            if rng.random() < 0.05:
                parts.append(nl[int(rng.integers(0, 3))])
            total -= k
    return b"".join(parts)


@pytest.mark.parametrize("seed,n_records,approx_len", [(1, 3, 2000), (2, 40, 30000), (3, 1, 3_000_000), (4, 300, 500)])
def test_device_fasta_ingest_equals_host_parser(tmp_path, monkeypatch, seed, n_records, approx_len):
    """mpcr_fasta_index + mpcr_fasta_compact on the GPU == the host parser (pinned to the reference by the goldens)."""
    from merpcr_b200 import MerPCR
    from merpcr_b200.fasta import FASTALoader
    monkeypatch.setenv("MPCR_DEVICE_INGEST_MIN_BYTES", "1")
    p = tmp_path / "x.fa"
    p.write_bytes(_random_fasta(seed, n_records, approx_len))
    eng = MerPCR()
    host = FASTALoader.load_file(str(p))
    dev = eng.load_fasta_file(str(p))
    assert len(dev) == len(host) == n_records
    assert all(r.sequence_device is not None and r.sequence_device.is_cuda for r in dev)
    for a, b in zip(dev, host):
        assert (a.defline, a.label, len(a)) == (b.defline, b.label, len(b))
        assert np.array_equal(a.sequence_bytes, b.sequence_bytes)


def test_fuzz_goldens_bit_exact_with_device_ingest(monkeypatch):
    """The reference-generated goldens once more, with every FASTA file going through the device-side ingest."""
    from merpcr_b200 import MerPCR
    monkeypatch.setenv("MPCR_DEVICE_INGEST_MIN_BYTES", "1")
    for c in goldens.fuzz_cases():
        parity.check_fuzz_case(c, MerPCR)


def test_search_from_device_resident_records(tmp_path, monkeypatch):
    """load_fasta_file (device ingest) -> search: the planes are packed straight from HBM, same hits as from host bytes."""
    from merpcr_b200 import MerPCR
    monkeypatch.setenv("MPCR_DEVICE_INGEST_MIN_BYTES", "1")
    rng = synth.Rng(901)
    contigs = [rng.dna(n) for n in (150000, 80000)]
    sts = synth.make_sts_set(902, 100, 18, 25, 100, 700)
    synth.plant_amplicons(903, contigs, sts, 50, sub_mode="cfg3")
    stsf = _write(tmp_path, "s.sts", synth.sts_lines(sts))
    fa = b"".join(b">c%d\n" % i + b"\n".join(c[j:j + 60].tobytes() for j in range(0, len(c), 60)) + b"\n"
                  for i, c in enumerate(contigs))
    faf = _write(tmp_path, "g.fa", fa)
    eng = MerPCR(mismatches=1)
    assert eng.load_sts_file(stsf)
    recs = eng.load_fasta_file(faf)
    assert all(r.sequence_device is not None for r in recs)
    got = parity.engine_hits(eng, recs)
    want = parity.oracle_hits(dict(wordsize=11, margin=50, mismatches=1), synth.sts_lines(sts).decode(),
                              [c.tobytes() for c in contigs])
    assert np.array_equal(got, want) and len(got) > 50
    assert eng.last_h2d_bytes == 0


def test_seed_extension_two_tables_equal_one_on_device(tmp_path, monkeypatch):
    """mpcr_ctx_set_seed_extension on the GPU: exact search keyed on 11-letter words (+ the table of records that cannot
    be extended) == the plain one-table search == the oracle."""
    from merpcr_b200 import MerPCR
    rng = synth.Rng(711)
    contigs = [rng.dna(n) for n in (400_000, 150_000, 9)]
    sts = synth.make_sts_set(712, 30000, 9, 25, 60, 400)
    sts["p1"][::7, 9] = ord("N")
    expected = synth.plant_amplicons(713, contigs[:2], sts, 30, plant_count=200)
    text = synth.sts_lines(sts)
    stsf = _write(tmp_path, "s.sts", text)
    params = dict(wordsize=8, margin=30, mismatches=0)
    want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
    for flag, parts in (("0", "1"), ("1", "1"), ("1", "3")):
        monkeypatch.setenv("MPCR_SEED_EXTENSION", flag)
        monkeypatch.setenv("MPCR_SEED_PARTS", parts)      # several extended tables, every STS line in exactly one
        eng = MerPCR(**params)
        assert eng.load_sts_file(stsf)
        assert len(eng._ctx_exts) == (int(parts) if flag == "1" else 0)
        got = parity.engine_hits(eng, _records(contigs))
        assert np.array_equal(got, want), flag
        eng.close()
    assert len(want) >= len(expected) > 100


def test_fuzz_goldens_with_seed_extension_on_device(monkeypatch):
    from merpcr_b200 import MerPCR
    monkeypatch.setenv("MPCR_SEED_EXTENSION", "1")
    n = 0
    for c in goldens.fuzz_cases():
        if c["params"].get("mismatches", 0) == 0 and not c["params"].get("iupac_mode", 0):
            parity.check_fuzz_case(c, MerPCR)
            n += 1
    assert n > 5


@pytest.mark.parametrize("wordsize,mismatches,block_env,n_sts", [(8, 1, "1", 30000), (8, 1, "2", 3000), (10, 2, "1", 30000),
                                                                 (6, 1, "3", 3000), (11, 1, "1", 30000)])
def test_block_tables_equal_one_table_on_device(tmp_path, monkeypatch, wordsize, mismatches, block_env, n_sts):
    """mpcr_ctx_set_seed_blocks on the GPU: a search that allows mismatches keyed on seed + one of N + 1 blocks (split
    keys through the sparse scanner, direct and open-addressed slot tables) + the table of the records that cannot be
    keyed that way == the plain one-table search == the oracle, every site once."""
    from merpcr_b200 import MerPCR
    contigs, text, expected = synth.block_table_case(900 + wordsize, wordsize, mismatches, n_sts=n_sts,
                                                     contig_lens=(400_000, 150_000, 11), plant_count=400)
    stsf = _write(tmp_path, "s.sts", text)
    params = dict(wordsize=wordsize, margin=30, mismatches=mismatches)
    want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
    assert len(want) > 150
    for flag, parts in (("0", "1"), (block_env, "1"), (block_env, "3")):
        monkeypatch.setenv("MPCR_SEED_BLOCKS", flag)
        monkeypatch.setenv("MPCR_SEED_PARTS", parts)
        eng = MerPCR(**params)
        assert eng.load_sts_file(stsf)
        assert len(eng._ctx_exts) == (int(parts) * (mismatches + 1) if flag != "0" else 0)
        got = parity.engine_hits(eng, _records(contigs))
        assert np.array_equal(got, want), (flag, parts)
        eng.close()


def test_fuzz_goldens_with_block_tables_on_device(monkeypatch):
    from merpcr_b200 import MerPCR
    monkeypatch.setenv("MPCR_SEED_BLOCKS", "1")
    n = 0
    for c in goldens.fuzz_cases():
        if c["params"].get("mismatches", 0) >= 1 and not c["params"].get("iupac_mode", 0):
            parity.check_fuzz_case(c, MerPCR)
            n += 1
    assert n > 20


def test_cli_subprocess_on_the_fixture(tmp_path):
    """`python -m merpcr_b200 <sts> <fasta>` -- the reference's CLI surface end to end on the GPU (cli.py:217-266)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "hits.txt"
    env = dict(os.environ, PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-m", "merpcr_b200", goldens.FIXTURE_STS, goldens.FIXTURE_FA, "-W", "11", "-N", "0",
                        "-M", "50", "-X", "1", "-T", "1", "-Q", "0", "-O", str(out)], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == goldens.FIXTURE_LINE
    assert "Reading STS file" in r.stderr and "Total hits found: 1" in r.stderr
    # me-PCR style K=V arguments and stdout output (cli.py:19-62)
    r = subprocess.run([sys.executable, "-m", "merpcr_b200", goldens.FIXTURE_STS, goldens.FIXTURE_FA, "W=11", "M=50", "N=0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout == goldens.FIXTURE_LINE
    r = subprocess.run([sys.executable, "-m", "merpcr_b200", goldens.FIXTURE_STS, goldens.FIXTURE_FA, "-W", "2"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 2


@pytest.mark.parametrize("cap", ["0", "3"])
def test_survivor_list_overflow_verifies_in_place(tmp_path, monkeypatch, cap):
    """A full survivor list must not lose anything: the scanner verifies the overflow itself (verify_serial)."""
    from merpcr_b200 import MerPCR
    monkeypatch.setenv("MPCR_SURVIVOR_CAP", cap)
    name, lengths, n_sts, params, sub_mode, decorate, ranged = CASES[1]
    contigs, sts_text, expected = _make_workload(4242, lengths, n_sts, params, sub_mode, decorate, ranged)
    path = _write(tmp_path, "w.sts", sts_text)
    eng = MerPCR(**params)
    assert eng.load_sts_file(path)
    got = parity.engine_hits(eng, _records(contigs))
    want = parity.oracle_hits(params, sts_text.decode(), [c.tobytes() for c in contigs])
    assert np.array_equal(got, want) and len(got) > 100


def test_true_strands_option_on_device():
    """mpcr_ctx_set_true_strands, pinned to the reference through per-strand identities."""
    from merpcr_b200 import MerPCR
    from test_host_logic import _true_strands_check
    _true_strands_check(MerPCR, _records)


@pytest.mark.parametrize("n,long_run", [(2, False), (37, False), (1000, True), (9000, False), (9000, True), (16384, False),
                                        (16385, False), (70000, False), (70000, True), (131072, False), (131073, False),
                                        (300000, True)])
def test_sort_with_the_count_on_the_device(n, long_run):
    """mpcr_sort_hits_dev (count read on the device; bucket sort on the global coordinate for lists of up to 2^17 hits,
    radix passes + tie kernels for longer ones and for lists that pile up in one place; any hint) and mpcr_sort_hits
    order random hit lists -- many ties on (contig, pos1), optionally one run of several hundred equal positions --
    exactly like a lexicographic sort on the reference's order key."""
    import torch
    from merpcr_b200 import MerPCR, _capi
    eng = MerPCR()
    lib = eng._be.lib
    rng = np.random.default_rng(n)
    h = np.zeros(n, dtype=_capi.HIT_DTYPE)
    h["contig"] = rng.integers(0, 24, n)
    h["pos1"] = rng.integers(0, max(4, n // 3), n)             # plenty of equal (contig, pos1)
    if long_run:
        h["contig"][:700] = 5
        h["pos1"][:700] = 77                                    # a long run: bucket overflow -> radix + order_long_runs
    h["pos2"] = h["pos1"] + rng.integers(50, 900, n)
    h["rec"] = rng.permutation(n).astype(np.uint32)            # unique: the order key is total
    h["rank"] = rng.integers(0, 100, n)
    h["hash_off"] = rng.integers(0, 3, n)
    order = np.lexsort((h["rank"], h["rec"], h["hash_off"], h["pos1"], h["contig"]))
    want = h[order]
    # the digit counts of the radix passes come from the layout of the last scan: give the context one
    layout = eng.make_layout([max(4, n // 3) + 1000] * 24)
    cg = layout["contigs"]
    eng._be.check(lib.mpcr_scan_prepare(eng._ctx, cg.ctypes.data, len(cg), 0, 0, layout["total"], eng._stream()))
    dev = torch.device("cuda", eng.device)
    cap = n + 1000
    for hint in (None, 0, 5, n, 10 * n + 50000):
        buf = torch.zeros(cap * h.itemsize, dtype=torch.uint8, device=dev)
        buf[: n * h.itemsize] = torch.from_numpy(h.view(np.uint8).copy()).to(dev)
        if hint is None:
            eng._be.check(lib.mpcr_sort_hits(eng._ctx, buf.data_ptr(), n, eng._stream()))
        else:
            count = torch.tensor([n], dtype=torch.int64, device=dev)
            eng._be.check(lib.mpcr_sort_hits_dev(eng._ctx, buf.data_ptr(), count.data_ptr(), cap, hint, eng._stream()))
        torch.cuda.synchronize()
        got = buf[: n * h.itemsize].cpu().numpy().view(_capi.HIT_DTYPE)
        assert np.array_equal(got, want), (n, hint)
    # a count beyond the capacity sorts the records that were stored
    buf = torch.from_numpy(h.view(np.uint8).copy()).to(dev)
    count = torch.tensor([n + 12345], dtype=torch.int64, device=dev)
    eng._be.check(lib.mpcr_sort_hits_dev(eng._ctx, buf.data_ptr(), count.data_ptr(), n, 0, eng._stream()))
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy().view(_capi.HIT_DTYPE), want)
    eng.close()


def test_scan_rejects_planes_that_are_too_small(tmp_path):
    """mpcr_scan checks the plane extent it is told about against what the kernels read (whole units + read-ahead,
    the last mate window, nothing in front of the origin) and fails with MPCR_EINVAL instead of reading out of bounds."""
    import torch
    from merpcr_b200 import MerPCR
    rng = synth.Rng(31)
    contigs = [rng.dna(50_000), rng.dna(20_000)]
    sts = synth.make_sts_set(32, 50, 18, 25, 100, 600)
    synth.plant_amplicons(33, contigs, sts, 50, sub_mode="none")
    path = _write(tmp_path, "b.sts", synth.sts_lines(sts))
    eng = MerPCR()
    assert eng.load_sts_file(path)
    lib = eng._be.lib
    layout = eng.make_layout([len(c) for c in contigs])
    sh = eng.upload(layout, contigs)
    cg = layout["contigs"]
    dev = torch.device("cuda", eng.device)
    hits = torch.empty(4096 * 24, dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)

    def scan(origin, plane_bases, lo, hi):
        return lib.mpcr_scan(eng._ctx, cg.ctypes.data, len(cg), sh.plane2.data_ptr(), sh.plane4.data_ptr(),
                             sh.valid.data_ptr(), origin, plane_bases, lo, hi, hits.data_ptr(), 4096, count.data_ptr(),
                             eng._stream())
    assert scan(sh.origin, sh.alloc, sh.begin, sh.end) == 0
    n_ok = int(count.item())
    assert n_ok >= 30
    last = int(cg[1]["gstart"]) + int(cg[1]["length"])
    for too_small in (0, 128, last - 1, last + 64):            # even the true genome length lacks the staging read-ahead
        with pytest.raises(ValueError):
            eng._be.check(scan(sh.origin, too_small, sh.begin, sh.end))
    with pytest.raises(ValueError):                            # a range that starts in front of the planes' origin
        eng._be.check(scan(1 << 20, sh.alloc, 0, sh.end))
    assert scan(sh.origin, sh.alloc, sh.begin, sh.end) == 0 and int(count.item()) == n_ok
    eng.close()


def test_host_packed_nibble_ingest_builds_the_same_planes(tmp_path):
    """Host-resident sequence shipped as packed nibbles (mpcr_host_pack_nibbles -> H2D into plane4 -> mpcr_derive_planes)
    must leave exactly the planes the ASCII path (H2D -> pack_kernel) builds -- odd lengths, partial last strips, every
    FASTA letter in both cases -- and the same hits; a 'U' outside IUPAC mode falls back to ASCII for its piece."""
    import torch
    from merpcr_b200 import MerPCR
    rng = synth.Rng(909)
    contigs = [rng.dna(3_000_001), rng.dna(777), rng.dna(64 * 4096), rng.dna(1_234_567)]
    letters = np.frombuffer(b"NRYKMSWBDHVXnrykmswbdhvxacgt", dtype=np.uint8)
    for c in contigs:
        pos = rng.ints(0, len(c) - 1, max(4, len(c) // 500))
        c[pos] = letters[rng.ints(0, len(letters) - 1, len(pos))]
    sts = synth.make_sts_set(910, 1500, 18, 25, 100, 700)
    synth.plant_amplicons(911, contigs, sts, 50, sub_mode="cfg3")
    path = _write(tmp_path, "h.sts", synth.sts_lines(sts))
    for iupac, with_u in ((0, False), (1, True), (0, True)):
        seqs = [c.copy() for c in contigs]
        if with_u:
            seqs[3][[5, 600_000]] = ord("U")
        planes, hits = [], []
        for host_pack in (True, False):
            eng = MerPCR(wordsize=11, margin=50, mismatches=1, iupac_mode=iupac)
            eng.host_pack = host_pack
            assert eng.load_sts_file(path)
            layout = eng.make_layout([len(c) for c in seqs])
            sh, ht, n = eng.upload_and_scan(layout, [torch.from_numpy(c) for c in seqs])
            torch.cuda.synchronize()
            used = (layout["total"] + 127) // 128 * 128
            planes.append((sh.plane2[: used // 4].cpu(), sh.plane4[: used // 2].cpu(), sh.valid[: used // 8].cpu()))
            hits.append(eng._hits_to_host(ht, n))
            if host_pack:
                expect = sum((len(c) + 1) // 2 for c in seqs)
                if with_u and not iupac:
                    expect += len(seqs[3]) - (len(seqs[3]) + 1) // 2      # that contig went up as ASCII
                assert eng.last_h2d_bytes == expect
            else:
                assert eng.last_h2d_bytes == sum(len(c) for c in seqs)
            eng.close()
        for a, b in zip(*planes):
            assert torch.equal(a, b)
        assert np.array_equal(hits[0], hits[1]) and len(hits[0]) > 1000


def test_position_sampling_equals_the_unsampled_search_on_device(tmp_path, monkeypatch):
    """mpcr_ctx_set_sampling on the GPU (sampled_scan_kernel: every S-th position probed through the global-memory
    filter) -- same matrix as the CPU-tier test: strides, with and without the seed extension, shards."""
    from merpcr_b200 import MerPCR
    from test_host_logic import _sampling_case
    contigs, text, stsf, expected = _sampling_case(tmp_path, "samp_gpu")
    params = dict(wordsize=8, margin=30, mismatches=0)
    want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
    assert len(want) >= len(expected) > 50
    for ext, stride in (("0", "0"), ("0", "3"), ("1", "3"), ("1", "2"), ("0", "5"), ("1", "7")):
        monkeypatch.setenv("MPCR_SEED_EXTENSION", ext)
        monkeypatch.setenv("MPCR_SAMPLING", stride)
        eng = MerPCR(**params)
        assert eng.load_sts_file(stsf)
        assert (eng._ctx_samp is not None) == (stride != "0")
        got = parity.engine_hits(eng, _records(contigs))
        assert np.array_equal(got, want), (ext, stride)
        eng.close()
    monkeypatch.setenv("MPCR_SEED_EXTENSION", "1")
    monkeypatch.setenv("MPCR_SAMPLING", "3")
    parts = []
    for shard in (None, (0, 3), (1, 3), (2, 3)):
        eng = MerPCR(**params, shard=shard)
        assert eng.load_sts_file(stsf)
        parts.append(eng.search_hits(_records(contigs)))
        eng.close()
    merged = np.concatenate(parts[1:])
    order = np.lexsort((merged["rank"], merged["rec"], merged["hash_off"], merged["pos1"], merged["contig"]))
    assert np.array_equal(merged[order], parts[0])


def test_fuzz_goldens_with_position_sampling_on_device(monkeypatch):
    from merpcr_b200 import MerPCR
    for stride in ("2", "3"):
        monkeypatch.setenv("MPCR_SAMPLING", stride)
        n = 0
        for c in goldens.fuzz_cases():
            if c["params"].get("mismatches", 0) == 0 and not c["params"].get("iupac_mode", 0) and \
                    c["params"].get("wordsize", 11) < 16:
                parity.check_fuzz_case(c, MerPCR)
                n += 1
        assert n > 5


def test_two_steps_in_flight_on_device(tmp_path):
    from merpcr_b200 import MerPCR
    from test_host_logic import _two_in_flight_check
    _two_in_flight_check(MerPCR, _records, tmp_path)
