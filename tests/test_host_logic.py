"""CPU tier: the Python host (parsing, alphabet, layout, sharding, formatting, CLI) and the shared per-position
semantics of merpcr_b200/csrc/mpcr_core.cuh, exercised through tests/host_emul's serial emulation of the C ABI.
The emulation is test infrastructure (explicitly injected here); the GPU parity tests never use it."""
import io
import logging
import os
import sys

import numpy as np
import pytest

import emul
import goldens
import parity
import synth


@pytest.fixture(scope="module", autouse=True)
def _emulated_backend():
    emul.inject()
    yield
    emul.restore()


def _engine(**kw):
    from merpcr_b200 import MerPCR
    return MerPCR(**kw)


def test_fixture_golden_line(tmp_path):
    eng = _engine()
    assert eng.load_sts_file(goldens.FIXTURE_STS)
    assert len(eng.sts_records) == 6 and eng.max_pcr_size == 193        # reference test_comprehensive.py:42
    assert sorted(eng.sts_table) == sorted([3638181, 3526114, 2555953, 3737721, 2062650, 476488])
    recs = eng.load_fasta_file(goldens.FIXTURE_FA)
    assert [(r.label, len(r.sequence)) for r in recs] == [("L78833", 117143)]
    out = tmp_path / "o.txt"
    assert eng.search(recs, str(out)) == 1 and eng.total_hits == 1
    assert out.read_text() == goldens.FIXTURE_LINE


def test_fuzz_goldens_through_host_stack():
    from merpcr_b200 import MerPCR
    for c in goldens.fuzz_cases():
        parity.check_fuzz_case(c, MerPCR)


def test_constructor_validation_and_attributes():
    # reference tests/test_engine_internals.py:202-235, test_core_engine_comprehensive.py:20-59
    from merpcr_b200 import MerPCR
    e = MerPCR()
    assert (e.wordsize, e.margin, e.mismatches, e.three_prime_match, e.iupac_mode, e.default_pcr_size, e.threads,
            e.max_sts_line_length) == (11, 50, 0, 1, 0, 240, 1, 1022)
    assert e.sts_records == [] and e.sts_table == {} and e.max_pcr_size == 0 and e.total_hits == 0
    for bad in (dict(wordsize=2), dict(wordsize=17), dict(mismatches=-1), dict(mismatches=11), dict(margin=-1),
                dict(margin=10001), dict(three_prime_match=-1), dict(default_pcr_size=0), dict(default_pcr_size=10001)):
        with pytest.raises(ValueError):
            MerPCR(**bad)


def test_private_helpers_known_answers():
    from merpcr_b200 import MerPCR
    from merpcr_b200.utils import hash_value, reverse_complement
    assert MerPCR(wordsize=4)._hash_value("ATCG") == (0, 54) == hash_value("ATCG", 4)
    assert MerPCR(wordsize=8)._hash_value("TTTTTTTT") == (0, 65535)
    assert MerPCR(wordsize=8)._hash_value("NNNATCGATCGATCG")[0] == 3
    assert MerPCR()._hash_value("ACGUACGUACG") == (0, 444102)
    assert MerPCR(wordsize=8)._hash_value("ACGT") == (-1, 0)
    e = MerPCR(mismatches=1, three_prime_match=2)
    assert e._reverse_complement("ACGU-XZ") == "NXNACGT" == reverse_complement("ACGU-XZ")
    assert e._reverse_complement("RWYS") == "SRWY"
    assert e._compare_seqs("TTCGATCG", "ATCGATCG", "+") and not e._compare_seqs("ATCGATCA", "ATCGATCG", "+")
    assert not e._compare_seqs("TTCGATCG", "ATCGATCG", "-") and e._compare_seqs("ATCGATCA", "ATCGATCG", "-")
    i = MerPCR(iupac_mode=1)
    assert i._compare_seqs("A", "R", "+") and not i._compare_seqs("C", "R", "+") and i._compare_seqs("N", "A", "+")


def test_loader_failure_modes(tmp_path):
    # reference tests/test_core_engine_comprehensive.py:85-140, test_io_modules.py:110-148
    e = _engine()
    empty = tmp_path / "e.sts"
    empty.write_text("")
    assert e.load_sts_file(str(empty)) is False
    assert e.load_fasta_file(str(empty)) == []
    with pytest.raises(FileNotFoundError):
        e.load_sts_file(str(tmp_path / "missing.sts"))
    with pytest.raises(FileNotFoundError):
        e.load_fasta_file(str(tmp_path / "missing.fa"))
    bad = tmp_path / "b.sts"
    bad.write_text("id\tACGTACGTACGT\tACGTACGTACGT\n")
    assert e.load_sts_file(str(bad)) is False and e.sts_records == []
    fa = tmp_path / "f.fa"
    fa.write_text(">seq1\nATCG123NNNN456ATCG\nWXYZ789GCTA\n\n>seq2 only header\n")
    recs = e.load_fasta_file(str(fa))
    assert [(r.label, r.sequence) for r in recs] == [("seq1", "ATCGNNNNATCGWXYGCTA"), ("seq2", "")]


def test_search_edge_cases(tmp_path, capsys):
    from merpcr_b200 import FASTARecord, MerPCR
    e = MerPCR(wordsize=8)
    assert e.search([]) == 0                                   # nothing loaded, nothing to search
    sts = tmp_path / "a.sts"
    p1, p2 = "ACGTTGCAAGGCTA", "TTGACCGGTATCAG"
    sts.write_text(f"S1\t{p1}\t{p2}\t60\talias one\nS2\t{p1}\t{p2}\t60\n")
    assert e.load_sts_file(str(sts))
    filler = "A" * (60 - len(p1) - len(p2))
    seq = "CCCCC" + p1 + filler + p2 + "GGGGG"
    recs = [FASTARecord(defline=">c1 test", sequence=seq), FASTARecord(defline=">empty", sequence=""),
            FASTARecord(defline="noangle x", sequence=seq.lower(), label="given")]
    n = e.search(recs)                                          # stdout; identical primers list in file order (A.7)
    out = capsys.readouterr().out
    assert n == 4 and out == ("c1\t6..65\tS1\talias one\t(+)\nc1\t6..65\tS2\t\t(+)\n"
                              "given\t6..65\tS1\talias one\t(+)\ngiven\t6..65\tS2\t\t(+)\n")
    assert e.search(recs, "STDOUT") == 4
    capsys.readouterr()


def test_api_sequences_with_unusual_letters():
    """Sequences built through the API may carry letters FASTA files cannot: U hashes as T but
    only equals U outside IUPAC mode; any other letter matches only itself."""
    from merpcr_b200 import FASTARecord, MerPCR
    from oracle.oracle import Oracle
    import tempfile
    p1, p2 = "ACGUUGCAAGGCTA", "TTGACCGGTATCAG"
    seq = "CC" + p1 + "A" * 30 + p2 + "GG" + "ACGTTGCAAGGCTA" + "A" * 30 + p2
    for iupac in (0, 1):
        with tempfile.NamedTemporaryFile("w", suffix=".sts", delete=False) as f:
            f.write(f"S1\t{p1}\t{p2}\t58\n")
        o = Oracle(wordsize=8, iupac_mode=iupac)
        assert o.load_sts_file(f.name)
        e = MerPCR(wordsize=8, iupac_mode=iupac)
        assert e.load_sts_file(f.name)
        import io as _io, contextlib
        buf = _io.StringIO()
        with contextlib.redirect_stdout(buf):
            e.search([FASTARecord(">u", seq)])
        assert buf.getvalue() == o.search_text("u", seq)[1] and buf.getvalue().count("\n") == (2 if iupac else 1)
        os.unlink(f.name)


def test_sharded_scan_equals_whole(tmp_path):
    """bp-balanced shards with halos give, merged, exactly the unsharded hit list."""
    from merpcr_b200 import FASTARecord, MerPCR
    rng = synth.Rng(77)
    contigs = [rng.dna(n) for n in (90000, 70001, 1500, 130000)]
    sts = synth.make_sts_set(78, 120, 18, 25, 100, 700)
    expected = synth.plant_amplicons(79, contigs, sts, 50, sub_mode="cfg3")
    stsf = tmp_path / "s.sts"
    stsf.write_bytes(synth.sts_lines(sts))
    params = dict(wordsize=11, margin=50, mismatches=1)
    recs = [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)]
    whole = MerPCR(**params)
    assert whole.load_sts_file(str(stsf))
    ref = parity.engine_hits(whole, recs)
    want = parity.oracle_hits(params, synth.sts_lines(sts).decode(), [c.tobytes() for c in contigs])
    assert np.array_equal(ref, want) and len(ref) >= len(expected) > 50
    found = {(r[0], r[1], r[2]) for r in ref.tolist()}
    assert all((ci, a, b) in found for ci, a, b, _, _ in expected)      # planted truth, independent of the oracle
    for world in (2, 3, 8):
        parts = []
        for rank in range(world):
            e = MerPCR(**params, shard=(rank, world))
            assert e.load_sts_file(str(stsf))
            parts.append(e.search_hits(recs))
        merged = np.concatenate(parts)
        order = np.lexsort((merged["rank"], merged["rec"], merged["hash_off"], merged["pos1"], merged["contig"]))
        assert np.array_equal(merged[order], whole.search_hits(recs)), world


def test_append_mode_range_scans_equal_one_scan(tmp_path):
    """mpcr_ctx_set_append + mpcr_scan over consecutive ranges of the padded coordinate (what `upload_and_scan` does
    while a genome is still uploading): ranges cut anywhere on a multiple of 128 -- inside contigs too -- own disjoint
    2048-position units, so the appended hits are exactly the hits of one scan."""
    import ctypes as C
    import torch
    from merpcr_b200 import FASTARecord, MerPCR, _capi
    rng = synth.Rng(901)
    contigs = [rng.dna(n) for n in (50000, 12, 33000, 2047, 2049, 41000)]
    sts = synth.make_sts_set(902, 150, 18, 25, 100, 600)
    synth.plant_amplicons(903, [contigs[0], contigs[2], contigs[5]], sts, 50, sub_mode="cfg3")
    stsf = tmp_path / "s.sts"
    stsf.write_bytes(synth.sts_lines(sts))
    eng = MerPCR(wordsize=11, margin=50, mismatches=1)
    assert eng.load_sts_file(str(stsf))
    recs = [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)]
    want = eng.search_hits(recs)
    assert len(want) > 100
    layout = eng.make_layout([len(c) for c in contigs])
    sh = eng.upload(layout, [r.sequence_bytes for r in recs])
    lib, cg = eng._be.lib, layout["contigs"]
    isz = _capi.HIT_DTYPE.itemsize
    for cuts in ([int(c["gstart"]) for c in cg[1:]],                       # contig boundaries
                 [128 * k for k in (7, 100, 101, 395, 396, 640, 900, 1200)],  # anywhere, inside contigs
                 [2048 * 5 + 128, 2048 * 5 + 256]):                        # two cuts inside one unit
        bounds = [layout["begin"]] + [c for c in sorted(cuts) if layout["begin"] < c < layout["end"]] + [layout["end"]]
        hits = torch.zeros(4 * len(want) * isz, dtype=torch.uint8)
        count = torch.zeros(1, dtype=torch.int64)
        eng._be.check(lib.mpcr_ctx_set_append(eng._ctx, 1))
        try:
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                eng._be.check(lib.mpcr_scan(eng._ctx, cg.ctypes.data, len(cg), sh.plane2.data_ptr(), sh.plane4.data_ptr(),
                                            sh.valid.data_ptr(), sh.origin, sh.alloc, lo, hi, hits.data_ptr(),
                                            4 * len(want), count.data_ptr(), 0))
        finally:
            eng._be.check(lib.mpcr_ctx_set_append(eng._ctx, 0))
        n = int(count.item())
        assert n == len(want), cuts
        eng._be.check(lib.mpcr_sort_hits(eng._ctx, hits.data_ptr(), n, 0))
        got = hits[: n * isz].numpy().view(_capi.HIT_DTYPE)
        assert np.array_equal(got, want), cuts


def test_pipelined_search_scans_contig_groups_as_they_complete(tmp_path, monkeypatch):
    """`upload_and_scan` (what `search` runs for host-resident sequences): contig groups are scanned in append mode as
    soon as they are packed -- several mpcr_scan calls, one hit buffer, one sort -- and give exactly the hits of
    upload + one scan; also as ranks of a sharded run (ranges cut inside contigs at the shard bounds)."""
    import merpcr_b200.engine as E
    from merpcr_b200 import FASTARecord, MerPCR, multi
    monkeypatch.setattr(E, "STREAM_SCAN_BASES", 30000)
    rng = synth.Rng(1201)
    contigs = [rng.dna(n) for n in (50000, 12, 33000, 2047, 2049, 41000, 10, 29000)]
    sts = synth.make_sts_set(1202, 200, 18, 25, 100, 600)
    synth.plant_amplicons(1203, [c for c in contigs if len(c) > 20000], sts, 50, sub_mode="cfg3")
    f = tmp_path / "p.sts"
    f.write_bytes(synth.sts_lines(sts))
    params = dict(wordsize=11, margin=50, mismatches=1)
    eng = MerPCR(**params)
    assert eng.load_sts_file(str(f))
    layout = eng.make_layout([len(c) for c in contigs])
    want = eng.scan(layout, eng.upload(layout, contigs))
    assert len(want) > 100
    before = eng.gpu_launches
    _, hits_t, n = eng.upload_and_scan(layout, contigs)
    assert eng.gpu_launches - before >= 4                      # several range scans (the emulation counts one each)
    assert np.array_equal(eng._hits_to_host(hits_t, n), want)
    for world in (2, 3):
        parts = []
        for rank in range(world):
            e = MerPCR(**params, shard=(rank, world))
            assert e.load_sts_file(str(f))
            lay = e.make_layout([len(c) for c in contigs])
            _, hits_t, n = e.upload_and_scan(lay, contigs)
            parts.append(e._hits_to_host(hits_t, n))
        assert np.array_equal(multi.merge_hits(parts), want), world
    # a hit buffer that is too small: the pipeline notices at the end and rescans the resident planes with room
    import torch
    sh = eng._prepare_shard(layout, None)
    sh.hits = torch.empty(8 * want.dtype.itemsize, dtype=torch.uint8)
    _, hits_t, n = eng.upload_and_scan(layout, contigs, shard=sh)
    assert np.array_equal(eng._hits_to_host(hits_t, n), want)


def test_long_primers_equal_the_oracle(tmp_path):
    """Primers of 30..120 bases (multi-word compare, no hoisted primer view past 32 bases) through the host stack."""
    from merpcr_b200 import FASTARecord, MerPCR
    for seed, params in ((41, dict(wordsize=11, margin=50, mismatches=2, three_prime_match=1)),
                         (47, dict(wordsize=9, margin=30, mismatches=1, three_prime_match=0, iupac_mode=1))):
        rng = synth.Rng(seed)
        contigs = [rng.dna(n) for n in (120_000, 60_000, 150)]
        sts = synth.make_sts_set(seed + 1, 120, 30, 120, 260, 1000)
        synth.plant_amplicons(seed + 2, contigs[:2], sts, params["margin"], sub_mode="cfg3", plant_count=60)
        sts["p1"][::9, 40] = ord("N")
        sts["p2"][::7, 3] = ord("R")
        text = synth.sts_lines(sts)
        f = tmp_path / f"s{seed}.sts"
        f.write_bytes(text)
        eng = MerPCR(**params)
        assert eng.load_sts_file(str(f))
        got = parity.engine_hits(eng, [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)])
        want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
        assert got.shape == want.shape and np.array_equal(got, want) and len(want) > 20, params


def test_cli_in_process(tmp_path, monkeypatch, capsys):
    # reference tests/test_cli.py, test_cli_enhanced.py:23-156 (flag set, K=V conversion, exit codes)
    from merpcr_b200 import cli
    assert cli.convert_mepcr_arguments(["M=50", "N=1", "P=3", "-help", "x.sts", "O=out"]) == \
        ["-M", "50", "-N", "1", "--help", "x.sts", "-O", "out"]
    out = tmp_path / "hits.txt"
    monkeypatch.setattr(sys, "argv", ["merpcr", goldens.FIXTURE_STS, goldens.FIXTURE_FA, "W=11", "-N", "0", "-O", str(out)])
    assert cli.main() == 0 and out.read_text() == goldens.FIXTURE_LINE
    monkeypatch.setattr(sys, "argv", ["merpcr", goldens.FIXTURE_STS, goldens.FIXTURE_FA, "-M", "7"])
    assert cli.main() == 0 and capsys.readouterr().out == ""
    monkeypatch.setattr(sys, "argv", ["merpcr", str(tmp_path / "nope.sts"), goldens.FIXTURE_FA])
    assert cli.main() == 1
    for bad in (["-W", "2"], ["-N", "11"], ["-M", "10001"], ["-Z", "0"], ["-T", "0"], ["-I", "2"]):
        monkeypatch.setattr(sys, "argv", ["merpcr", goldens.FIXTURE_STS, goldens.FIXTURE_FA] + bad)
        with pytest.raises(SystemExit) as ei:
            cli.main()
        assert ei.value.code == 2
    capsys.readouterr()


def test_verbose_log_lines(tmp_path, caplog):
    # reference tests/test_cli.py:45-56 greps these strings with -Q 0
    e = _engine()
    with caplog.at_level(logging.INFO, logger="merpcr"):
        e.load_sts_file(goldens.FIXTURE_STS)
        e.search(e.load_fasta_file(goldens.FIXTURE_FA), str(tmp_path / "o"))
    text = caplog.text
    assert "Reading STS file" in text and "Processing sequence: L78833 (117143 bp)" in text
    assert "Total hits found: 1" in text and "Reading FASTA file" in text


FASTA_TEXTS = [
    b">a desc\nACGT\nacgtn\n>b\n\nNNNN\n",
    b"junk before\nACGT\n>first\r\nAC GT\r\n  >second  \r\nTTTT>AAAA\r\n>third",
    b"\n\n  \t>x\rACGTRYKM\r\rBDHVSWX\r>y\rUUUU1234acgu\r",
    b"no header at all\nACGT\n",
    b">only header\n",
    b">\x1f>odd\nAC\x0bGT\n\x0c>z\nGG-GG*\n",
    b">h1\nACGT\n>h2 > not a header marker\nAAAA >CCCC\n>h3\n" + b"ACGTN" * 3000 + b"\n",
]


@pytest.mark.parametrize("i", range(len(FASTA_TEXTS)))
def test_device_ingest_protocol_equals_host_parser(tmp_path, monkeypatch, i):
    """The mpcr_fasta_index / mpcr_fasta_compact protocol (driven by merpcr_b200/fasta.py) gives exactly what the host
    parser gives -- which test_fuzz_goldens pins to the reference's FASTALoader."""
    from merpcr_b200.fasta import FASTALoader
    monkeypatch.setenv("MPCR_DEVICE_INGEST_MIN_BYTES", "1")
    p = tmp_path / "x.fa"
    p.write_bytes(FASTA_TEXTS[i])
    eng = _engine()
    host = FASTALoader.load_file(str(p))
    try:
        dev = FASTALoader.load_file(str(p), engine=eng)
    except IndexError:   # a bare '>' header raises in FASTARecord (models.py:43-49) on both paths
        with pytest.raises(IndexError):
            FASTALoader.load_file(str(p))
        return
    assert [(r.defline, r.label, r.sequence) for r in dev] == [(r.defline, r.label, r.sequence) for r in host]
    assert all(r.sequence_device is not None for r in dev)


def test_seed_extension_two_tables_equal_one(tmp_path, monkeypatch):
    """Exact searches may be keyed on 11-letter words (mpcr_ctx_set_seed_extension): records that extend + records
    that do not, scanned separately and merged, must give the one-table result (== the oracle)."""
    from merpcr_b200 import FASTARecord, MerPCR
    rng = synth.Rng(611)
    contigs = [rng.dna(n) for n in (60000, 25000, 9)]
    sts = synth.make_sts_set(612, 4000, 9, 25, 60, 400)       # short primers: many cannot be extended to 11
    sts["p1"][::7, 9] = ord("N")                                # ... and some have an ambiguity right after the seed
    expected = synth.plant_amplicons(613, contigs[:2], sts, 30, plant_count=60)
    text = synth.sts_lines(sts)
    stsf = tmp_path / "s.sts"
    stsf.write_bytes(text)
    params = dict(wordsize=8, margin=30, mismatches=0)
    recs = [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)]
    want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
    for flag, parts in (("0", "1"), ("1", "1"), ("1", "3")):
        monkeypatch.setenv("MPCR_SEED_EXTENSION", flag)
        monkeypatch.setenv("MPCR_SEED_PARTS", parts)      # several extended tables, every STS line in exactly one
        eng = MerPCR(**params)
        assert eng.load_sts_file(str(stsf))
        assert len(eng._ctx_exts) == (int(parts) if flag == "1" else 0)
        got = parity.engine_hits(eng, recs)
        assert np.array_equal(got, want), flag
        eng.close()
    assert len(want) >= len(expected) > 20


@pytest.mark.parametrize("wordsize,mismatches,block_env", [(8, 1, "1"), (8, 1, "2"), (10, 2, "1"), (6, 1, "3"), (11, 1, "1")])
def test_block_tables_equal_one_table(tmp_path, monkeypatch, wordsize, mismatches, block_env):
    """Searches that allow mismatches may be keyed on seed + one of N + 1 blocks behind it (mpcr_ctx_set_seed_blocks):
    the block tables and the table of the records that cannot be keyed that way, scanned separately and merged, must
    give the one-table result (== the oracle) -- every site once, whichever blocks its mismatches fall into."""
    from merpcr_b200 import FASTARecord, MerPCR
    contigs, text, expected = synth.block_table_case(700 + wordsize, wordsize, mismatches)
    stsf = tmp_path / "s.sts"
    stsf.write_bytes(text)
    params = dict(wordsize=wordsize, margin=30, mismatches=mismatches)
    recs = [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)]
    want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
    assert len(want) > 40
    for flag, parts in (("0", "1"), (block_env, "1"), (block_env, "2")):
        monkeypatch.setenv("MPCR_SEED_BLOCKS", flag)
        monkeypatch.setenv("MPCR_SEED_PARTS", parts)
        eng = MerPCR(**params)
        assert eng.load_sts_file(str(stsf))
        assert len(eng._ctx_exts) == (int(parts) * (mismatches + 1) if flag != "0" else 0)
        if flag != "0":    # blockable records went to the block tables, the others stayed
            items = [int(eng._be.lib.mpcr_table_items(c)) for c in eng._all_ctxs()]
            assert items[0] > 0 and all(x > 0 for x in items[1:]) and len(set(items[1::int(parts)])) == 1
        got = parity.engine_hits(eng, recs)
        assert np.array_equal(got, want), (flag, parts)
        eng.close()
    # two shards of the blocked search merge to the whole
    monkeypatch.setenv("MPCR_SEED_BLOCKS", block_env)
    monkeypatch.setenv("MPCR_SEED_PARTS", "1")
    pieces = []
    for rank in range(2):
        eng = MerPCR(**params, shard=(rank, 2))
        assert eng.load_sts_file(str(stsf))
        pieces.append(eng.search_hits(recs))
        eng.close()
    eng = MerPCR(**params)
    assert eng.load_sts_file(str(stsf))
    whole = eng.search_hits(recs)
    # the C ABI refuses settings that would lose sites
    lib, chk = eng._be.lib, eng._be.check
    with pytest.raises(ValueError):
        chk(lib.mpcr_ctx_set_seed_blocks(eng._ctx, 2, mismatches, 1))          # needs more blocks than mismatches
    with pytest.raises(ValueError):
        chk(lib.mpcr_ctx_set_seed_blocks(eng._ctx, 9, mismatches + 1, 1))      # seed + blocks beyond 16 letters
    eng.close()
    merged = np.concatenate(pieces)
    order = np.lexsort((merged["rank"], merged["rec"], merged["hash_off"], merged["pos1"], merged["contig"]))
    assert np.array_equal(merged[order], whole)
    eng = MerPCR(wordsize=8, mismatches=1, iupac_mode=1)
    with pytest.raises(ValueError):
        eng._be.check(eng._be.lib.mpcr_ctx_set_seed_blocks(eng._ctx, 4, 2, 1))   # IUPAC compares are not letter identity
    eng.close()


def test_fuzz_goldens_with_block_tables(monkeypatch):
    monkeypatch.setenv("MPCR_SEED_BLOCKS", "1")
    from merpcr_b200 import MerPCR
    n = 0
    for c in goldens.fuzz_cases():
        if c["params"].get("mismatches", 0) >= 1 and not c["params"].get("iupac_mode", 0):
            parity.check_fuzz_case(c, MerPCR)
            n += 1
    assert n > 20


def test_fuzz_goldens_with_seed_extension(monkeypatch):
    monkeypatch.setenv("MPCR_SEED_EXTENSION", "1")
    from merpcr_b200 import MerPCR
    for c in goldens.fuzz_cases():
        if c["params"].get("mismatches", 0) == 0 and not c["params"].get("iupac_mode", 0):
            parity.check_fuzz_case(c, MerPCR)


def _weird_sts_text(seed: int) -> bytes:
    rng = np.random.default_rng(seed)
    sizes = ["100", "0", "007", " 12", "12 ", "+5", "-5", "5-", "1_000", "100-200", "100-200-300", "a-b", "-", "",
             "999999999", "12345678901234567890", "50-60 ", "1e3", "٣", "100--200", "3-2", "0-0"]
    nl = ["\n", "\r\n", "\r"]
    letters = list("ACGTacgtNRYnU-*")
    w = np.array([20, 20, 20, 20, 3, 3, 3, 3, 1, 1, 1, 1, 1, .5, .5])
    prim = lambda k: "".join(rng.choice(letters, size=k, p=w / w.sum()))
    out = []
    for i in range(300):
        kind = rng.integers(0, 20)
        if kind == 0:
            out.append("# comment\tx\ty\tz")
        elif kind == 1:
            out.append("   ")
        elif kind == 2:
            out.append(f"\tSTS{i}\t{prim(20)}\t{prim(20)}\t100\talias")           # leading tab is stripped away
        else:
            size = sizes[int(rng.integers(0, len(sizes)))]
            if not size.isascii():
                size = "77"
            fields = [f"id {i}", prim(int(rng.integers(5, 30))), prim(int(rng.integers(5, 30))), size]
            if rng.random() < 0.6 or size.strip() == "":
                fields.append(f"alias {i}  x")
            if rng.random() < 0.2:
                fields += ["extra", "more"]
            if rng.random() < 0.1:
                fields[-1] += "\t"                                              # trailing tab
            out.append(("  " if rng.random() < 0.1 else "") + "\t".join(fields))
        out[-1] += nl[int(rng.integers(0, 3))]
    return "".join(out).encode("ascii")


@pytest.mark.parametrize("seed", range(6))
def test_native_sts_parser_equals_python_loop(tmp_path, seed):
    """mpcr_sts_parse + the vectorised size rules == the reference's line loop kept in _parse_sts_python."""
    p = tmp_path / "w.sts"
    p.write_bytes(_weird_sts_text(seed))
    for W in (8, 11):
        eng = _engine(wordsize=W)
        nat = eng._parse_sts_native(np.fromfile(str(p), dtype=np.uint8))
        py = eng._parse_sts_python(str(p))
        assert nat is not None and nat is not False and py is not False
        a, b = nat[0], py[0]
        assert (nat[1], nat[2]) == (py[1], py[2])
        assert a.n == b.n and [int(x) for x in a.sizes] == [int(x) for x in b.sizes]
        assert [int(x) for x in a.line_nos] == b.line_nos
        assert (a.ids, a.aliases, a.p1s, a.p2s) == (b.ids, b.aliases, b.p1s, b.p2s)


def test_native_sts_parser_malformed_and_non_ascii(tmp_path):
    eng = _engine()
    bad = tmp_path / "bad.sts"
    bad.write_bytes(b"ok\tACGTACGTACGTA\tACGTACGTACGTA\t100\nbroken line\tonly two\n")
    assert eng._parse_sts_native(np.fromfile(str(bad), dtype=np.uint8)) is False
    assert eng.load_sts_file(str(bad)) is False
    na = tmp_path / "na.sts"
    na.write_bytes("id\tACGTACGTACGTA\tACGTACGTACGTA\t100\tnaïve alias\n".encode("utf-8"))
    assert eng._parse_sts_native(np.fromfile(str(na), dtype=np.uint8)) is None
    assert eng.load_sts_file(str(na)) and eng.sts_records[0].alias == "naïve alias"


def _true_strands_check(MerPCRcls, record_factory):
    """true_strands=True is not reference behaviour, but it is pinned to the reference through two identities:
    its (+) hits are the reference's (+) hits on the STS file with primer2 reverse-complemented, and its (-) hits are
    the reference's (-) hits on the original file.  Planted, biologically normal amplicons must come out as (+)."""
    import tempfile
    rng = synth.Rng(8080)
    contigs = [rng.dna(n) for n in (50000, 30000)]
    sts = synth.make_sts_set(8081, 60, 18, 25, 100, 600)
    # plant forward amplicons p1 ... rc(p2) by planting the reference's "+" form of the transformed table
    sts_rc = dict(sts)
    sts_rc["p2"] = sts["p2"].copy()
    for i in range(len(sts["l2"])):
        k = int(sts["l2"][i])
        sts_rc["p2"][i, :k] = synth.revcomp_bytes(sts["p2"][i, :k])
    planted = synth.plant_amplicons(8082, contigs, sts_rc, 50)
    text, text_rc = synth.sts_lines(sts), synth.sts_lines(sts_rc)
    params = dict(wordsize=11, margin=50, mismatches=1)
    with tempfile.NamedTemporaryFile("wb", suffix=".sts", delete=False) as f:
        f.write(text)
    eng = MerPCRcls(**params, true_strands=True)
    assert eng.load_sts_file(f.name)
    os.unlink(f.name)
    got = parity.engine_hits(eng, record_factory(contigs))
    seqs = [c.tobytes() for c in contigs]
    ref_plain = parity.oracle_hits(params, text.decode(), seqs)
    ref_rc = parity.oracle_hits(params, text_rc.decode(), seqs)
    as_set = lambda a, strand: {tuple(r) for r in a.tolist() if r[4] == strand}
    assert as_set(got, 0) == as_set(ref_rc, 0) and as_set(got, 1) == as_set(ref_plain, 1)
    plus = {(r[0], r[1], r[2]) for r in got.tolist() if r[4] == 0}
    fwd = [(ci, a, b) for ci, a, b, _, strand in planted if strand == "+"]
    assert len(fwd) > 10 and all(x in plus for x in fwd)
    rec = next(r for r in eng.sts_records if r.direct == "+")
    assert rec.primer2 == eng._reverse_complement(sts["p2"][0, : sts["l2"][0]].tobytes().decode())
    eng.close()


def test_true_strands_option():
    from merpcr_b200 import FASTARecord, MerPCR
    _true_strands_check(MerPCR, lambda cs: [FASTARecord(f">c{i}", c) for i, c in enumerate(cs)])


def test_fullsize_oracle_comparison_helpers(tmp_path):
    """tests/fullsize.py's whole-genome and slice comparisons (what the GPU tier runs at BASELINE sizes), here on a
    small genome through the emulated backend -- including that they DO notice a missing or reordered hit."""
    import torch
    import fullsize
    from merpcr_b200 import FASTARecord, MerPCR
    rng = synth.Rng(515)
    lengths = [180_000, 40_000, 90_001]
    contigs = [rng.dna(n) for n in lengths]
    sts = synth.make_sts_set(516, 400, 18, 25, 100, 600)
    planted = synth.plant_amplicons(517, contigs, sts, 50, sub_mode="cfg3")
    text = synth.sts_lines(sts)
    sp = tmp_path / "f.sts"
    sp.write_bytes(text)
    params = dict(wordsize=11, margin=50, mismatches=1, three_prime_match=1)
    eng = MerPCR(**params)
    assert eng.load_sts_file(str(sp))
    hits = eng.search_hits([FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)])
    assert len(hits) >= len(planted) > 100
    tens = [torch.from_numpy(c) for c in contigs]
    safe = int(sts["size"].max()) + 40 + 50 + 64
    whole = fullsize.compare_with_oracle(eng, hits, params, text, tens, lengths, safe, "whole")
    assert whole["oracle_bit_exact"] and whole["oracle_hits"] == len(hits) and whole["oracle_bp"] == sum(lengths)
    jobs = fullsize.slice_jobs(lengths, 6, 30_000, safe)
    assert jobs[0][:2] == (0, 0) and jobs[-1][0] == 2 and jobs[-1][2] == lengths[2] and jobs[-1][3] == 30_000
    assert all(cut == 30_000 - safe for _, _, b, cut in jobs[1:-1])
    sl = fullsize.compare_with_oracle(eng, hits, params, text, tens, lengths, safe, ("slices", 6, 30_000))
    assert sl["oracle_bit_exact"] and sl["oracle_hits"] > 10
    # a dropped hit and two swapped neighbours are both caught
    assert not fullsize.compare_with_oracle(eng, np.delete(hits, len(hits) // 2), params, text, tens, lengths, safe,
                                            "whole")["oracle_bit_exact"]
    swapped = hits.copy()
    swapped[[3, 4]] = swapped[[4, 3]]
    assert not fullsize.compare_with_oracle(eng, swapped, params, text, tens, lengths, safe, "whole")["oracle_bit_exact"]
    k = int(np.flatnonzero((hits["contig"] == 0) & (hits["pos2"] < 30_000 - safe))[0])
    assert not fullsize.compare_with_oracle(eng, np.delete(hits, k), params, text, tens, lengths, safe,
                                            ("slices", 6, 30_000))["oracle_bit_exact"]
    eng.close()


@pytest.mark.parametrize("host_pack", ["1", "0"])
def test_nibble_ingest_and_ascii_ingest_agree_with_the_oracle(tmp_path, monkeypatch, host_pack):
    """Host-resident sequence reaches the planes either as host-packed nibbles (+ mpcr_derive_planes) or as ASCII
    (mpcr_pack_sequence); a piece holding 'U' outside IUPAC mode (it hashes like T but equals nothing, so its planes do
    not follow from its nibbles) must fall back to ASCII by itself.  Both give the oracle's hits."""
    from merpcr_b200 import FASTARecord, MerPCR
    monkeypatch.setenv("MPCR_HOST_PACK", host_pack)
    rng = synth.Rng(808)
    contigs = [rng.dna(70_001), rng.dna(33_333), rng.dna(64 * 700)]
    sts = synth.make_sts_set(809, 150, 18, 25, 100, 500)
    synth.plant_amplicons(810, contigs, sts, 50, sub_mode="cfg3")
    contigs[0][1000:1040] = np.frombuffer(b"NNNNNRYKMSWBDHVXnacgtryNNNNNNNNNNNNNNNNN"[:40], dtype=np.uint8)
    contigs[1][5:9] = np.frombuffer(b"UuUT", dtype=np.uint8)          # irregular under -I 0
    contigs[1][20_000] = ord("U")
    text = synth.sts_lines(sts)
    sp = tmp_path / "n.sts"
    sp.write_bytes(text)
    for params in (dict(wordsize=11, margin=50, mismatches=1), dict(wordsize=11, margin=50, mismatches=2, iupac_mode=1)):
        eng = MerPCR(**params)
        assert eng.host_pack == (host_pack == "1")
        assert eng.load_sts_file(str(sp))
        got = parity.engine_hits(eng, [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)])
        want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
        assert np.array_equal(got, want) and len(want) > 50
        eng.close()


def _sampling_case(tmp_path, tag):
    """Primers of 18..30 letters (most can be sampled, some cannot: short ones, an ambiguity inside the windows, a late
    hash offset), planted amplicons, a contig shorter than a window, tandem repeats that hit many probed positions."""
    rng = synth.Rng(1411)
    contigs = [rng.dna(n) for n in (90000, 30001, 14)]
    unit = np.frombuffer(b"ACGGTCATTGCAGTTTGACCGGTATCAGCATGCATGAACC", dtype=np.uint8)
    contigs[1][5000:5000 + 40 * 50] = np.tile(unit, 50)
    sts = synth.make_sts_set(1412, 3000, 12, 30, 60, 400)
    sts["p1"][::5, 14] = ord("N")          # an ambiguity inside the sampled windows: not sampleable
    sts["p1"][1::9, 2] = ord("R")          # ... or in front of them: a later hash offset
    expected = synth.plant_amplicons(1413, contigs[:2], sts, 30, plant_count=150)
    text = synth.sts_lines(sts) + b"REP\tACGGTCATTGCAGTTTGACC\tGGTATCAGCATGCATGAACC\t80\trepeat\n"
    stsf = tmp_path / f"{tag}.sts"
    stsf.write_bytes(text)
    return contigs, text, str(stsf), expected


def test_position_sampling_equals_the_unsampled_search(tmp_path, monkeypatch):
    """mpcr_ctx_set_sampling: the sampled table (every S-th position probed, S windows per record) + the tables of the
    records that cannot be sampled == the plain search == the oracle, for several strides, alone and on top of the
    seed extension, whole and cut into shards (a site belongs to the shard that holds its PROBED position)."""
    from merpcr_b200 import FASTARecord, MerPCR
    contigs, text, stsf, expected = _sampling_case(tmp_path, "samp")
    params = dict(wordsize=8, margin=30, mismatches=0)
    recs = [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)]
    want = parity.oracle_hits(params, text.decode(), [c.tobytes() for c in contigs])
    assert len(want) >= len(expected) > 50
    for ext, stride in (("0", "0"), ("0", "3"), ("1", "3"), ("1", "2"), ("0", "5"), ("1", "7")):
        monkeypatch.setenv("MPCR_SEED_EXTENSION", ext)
        monkeypatch.setenv("MPCR_SAMPLING", stride)
        eng = MerPCR(**params)
        assert eng.load_sts_file(stsf)
        assert (eng._ctx_samp is not None) == (stride != "0")
        if stride != "0":
            items = int(eng._be.lib.mpcr_table_items(eng._ctx_samp))
            assert items > 0 and items % int(stride) == 0
        got = parity.engine_hits(eng, recs)
        assert np.array_equal(got, want), (ext, stride)
        eng.close()
    # shards: the union of the ranks' hits, merged by the order key, is the whole
    monkeypatch.setenv("MPCR_SEED_EXTENSION", "1")
    monkeypatch.setenv("MPCR_SAMPLING", "3")
    parts = []
    for rank in range(3):
        eng = MerPCR(**params, shard=(rank, 3))
        assert eng.load_sts_file(stsf)
        parts.append(eng.search_hits(recs))
        eng.close()
    eng = MerPCR(**params)
    assert eng.load_sts_file(stsf)
    whole = eng.search_hits(recs)
    eng.close()
    merged = np.concatenate(parts)
    order = np.lexsort((merged["rank"], merged["rec"], merged["hash_off"], merged["pos1"], merged["contig"]))
    assert np.array_equal(merged[order], whole) and all(len(p) for p in parts[:2])
    # a search that allows mismatches cannot be sampled
    eng = MerPCR(wordsize=8, mismatches=1)
    with pytest.raises(ValueError):
        eng._be.check(eng._be.lib.mpcr_ctx_set_sampling(eng._ctx, 16, 3, 1))
    eng.close()


def test_fuzz_goldens_with_position_sampling(monkeypatch):
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


def _two_in_flight_check(MerPCR, make_records, tmp_path):
    """scan_device_async / scan_finish (two steps in flight, one hit buffer and count per slot) == scan_device, including a
    hit list that outgrows the slot's first buffer and a list that piles up in one place (the short-list sort gives up
    and scan_finish completes it)."""
    unit = "ACGGTCATTGCAGT" + "TTGACCGGTATCAG" + "CATGCATGAACC"
    seq = np.frombuffer(("G" * 300 + unit * 5000 + "C" * 300).encode(), dtype=np.uint8).copy()
    sts_text = "".join(f"R{i}\tACGGTCATTGCAGT\tTTGACCGGTATCAG\t{68 + (i % 3)}\trep\n" for i in range(4)).encode()
    p = tmp_path / "two.sts"
    p.write_bytes(sts_text)
    eng = MerPCR(wordsize=8, margin=45, mismatches=0)
    assert eng.load_sts_file(str(p))
    recs = make_records([seq, seq[:50_000].copy()])
    layout = eng.make_layout([len(r) for r in recs])
    sh = eng.upload(layout, [r.sequence_bytes for r in recs])
    first = eng.scan_device_async(layout, sh, slot=0)          # the slots start with room for 65 536 hits ...
    h0, n0 = eng.scan_finish(layout, sh, first)                  # ... so this one is scanned again with more
    got0 = eng._hits_to_host(h0, n0)
    hits, n = eng.scan_device(layout, sh)
    want = eng._hits_to_host(hits, n)
    assert n == n0 > 65536 and np.array_equal(got0, want)
    handles = [eng.scan_device_async(layout, sh, slot=0), eng.scan_device_async(layout, sh, slot=1)]
    for k in range(4):
        h, m = eng.scan_finish(layout, sh, handles[k & 1])
        assert m == n and np.array_equal(eng._hits_to_host(h, m), want), k
        handles[k & 1] = eng.scan_device_async(layout, sh, slot=k & 1)
    for hd in handles:
        h, m = eng.scan_finish(layout, sh, hd)
        assert m == n and np.array_equal(eng._hits_to_host(h, m), want)
    eng.close()


def test_two_steps_in_flight(tmp_path):
    from merpcr_b200 import FASTARecord, MerPCR
    _two_in_flight_check(MerPCR, lambda seqs: [FASTARecord(f">c{i}", s) for i, s in enumerate(seqs)], tmp_path)


def test_linear_filter_map_properties():
    """mpcr_core.cuh's linear filter map (11-letter keys): the float fma must give an in-range, monotone word index
    for every key, and the keys of one word must get (all but a few) distinct bit pairs -- what keeps its false-positive
    rate low: of 128 consecutive keys six pairs share a mask (bits [0,5) and [2,7) swapped)."""
    import ctypes
    lib = ctypes.CDLL(emul.build())
    lib.emul_filter_linear.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32),
                                       ctypes.POINTER(ctypes.c_uint32)]
    w, m = ctypes.c_uint32(), ctypes.c_uint32()

    def probe(n_words, key):
        assert lib.emul_filter_linear(n_words, key, ctypes.byref(w), ctypes.byref(m)) == 1
        return w.value, m.value

    rng = np.random.default_rng(5)
    for n_words in (64, 1000, 39616, 40128, 57000):
        assert probe(n_words, 0)[0] == 0 and probe(n_words, (1 << 22) - 1)[0] == n_words - 1
        keys = np.unique(np.concatenate([rng.integers(0, 1 << 22, 4000), np.arange(0, 3000), np.arange((1 << 22) - 3000, 1 << 22)]))
        res = [probe(n_words, int(k)) for k in keys]
        words = np.array([r[0] for r in res])
        assert words.min() >= 0 and words.max() < n_words
        assert np.all(np.diff(words) >= 0)                       # monotone in the key
        assert all(bin(r[1]).count("1") in (1, 2) for r in res)
        # consecutive keys that share a word: pairwise distinct masks (here on the dense runs at both ends)
        dense = np.arange(0, 3000)
        by_word = {}
        for k in dense:
            wd, mk = probe(n_words, int(k))
            by_word.setdefault(wd, []).append(mk)
        if n_words >= 1 << 15:                                   # <= 128 keys per word: bits [0,5) + [2,7) separate them
            assert all(len(v) - len(set(v)) <= 6 for v in by_word.values())
            assert sum(len(v) - len(set(v)) for v in by_word.values()) <= 0.06 * len(dense)
    assert lib.emul_filter_linear(1, 0, ctypes.byref(w), ctypes.byref(m)) == 0      # refused: the caller keeps the other map


def test_fasta_keep_flags_four_bytes_at_a_time():
    """mpcr_core.cuh: the bit-sliced keep test of the device-side FASTA ingest against io/fasta.py:60's letter set, for
    every byte value in every byte position next to arbitrary neighbours; and the byte-equality trigger."""
    import ctypes
    lib = ctypes.CDLL(emul.build())
    lib.emul_fasta_keep_flags4.argtypes = [ctypes.c_uint32]
    lib.emul_fasta_keep_flags4.restype = ctypes.c_uint32
    lib.emul_bytes_equal_trigger4.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
    lib.emul_bytes_equal_trigger4.restype = ctypes.c_uint32
    keep = set(b"ACGTBDHKMNRSVWXYacgtbdhkmnrsvwxy")
    rng = np.random.default_rng(11)
    for v in range(256):
        for pos in range(4):
            for other in (0x00000000, 0xFFFFFFFF, 0x41414141, int(rng.integers(0, 1 << 32))):
                w = (other & ~(0xFF << (8 * pos)) | (v << (8 * pos))) & 0xFFFFFFFF
                flags = lib.emul_fasta_keep_flags4(w)
                assert flags & ~0x01010101 == 0
                assert ((flags >> (8 * pos)) & 1) == (1 if v in keep else 0), (v, pos, hex(w), hex(flags))
    for _ in range(20000):
        w = int(rng.integers(0, 1 << 32))
        if rng.random() < 0.5:
            w = (w & ~(0xFF << (8 * int(rng.integers(0, 4)))) | (0x3E << (8 * int(rng.integers(0, 4))))) & 0xFFFFFFFF
        has = any(((w >> (8 * k)) & 0xFF) == 0x3E for k in range(4))
        trig = lib.emul_bytes_equal_trigger4(w, 0x3E3E3E3E)
        assert (trig != 0) == has
        if has:   # the lowest flagged byte is a true match
            low = min(k for k in range(4) if (trig >> (8 * k + 7)) & 1)
            assert ((w >> (8 * low)) & 0xFF) == 0x3E


def test_rolling_mate_precheck_equals_the_per_position_check():
    """mpcr_core.cuh: mate_precheck32 (the verifier's rolling 8-base check over up to 32 consecutive mate positions) must
    be exactly compare_view's own first check at every position, for every alignment of the block, both compare modes."""
    import ctypes
    lib = ctypes.CDLL(emul.build())
    lib.emul_mate_precheck_diff.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.emul_mate_precheck_diff.restype = ctypes.c_uint32
    rng = np.random.default_rng(21)
    n_bases = 4096
    # plane4: mostly A/C/G/T nibbles (1, 2, 4, 8), some N (15), some IUPAC masks, some zero (X / foreign)
    nib = rng.choice(np.array([1, 2, 4, 8, 15, 3, 5, 0], dtype=np.uint8), size=n_bases + 128, p=[.23, .23, .23, .23, .03, .02, .02, .01])
    plane = np.zeros((n_bases + 128) // 16, dtype=np.uint64)
    for i, v in enumerate(nib):
        plane[i // 16] |= np.uint64(int(v)) << np.uint64(4 * (i % 16))
    checked = 0
    for trial in range(300):
        length = int(rng.integers(8, 33))
        pos = int(rng.integers(0, n_bases - 200))
        # primer = a genome stretch with a few substitutions / a degenerate letter, so that some positions pass
        pn = nib[pos: pos + length].copy()
        pn[pn == 0] = 1
        for _ in range(int(rng.integers(0, 3))):
            pn[int(rng.integers(0, length))] = int(rng.choice([1, 2, 4, 8, 15, 5]))
        nw = (length + 15) // 16
        words = np.zeros(2 * nw, dtype=np.uint64)       # nibble words, then aux words (no flags)
        for i, v in enumerate(pn):
            words[i // 16] |= np.uint64(int(v)) << np.uint64(4 * (i % 16))
        for gb in (pos - int(rng.integers(0, 31)), pos - 17, pos):
            gb = max(gb, 0)
            for m in (32, int(rng.integers(1, 32))):
                for N, X, iupac in ((0, 0, 0), (1, 1, 0), (2, 2, 1), (0, 1, 1)):
                    d = lib.emul_mate_precheck_diff(plane.ctypes.data, gb, m, words.ctypes.data, length, N, X, iupac)
                    assert d == 0, (trial, gb, m, N, X, iupac)
                    checked += 1
    assert checked > 5000
