"""Pins the CPU oracle (oracle/merpcr_oracle.c) to the reference: fixture golden line, unit known-answers and
the committed reference-generated fuzz / threaded vectors.  CPU only."""
import hashlib

import pytest

import goldens
import synth
from oracle.oracle import Oracle, load_fasta_text


def test_fixture_golden_line():
    # reference tests/test_comprehensive.py:65-95 -- defaults give exactly one hit
    o = Oracle()
    status, hits, text = o.run_files(goldens.FIXTURE_STS, goldens.FIXTURE_FA)
    assert (status, hits, text) == (0, 1, goldens.FIXTURE_LINE)
    assert o.num_records == 6 and o.max_pcr_size == 193          # test_comprehensive.py:42, SURVEY App. B
    assert [r["hash"] for r in o.records()] == [3638181, 3526114, 2555953, 3737721, 2062650, 476488]


@pytest.mark.parametrize("flags,expect_hit", [
    (dict(mismatches=1), True), (dict(mismatches=2), True), (dict(wordsize=3), True),
    (dict(wordsize=8, mismatches=2, margin=500, three_prime_match=0), True), (dict(iupac_mode=1), True),
    (dict(margin=8), True), (dict(margin=7), False), (dict(margin=5), False),
])
def test_fixture_flag_sweep(flags, expect_hit):
    # SURVEY Appendix B table (produced by the reference)
    _, hits, text = Oracle(**flags).run_files(goldens.FIXTURE_STS, goldens.FIXTURE_FA)
    assert text == (goldens.FIXTURE_LINE if expect_hit else "") and hits == int(expect_hit)


def test_unit_known_answers():
    # reference tests/test_engine_internals.py:26-62, tests/test_utils_comprehensive.py:173-181,27-54
    assert Oracle(wordsize=4).hash_value("ATCG") == (0, 54)
    assert Oracle(wordsize=8).hash_value("AAAAAAAA") == (0, 0)
    assert Oracle(wordsize=8).hash_value("TTTTTTTT") == (0, 65535)
    assert Oracle(wordsize=8).hash_value("NNNATCGATCGATCG")[0] == 3
    assert Oracle().hash_value("ACGUACGUACG") == Oracle().hash_value("ACGTACGTACG") == (0, 444102)
    assert Oracle(wordsize=8).hash_value("ACGT") == (-1, 0)
    o = Oracle()
    assert o.reverse_complement("ACGU-XZ") == "NXNACGT"
    assert o.reverse_complement("RWYS") == "SRWY" and o.reverse_complement("BDHV") == "BDHV"
    assert o.reverse_complement("acgt") == "acgt"


def test_compare_three_prime_protection():
    # reference tests/test_engine_internals.py:78-122 (X=2, N=1)
    o = Oracle(mismatches=1, three_prime_match=2)
    assert o.compare_seqs("ATCGATCG", "ATCGATCG", "+")
    assert o.compare_seqs("TTCGATCG", "ATCGATCG", "+")        # mismatch at the 5' end is allowed
    assert not o.compare_seqs("ATCGATCA", "ATCGATCG", "+")    # last base protected on '+'
    assert not o.compare_seqs("ATCGATTG", "ATCGATCG", "+")
    assert not o.compare_seqs("TTCGATCG", "ATCGATCG", "-")    # first base protected on '-'
    assert o.compare_seqs("ATCGATCA", "ATCGATCG", "-")
    assert not o.compare_seqs("ATCG", "ATCGA", "+")           # length mismatch


def test_compare_iupac():
    # reference tests/test_engine_internals.py:133-154, tests/test_comprehensive.py:132-146
    o = Oracle(iupac_mode=1)
    assert o.compare_seqs("A", "R", "+") and o.compare_seqs("G", "R", "+") and not o.compare_seqs("C", "R", "+")
    assert o.compare_seqs("N", "A", "+") and o.compare_seqs("T", "W", "+") and not o.compare_seqs("G", "W", "+")
    assert o.compare_seqs("X", "X", "+") and not o.compare_seqs("X", "A", "+") and not o.compare_seqs("X", "N", "+")
    p = Oracle(iupac_mode=0, three_prime_match=0)
    assert p.compare_seqs("N", "N", "+") and not p.compare_seqs("N", "A", "+")     # Q6


def test_fasta_filter_known_answer():
    # reference tests/test_io_modules.py:88-100
    recs = load_fasta_text(">seq1\nATCG123NNNN456ATCG\nWXYZ789GCTA\n")
    assert recs == [(">seq1", "seq1", "ATCGNNNNATCGWXYGCTA")]
    assert load_fasta_text("") == []
    assert load_fasta_text(">a\n>b\nAC\n") == [(">a", "a", ""), (">b", "b", "AC")]
    with pytest.raises(IndexError):
        load_fasta_text(">\nACGT\n")                          # Q12 bare header


def test_sts_loading_rules():
    # reference tests/test_io_modules.py:190-222, SURVEY A.2
    o = Oracle()
    assert o.load_sts_text("S1\tATCGATCGATCG\tGCTAGCTAGCTA\t150-250\n")
    assert o.record(0)["pcr_size"] == 200
    assert o.load_sts_text("S1\tATCG\tGCTA\t100\n") and o.num_records == 0      # short primers skipped
    assert not o.load_sts_text("") and not o.load_sts_text("a\tb\tc\n")
    assert o.load_sts_text("S\tATCGATCGATCG\tGCTAGCTAGCTA\t-100\n") and o.record(0)["pcr_size"] == 240


def _check_case(c):
    e = c["expect"]
    o = Oracle(**c["params"])
    ok = o.load_sts_text(c["sts_text"])
    assert ok == e["load_ok"], c["seed"]
    if not ok:
        return
    recs = [[r["id"], r["direct"], r["hash_offset"], r["pcr_size"], r["offset"], r["primer1"], r["primer2"],
             r["alias"]] for r in o.records()]
    assert recs == e["records"], c["seed"]
    assert o.max_pcr_size == e["max_pcr_size"]
    try:
        fa = load_fasta_text(c["fasta_text"])
    except IndexError:
        assert e["error"] == "fasta:IndexError"
        return
    assert e["error"] is None
    assert [[l, len(s), hashlib.sha256(s.encode()).hexdigest()[:16]] for _, l, s in fa] == e["fasta"], c["seed"]
    n, text = o.search([(l, s) for _, l, s in fa])
    assert (n, text) == (e["hits"], e["output"]), (c["seed"], c["params"])


def test_fuzz_goldens_bit_exact():
    cases = goldens.fuzz_cases()
    assert len(cases) >= 500
    for c in cases:
        _check_case(c)
    assert sum(c["expect"]["hits"] for c in cases) > 1000


def test_threaded_goldens_bit_exact():
    """The reference's multi-process path (engine.py:381-431), duplicates and all."""
    import hashlib as H
    for i, c in enumerate(goldens.threaded_cases()):
        genome = synth.dna_chunked(c["seed"], c["length"])
        sts = synth.make_sts_set(c["seed"] + 1000, c["n_sts"], 18, 25, 100, 600)
        synth.plant_amplicons(c["seed"] + 2000, [genome], sts, c["params"]["margin"], sub_mode="cfg3")
        sts_text = synth.sts_lines(sts).decode()
        fasta_text = ">big%d synthetic\n" % i + "\n".join(
            genome[j: j + 60].tobytes().decode() for j in range(0, c["length"], 60)) + "\n"
        assert H.sha256((sts_text + fasta_text).encode()).hexdigest() == c["sha"], "synthetic generator drifted"
        o = Oracle(**c["params"])
        assert o.load_sts_text(sts_text)
        fa = load_fasta_text(fasta_text)
        n, text = o.search([(l, s) for _, l, s in fa], threads=c["threads"])
        assert (n, text) == (c["expect"]["hits"], c["expect"]["output"]), i
        n1, text1 = o.search([(l, s) for _, l, s in fa], threads=1)
        assert set(text1.splitlines()) == set(text.splitlines())      # threads only ever add duplicates
