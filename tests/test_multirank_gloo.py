"""CPU tier: the N > 1 path as the bench launches it -- one process per rank, torch.distributed (gloo here, NCCL on
the GPU box) only for the barrier / gather around the scan, no collective on the data path.
Each rank scans its bp-balanced shard through tests/host_emul; rank 0 merges and checks against the oracle."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = dict(wordsize=11, margin=50, mismatches=1)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_two_ranks_over_gloo_equal_the_oracle(tmp_path, world):
    import torch.multiprocessing as mp
    import emul
    import parity
    import synth
    emul.build()
    rng = synth.Rng(501)
    contigs = [rng.dna(n) for n in (120000, 64000, 900, 75001)]
    sts = synth.make_sts_set(502, 150, 18, 25, 100, 700)
    expected = synth.plant_amplicons(503, contigs, sts, 50, sub_mode="cfg3")
    sts_path = str(tmp_path / "s.sts")
    with open(sts_path, "wb") as f:
        f.write(synth.sts_lines(sts))
    # the ranks regenerate the same contigs from the seed; planting must be reproduced there too
    np.save(str(tmp_path / "contigs.npy"), np.concatenate(contigs))
    out_path = str(tmp_path / "merged.npy")
    port = _free_port()
    mp.spawn(_rank_main_planted, args=(world, port, sts_path, out_path, str(tmp_path / "contigs.npy"),
                                       [len(c) for c in contigs]), nprocs=world, join=True)
    merged = np.load(out_path)
    want = parity.oracle_hits(PARAMS, synth.sts_lines(sts).decode(), [c.tobytes() for c in contigs])
    assert len(merged) == len(want) >= len(expected) > 50
    assert np.array_equal(merged["contig"], want[:, 0]) and np.array_equal(merged["pos1"], want[:, 1])
    assert np.array_equal(merged["pos2"], want[:, 2]) and np.array_equal(merged["rec"] & 1, want[:, 4])
    found = set(zip(merged["contig"].tolist(), merged["pos1"].tolist(), merged["pos2"].tolist()))
    assert all((ci, a, b) in found for ci, a, b, _, _ in expected)


def _rank_main_planted(rank, world, port, sts_path, out_path, contigs_path, lengths):
    for p in (os.path.dirname(HERE), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import emul
    emul.inject()
    from merpcr_b200 import FASTARecord, MerPCR
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        flat = np.load(contigs_path)
        contigs, a = [], 0
        for n in lengths:
            contigs.append(flat[a:a + n])
            a += n
        recs = [FASTARecord(f">c{i}", c) for i, c in enumerate(contigs)]
        eng = MerPCR(**PARAMS, shard=(rank, world))
        assert eng.load_sts_file(sts_path)
        dist.barrier()
        mine = eng.search_hits(recs)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)          # the "final hit gather" -- outside the scan path
        if rank == 0:
            merged = np.concatenate(gathered)
            order = np.lexsort((merged["rank"], merged["rec"], merged["hash_off"], merged["pos1"], merged["contig"]))
            np.save(out_path, merged[order])
        dist.barrier()
    finally:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# the product's own multi-GPU path (merpcr_b200/multi.py): `search` and the CLI as ranks of a process group --
# every rank scans its shard, rank 0 gathers, merges and writes exactly the single-process output text
# ---------------------------------------------------------------------------------------------------------
def _fasta_and_sts(tmp_path, seed):
    import synth
    rng = synth.Rng(seed)
    contigs = [rng.dna(n) for n in (90000, 300, 51000, 12, 70001)]
    sts = synth.make_sts_set(seed + 1, 120, 18, 25, 100, 700)
    synth.plant_amplicons(seed + 2, contigs, sts, 50, sub_mode="cfg3")
    sts_path, fa_path = str(tmp_path / "m.sts"), str(tmp_path / "m.fa")
    with open(sts_path, "wb") as f:
        f.write(synth.sts_lines(sts))
    with open(fa_path, "wb") as f:
        for i, c in enumerate(contigs):
            f.write(b">ctg%d some text\n" % i)
            for a in range(0, len(c), 70):
                f.write(c[a:a + 70].tobytes() + b"\n")
    return sts_path, fa_path


def _single_process_text(sts_path, fa_path, out_path):
    import emul
    emul.inject()
    from merpcr_b200 import MerPCR
    eng = MerPCR(**PARAMS)
    assert eng.load_sts_file(sts_path)
    n = eng.search(eng.load_fasta_file(fa_path), out_path)
    return n, open(out_path).read()


def _rank_main_search(rank, world, port, sts_path, fa_path, out_path, totals_path):
    for p in (os.path.dirname(HERE), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import emul
    emul.inject()
    from merpcr_b200 import MerPCR
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        eng = MerPCR(**PARAMS, shard=(rank, world))
        assert eng.load_sts_file(sts_path)
        n = eng.search(eng.load_fasta_file(fa_path), out_path)
        assert n == eng.total_hits
        with open(f"{totals_path}.{rank}", "w") as f:
            f.write(str(n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_search_gathers_hits_on_rank0(tmp_path, world):
    import torch.multiprocessing as mp
    sts_path, fa_path = _fasta_and_sts(tmp_path, 611)
    n1, want = _single_process_text(sts_path, fa_path, str(tmp_path / "one.txt"))
    assert n1 > 50
    out_path, totals = str(tmp_path / "multi.txt"), str(tmp_path / "total")
    mp.spawn(_rank_main_search, args=(world, _free_port(), sts_path, fa_path, out_path, totals), nprocs=world, join=True)
    assert open(out_path).read() == want                       # rank 0 wrote the merged list, byte for byte
    assert [int(open(f"{totals}.{r}").read()) for r in range(world)] == [n1] * world   # every rank knows the total


def _rank_main_cli(rank, world, port, argv, rc_path):
    for p in (os.path.dirname(HERE), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import emul
    emul.inject()
    from merpcr_b200 import cli
    sys.argv = ["merpcr"] + argv
    rc = cli.main()
    with open(f"{rc_path}.{rank}", "w") as f:
        f.write(str(rc))


def test_cli_as_ranks_of_a_launch(tmp_path):
    """What `python -m merpcr_b200 --gpus 2 ...` runs in each of its worker processes (torchrun environment)."""
    import torch.multiprocessing as mp
    sts_path, fa_path = _fasta_and_sts(tmp_path, 733)
    n1, want = _single_process_text(sts_path, fa_path, str(tmp_path / "one.txt"))
    out_path, rc_path = str(tmp_path / "cli.txt"), str(tmp_path / "rc")
    argv = ["-M", "50", "N=1", "-W", "11", "--gpus", "2", "-O", out_path, sts_path, fa_path]
    mp.spawn(_rank_main_cli, args=(2, _free_port(), argv, rc_path), nprocs=2, join=True)
    assert [open(f"{rc_path}.{r}").read() for r in range(2)] == ["0", "0"]
    assert open(out_path).read() == want


def test_merge_hits_order_key_and_empty_ranks():
    """merge_hits: ranks without hits, a single rank, and boundaries that interleave (pos1 = seed - hash_offset)."""
    from merpcr_b200 import _capi, multi
    dt = _capi.HIT_DTYPE

    def mk(rows):
        a = np.zeros(len(rows), dtype=dt)
        for i, (c, p1, ho, rec, rank) in enumerate(rows):
            a[i]["contig"], a[i]["pos1"], a[i]["pos2"], a[i]["rec"], a[i]["rank"], a[i]["hash_off"] = c, p1, p1 + 100, rec, rank, ho
        return a
    r0 = mk([(0, 10, 0, 4, 0), (0, 995, 7, 9, 0), (0, 1000, 0, 2, 1)])
    r1 = mk([(0, 990, 12, 5, 0), (0, 1000, 0, 2, 0), (1, 3, 0, 1, 0)])     # starts before rank 0's last hits end
    empty = np.zeros(0, dtype=dt)
    m = multi.merge_hits([r0, empty, r1, None])
    key = list(zip(m["contig"].tolist(), m["pos1"].tolist(), m["hash_off"].tolist(), m["rec"].tolist(), m["rank"].tolist()))
    assert key == sorted(key) and len(m) == 6
    assert np.array_equal(multi.merge_hits([r0]), r0)
    assert len(multi.merge_hits([empty, empty])) == 0 and len(multi.merge_hits([])) == 0


# ---------------------------------------------------------------------------------------------------------
# rank-local FASTA ingest: every rank reads only its own byte range of the file (+ margins), one small exchange
# gives everybody the record table; the merged output must still be the single-process text, byte for byte
# ---------------------------------------------------------------------------------------------------------
def _nasty_fasta(tmp_path, seed):
    """Records of very different sizes, CRLF / CR / LF line ends, blank lines, indented headers, an empty record, sequence
    in front of the first header, lower case, letters that are filtered out -- so that the byte-range cuts of 2..5 ranks
    fall into headers, line ends and tiny records."""
    import synth
    rng = synth.Rng(seed)
    contigs = [rng.dna(n) for n in (70000, 17, 41000, 0, 300, 52001, 9000)]
    sts = synth.make_sts_set(seed + 1, 150, 18, 25, 100, 600)
    synth.plant_amplicons(seed + 2, [c for c in contigs if len(c) > 1000], sts, 50, sub_mode="cfg3")
    sts_path, fa_path = str(tmp_path / "n.sts"), str(tmp_path / "n.fa")
    with open(sts_path, "wb") as f:
        f.write(synth.sts_lines(sts))
    ends = [b"\n", b"\r\n", b"\r"]
    with open(fa_path, "wb") as f:
        f.write(b"ACGTACGTNNNN sequence in front of the first header is dropped\n\n")
        for i, c in enumerate(contigs):
            f.write(b"  " * (i % 2) + b">rec%d a header line with some text in it %s" % (i, b"x" * (40 * i)) + ends[i % 3])
            body = c.copy()
            if len(body) > 100:
                body[50:60] |= 0x20                                    # lower case is kept as it is
            width = 60 + 7 * i
            for a in range(0, len(body), width):
                f.write(body[a:a + width].tobytes() + (b" 12*-" if a % (width * 5) == 0 else b"") + ends[(a // width + i) % 3])
                if a % (width * 11) == 0:
                    f.write(ends[i % 3])                               # blank lines
    return sts_path, fa_path


def _rank_main_local_ingest(rank, world, port, sts_path, fa_path, out_path, info_path, margins):
    for p in (os.path.dirname(HERE), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MPCR_DEVICE_INGEST_MIN_BYTES="1", MPCR_RANK_MARGIN_LEFT=str(margins[0]),
                      MPCR_RANK_MARGIN_RIGHT=str(margins[1]))
    import torch.distributed as dist
    import emul
    emul.inject()
    from merpcr_b200 import MerPCR
    from merpcr_b200.fasta import ShardedRecords
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        eng = MerPCR(**PARAMS, shard=(rank, world))
        assert eng.load_sts_file(sts_path)
        recs = eng.load_fasta_file(fa_path)
        sharded = isinstance(recs, ShardedRecords)
        held = sum(r.piece[0].numel() for r in recs if getattr(r, "piece", None) is not None) if sharded else -1
        n = eng.search(recs, out_path)
        with open(f"{info_path}.{rank}", "w") as f:
            f.write(f"{int(sharded)} {n} {held} {sum(len(r) for r in recs)} {len(recs)}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,margins,expect_sharded", [(2, (4096, 60000), True), (3, (4096, 60000), True),
                                                          (5, (4096, 60000), True), (3, (64, 60000), False),
                                                          (2, (4096, 2000), False)])
def test_rank_local_fasta_ingest(tmp_path, world, margins, expect_sharded):
    """Each rank ingests its own byte range (margins hold the halos); too small a margin makes every rank fall back to
    the whole file.  Either way rank 0 writes exactly the single-process text."""
    import torch.multiprocessing as mp
    sts_path, fa_path = _nasty_fasta(tmp_path, 911)
    n1, want = _single_process_text(sts_path, fa_path, str(tmp_path / "one.txt"))
    assert n1 > 50
    out_path, info = str(tmp_path / "multi.txt"), str(tmp_path / "info")
    mp.spawn(_rank_main_local_ingest, args=(world, _free_port(), sts_path, fa_path, out_path, info, margins),
             nprocs=world, join=True)
    assert open(out_path).read() == want
    rows = [open(f"{info}.{r}").read().split() for r in range(world)]
    assert [int(r[0]) for r in rows] == [int(expect_sharded)] * world
    assert [int(r[1]) for r in rows] == [n1] * world
    total = int(rows[0][3])
    assert all(int(r[3]) == total and int(r[4]) == 7 for r in rows)        # every rank knows every record and its length
    if expect_sharded:
        assert all(int(r[2]) < (1.0 / world + 0.42) * total for r in rows)   # ... but holds only its own stretch (+ margins)
