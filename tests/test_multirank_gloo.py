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
