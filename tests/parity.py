"""Shared checkers: run a golden / oracle case through merpcr_b200.MerPCR (whatever backend is active)."""
import hashlib
import os
import tempfile

import numpy as np

from oracle.oracle import Oracle


def check_fuzz_case(c, MerPCR, **engine_kw):
    """One reference-generated golden case, end to end through the public API (files in, text out)."""
    ex = c["expect"]
    with tempfile.TemporaryDirectory() as d:
        sp, fp, op = (os.path.join(d, x) for x in ("in.sts", "in.fa", "out.txt"))
        with open(sp, "w", newline="") as f:
            f.write(c["sts_text"])
        with open(fp, "w", newline="") as f:
            f.write(c["fasta_text"])
        eng = MerPCR(**c["params"], **engine_kw)
        try:
            ok = eng.load_sts_file(sp)
            assert ok == ex["load_ok"], c["seed"]
            if not ok:
                return
            got = [[r.id, r.direct, r.hash_offset, r.pcr_size, r.offset, r.primer1, r.primer2, r.alias]
                   for r in eng.sts_records]
            assert got == ex["records"], c["seed"]
            assert eng.max_pcr_size == ex["max_pcr_size"]
            try:
                recs = eng.load_fasta_file(fp)
            except IndexError:
                assert ex["error"] == "fasta:IndexError"
                return
            assert ex["error"] is None
            assert [[r.label, len(r.sequence), hashlib.sha256(r.sequence.encode()).hexdigest()[:16]]
                    for r in recs] == ex["fasta"], c["seed"]
            n = eng.search(recs, op)
            with open(op, newline="") as f:
                text = f.read()
            assert (n, text) == (ex["hits"], ex["output"]), (c["seed"], c["params"])
            assert eng.total_hits == n
        finally:
            eng.close()


def oracle_hits(params, sts_text, contigs):
    """Oracle hit list as an array of (contig, pos1, pos2, line_index, strand) rows in output order."""
    o = Oracle(**params)
    assert o.load_sts_text(sts_text)
    recs = o.records()
    line = np.array([r["offset"] for r in recs], dtype=np.int64)
    minus = np.array([r["direct"] == "-" for r in recs], dtype=np.int64)
    rows = []
    for ci, seq in enumerate(contigs):
        h = o.search_hits(bytes(seq) if not isinstance(seq, (bytes, str)) else seq)
        if len(h):
            rows.append(np.stack([np.full(len(h), ci), h[:, 0], h[:, 1], line[h[:, 2]], minus[h[:, 2]]], axis=1))
    return np.concatenate(rows) if rows else np.zeros((0, 5), dtype=np.int64)


def engine_hits(eng, records):
    """Same rows from merpcr_b200 (rec = 2*accepted_line_index + strand -> source line number via sts_records)."""
    h = eng.search_hits(records)
    if h.size == 0:
        return np.zeros((0, 5), dtype=np.int64)
    idx = eng._rec_to_idx[h["rec"]]
    line = np.array([eng.sts_records[i].offset for i in idx.tolist()], dtype=np.int64)
    return np.stack([h["contig"].astype(np.int64), h["pos1"].astype(np.int64), h["pos2"].astype(np.int64), line,
                     (h["rec"] & 1).astype(np.int64)], axis=1)
