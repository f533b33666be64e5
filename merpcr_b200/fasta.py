"""FASTA ingest, host side: file bytes -> per-record filtered sequence bytes.

Same observable behaviour as the reference's `FASTALoader.load_file` (io/fasta.py:19-71) and
`FASTARecord.__post_init__` (core/models.py:40-49): text-mode universal newlines, lines are `strip()`ped
before the `>` test, blank lines skipped, data before the first header dropped, every sequence character
outside `ACGTBDHKMNRSVWXY` (either case) silently removed (so `U` vanishes), case preserved,
empty file -> `[]`, missing file -> the `os.path.getsize` exception propagates.

The sequence is kept as a uint8 array (1 byte/base); packing into the 2-bit / 4-bit device planes happens
on the GPU (`mpcr_pack_sequence`).
"""
from __future__ import annotations

import locale
import logging
import os
import time
from typing import List

import numpy as np

from .alphabet import FASTA_KEEP
from .models import FASTARecord

logger = logging.getLogger("merpcr.io.fasta")  # same logger name as the reference module

_WS = b" \t\n\r\x0b\x0c\x1c\x1d\x1e\x1f"      # ASCII characters removed by str.strip()
_KEEP = np.zeros(256, dtype=bool)
for _c in FASTA_KEEP + FASTA_KEEP.lower():
    _KEEP[ord(_c)] = True
_IS_WS = np.zeros(256, dtype=bool)
for _b in _WS:
    _IS_WS[_b] = True


def _filter_bytes(seg: np.ndarray) -> np.ndarray:
    """io/fasta.py:60 on a byte segment (line terminators are not in the keep set, so lines need no splitting)."""
    if seg.size == 0:
        return np.zeros(0, dtype=np.uint8)
    out = []
    step = 1 << 26
    for s in range(0, seg.size, step):
        chunk = seg[s: s + step]
        out.append(chunk[_KEEP[chunk]])
    return out[0] if len(out) == 1 else np.concatenate(out)


def _parse_ascii(a: np.ndarray) -> List[FASTARecord]:
    gt = np.flatnonzero(a == ord(">"))
    headers = []  # (line_start_of_'>', end_of_line)
    if gt.size:
        term = np.flatnonzero((a == 10) | (a == 13))
        for i in gt.tolist():
            k = int(np.searchsorted(term, i))              # terminators before i
            line_start = int(term[k - 1]) + 1 if k > 0 else 0
            if headers and line_start < headers[-1][1]:
                continue                                    # a '>' inside an already recognised header line
            if i > line_start and not _IS_WS[a[line_start:i]].all():
                continue                                    # '>' in the middle of a sequence line: filtered out later
            line_end = int(term[k]) if k < term.size else int(a.size)
            headers.append((i, line_end))
    records = []
    for h, (i, line_end) in enumerate(headers):
        defline = a[i:line_end].tobytes().decode("ascii").strip()
        nxt = headers[h + 1][0] if h + 1 < len(headers) else int(a.size)
        seq = _filter_bytes(a[line_end:nxt])
        rec = FASTARecord(defline=defline, sequence=seq)
        rec._from_loader = True
        records.append(rec)
    return records


def _parse_text(text: str) -> List[FASTARecord]:
    """Slow exact path for non-ASCII files (same rules, on decoded text)."""
    records, cur, parts = [], None, []
    keep = set(FASTA_KEEP)
    for raw in text.replace("\r\n", "\n").replace("\r", "\n").split("\n"):
        line = raw.strip()
        if not line:
            continue
        if line.startswith(">"):
            if cur is not None:
                records.append(FASTARecord(defline=cur, sequence="".join(parts)))
            cur, parts = line, []
        else:
            parts.append("".join(c for c in line if c.upper() in keep))
    if cur is not None:
        records.append(FASTARecord(defline=cur, sequence="".join(parts)))
    for r in records:
        r._from_loader = True
    return records


DEVICE_INGEST_MIN_BYTES = 1 << 20   # below this the host parser is as fast as the round trip
RANK_MARGIN_LEFT = 1 << 16          # rank-local ingest: file bytes read in front of / behind a rank's own byte range
RANK_MARGIN_RIGHT = 1 << 20


def _file_to_device(filename: str, lo: int, hi: int, engine):
    """File bytes [lo, hi) -> a device tensor, in 64 MiB pieces: every piece is read by several pread streams at once
    (mpcr_file_read; one thread copies out of the page cache at a few GB/s) into one of three pinned staging buffers and
    goes up on a copy stream while the next piece is being read, so the wall clock is max(read, PCIe), not their sum.
    Returns (tensor, bytes actually read)."""
    import torch
    lib, dev = engine._be.lib, engine._tdev
    gpu = dev.type == "cuda"
    size = hi - lo
    chunk = min(1 << 26, max(1 << 16, size))
    text = torch.empty(max(size, 1), dtype=torch.uint8, device=dev)
    stages = getattr(engine, "_file_stage", None)
    if stages is None or stages[0].numel() < chunk:
        stages = [torch.empty(chunk, dtype=torch.uint8, pin_memory=gpu) for _ in range(3)]
        engine._file_stage = stages
    threads = getattr(engine, "_pack_threads", 0)
    path_b = os.fsencode(filename)
    if gpu:
        compute = torch.cuda.current_stream(dev)
        if engine._copy_stream is None:
            engine._copy_stream = torch.cuda.Stream(dev)
        copy = engine._copy_stream
        copy.wait_stream(compute)
        free = [torch.cuda.Event() for _ in stages]
    got = 0
    for k, off in enumerate(range(0, size, chunk)):
        n = min(chunk, size - off)
        slot = k % len(stages)
        if gpu and k >= len(stages):
            free[slot].synchronize()
        r = int(lib.mpcr_file_read(path_b, lo + off, n, stages[slot].data_ptr(), threads))
        if r < 0:
            raise OSError(-r, os.strerror(-r), filename)
        if gpu:
            with torch.cuda.stream(copy):
                text[off: off + r].copy_(stages[slot][:r], non_blocking=True)
            free[slot].record(copy)
        else:
            text[off: off + r].copy_(stages[slot][:r])
        got += r
        if r < n:
            break
    if gpu:
        compute.wait_stream(copy)
    return text, got


def _index_text(text, size: int, engine, mode: int = 0):
    """mpcr_fasta_index_ex on a text tensor.  Returns (records array, workspace tensor, flags)."""
    import ctypes as C

    import torch

    from . import _capi
    lib, be, ctx, dev = engine._be.lib, engine._be, engine._ctx, engine._tdev
    stream = engine._stream()
    cap = 1 << 16
    while True:
        ws = torch.empty(int(lib.mpcr_fasta_workspace_bytes(size, cap)), dtype=torch.uint8, device=dev)
        recs = np.zeros(cap, dtype=_capi.FASTA_RECORD_DTYPE)
        n_rec, flags = C.c_uint32(0), C.c_uint32(0)
        rc = lib.mpcr_fasta_index_ex(ctx, text.data_ptr(), size, mode, recs.ctypes.data, cap, C.byref(n_rec), C.byref(flags),
                                     ws.data_ptr(), ws.numel(), stream)
        if rc == _capi.MPCR_EOVERFLOW:
            # header lines are not blanked yet at this point, so the call can simply be repeated with room for all
            # (a slice's first line is blanked by then, which repeating does not change)
            cap = int(n_rec.value) + 16
            continue
        be.check(rc)
        break
    return recs[: int(n_rec.value)], ws, int(flags.value)


def _deflines(filename: str, spans):
    """Header lines [begin, end) of the file as stripped text: a few are read straight from the file, very many in one go."""
    raw = np.fromfile(filename, dtype=np.uint8) if len(spans) > 4096 else None
    out = []
    with open(filename, "rb", buffering=0) as f:
        for hb, he in spans:
            line = raw[hb:he].tobytes() if raw is not None else os.pread(f.fileno(), he - hb, hb)
            out.append(line.decode("ascii").strip())
    return out


def _device_ingest(filename: str, size: int, engine) -> List[FASTARecord] | None:
    """FASTA text ingest on the GPU (mpcr_fasta_index / mpcr_fasta_compact): file bytes -> pinned host buffer -> HBM ->
    header table + filtered bases, which stay in device memory.  Returns None for a non-ASCII file (the host parser
    applies the locale rules)."""
    import torch
    lib, be, ctx, dev = engine._be.lib, engine._be, engine._ctx, engine._tdev
    text, size = _file_to_device(filename, 0, size, engine)
    recs, ws, flags = _index_text(text, size, engine)
    if flags & 1:
        return None
    n = len(recs)
    total = int(recs["seq_offset"][-1] + recs["seq_length"][-1]) if n else 0
    seq = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    be.check(lib.mpcr_fasta_compact(ctx, text.data_ptr(), size, ws.data_ptr(), seq.data_ptr(), engine._stream()))
    engine._sync()
    out = []
    deflines = _deflines(filename, [(int(r["header_begin"]), int(r["header_end"])) for r in recs])
    for r, defline in zip(recs, deflines):
        a, b = int(r["seq_offset"]), int(r["seq_offset"] + r["seq_length"])
        rec = FASTARecord(defline=defline, sequence=seq[a:b])
        rec._from_loader = True
        out.append(rec)
    return out


class ShardedFASTARecord(FASTARecord):
    """A record of a rank-local ingest: every rank knows the defline and the length of every record, but holds only
    the bases of its own byte range of the file (plus halos): `piece` = (device tensor, contig-local offset of its first
    base) or None."""

    __slots__ = ("_length", "piece")

    def __init__(self, defline: str, length: int, piece=None):
        FASTARecord.__init__(self, defline, np.zeros(0, dtype=np.uint8))
        self._length = int(length)
        self.piece = piece
        self._from_loader = True

    def __len__(self) -> int:
        return self._length

    @property
    def sequence(self) -> str:
        raise RuntimeError(f"sequence '{self.label}' is distributed over the ranks of a multi-GPU run "
                           "(rank-local FASTA ingest); only its length and this rank's piece are held here")

    @property
    def sequence_bytes(self):
        raise RuntimeError(f"sequence '{self.label}' is distributed over the ranks of a multi-GPU run")


class ShardedRecords(list):
    """What a rank-local ingest returns: all records (ShardedFASTARecord) + the range of the padded genome coordinate
    this rank owns (engine.search scans exactly that range)."""

    owned_range = None


def _ranked_ingest(filename: str, size: int, engine):
    """Rank-local FASTA ingest for one rank of a multi-GPU run: the rank reads only its own byte range of the file (plus
    margins that hold the halos), indexes and compacts that slice on its GPU, and one exchange of a few numbers per
    record boundary (all_gather of python objects; nothing on the scan path) gives every rank the complete record
    table -- deflines, lengths -- and tells it which stretch of which record its bytes are.  A rank then owns the scan
    positions from its first base up to the next rank's first base.  Returns None when the file cannot be split this way
    (non-ASCII text, lines longer than the margins, halos larger than the margins): every rank then ingests the whole file."""
    import torch
    import torch.distributed as dist
    lib, be, ctx, dev = engine._be.lib, engine._be, engine._ctx, engine._tdev
    rank, world = engine.shard
    left = int(os.environ.get("MPCR_RANK_MARGIN_LEFT", RANK_MARGIN_LEFT))
    right = int(os.environ.get("MPCR_RANK_MARGIN_RIGHT", RANK_MARGIN_RIGHT))
    bounds = [size * r // world for r in range(world + 1)]
    lo, hi = max(0, bounds[rank] - left), min(size, bounds[rank + 1] + right)
    text, got = _file_to_device(filename, lo, hi, engine)
    ok = got == hi - lo
    recs = ws = None
    if ok:
        recs, ws, flags = _index_text(text, got, engine, mode=3 if lo > 0 else 0)
        ok = flags == 0
    c0, c1 = bounds[rank] - lo, bounds[rank + 1] - lo           # my own byte range in slice coordinates
    segs, k0 = [], 0
    if ok:
        pos = np.array([c0, c1], dtype=np.uint64)
        kept = np.zeros(2, dtype=np.uint64)
        be.check(lib.mpcr_fasta_offsets_at(ctx, text.data_ptr(), got, ws.data_ptr(), pos.ctypes.data, 2, kept.ctypes.data,
                                           engine._stream()))
        k0, k1 = int(kept[0]), int(kept[1])
        total = int(recs["seq_offset"][-1] + recs["seq_length"][-1]) if len(recs) else 0
        # the margins must hold the halos (unless the slice reaches the end of the file on that side)
        need_l = int(lib.mpcr_halo_left(ctx)) + 256
        need_r = int(lib.mpcr_halo_right(ctx)) + 4096
        if (lo > 0 and k0 < need_l) or (hi < size and total - k1 < need_r):
            ok = False
        # my segments: for every record with bases or its header in my own range, (header span or None, bases there)
        for r in recs:
            hb = int(r["header_begin"])
            pseudo = int(r["header_end"]) == hb
            a, b = int(r["seq_offset"]), int(r["seq_offset"] + r["seq_length"])
            mine = max(0, min(b, k1) - max(a, k0))
            header_here = (not pseudo) and c0 <= hb < c1
            if header_here or mine > 0:
                segs.append(((hb + lo, int(r["header_end"]) + lo) if header_here else None, mine))
    everyone = [None] * world
    dist.all_gather_object(everyone, (ok, segs))
    if not all(o for o, _ in everyone):
        return None
    # the global record table; where every rank's first base lies
    spans, lengths, first = [], [], []          # first[r] = (record index, offset inside it) of rank r's first base
    for _, rsegs in everyone:
        first.append(None)
        for span, n in rsegs:
            if span is not None:
                spans.append(span)
                lengths.append(0)
            if not lengths:
                continue                        # sequence in front of the first header of the file: dropped (io/fasta.py:51)
            if first[-1] is None and n > 0:
                first[-1] = (len(lengths) - 1, lengths[-1])
            lengths[-1] += n
    deflines = _deflines(filename, spans)
    # which records do my slice's records correspond to?  consecutive records of the file, anchored at my first base
    seq = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    be.check(lib.mpcr_fasta_compact(ctx, text.data_ptr(), got, ws.data_ptr(), seq.data_ptr(), engine._stream()))
    engine._sync()
    pieces = {}
    anchor = None                               # (slice record index, global record index, offset of the slice data in it)
    if first[rank] is not None:
        g_idx, g_off = first[rank]
        for j, r in enumerate(recs):
            a, b = int(r["seq_offset"]), int(r["seq_offset"] + r["seq_length"])
            if a <= k0 < b or (a == k0 and b > a) or (j + 1 == len(recs) and k0 >= a):
                # my first own base lies in slice record j, (k0 - a) bases into its slice data
                anchor = (j, g_idx, g_off - (k0 - a))
                break
    if anchor is not None:
        j0, g0, off0 = anchor
        for j, r in enumerate(recs):
            g = g0 + (j - j0)
            if g < 0 or g >= len(lengths):
                continue
            a, b = int(r["seq_offset"]), int(r["seq_offset"] + r["seq_length"])
            start = off0 if j == j0 else 0       # other slice records begin with their header: offset 0
            if j < j0:                           # a record that began in my left margin ends where the next begins
                start = 0
            if b > a:
                pieces[g] = (seq[a:b], start)
    out = ShardedRecords()
    for g, (defline, n) in enumerate(zip(deflines, lengths)):
        out.append(ShardedFASTARecord(defline, n, pieces.get(g)))
    # ownership: from my first base (rounded down to the plane granularity) to the next rank's first base
    layout = engine.make_layout(lengths)
    starts = []
    for r in range(world):
        if first[r] is None:
            starts.append(None)
        else:
            gi, go = first[r]
            starts.append((int(layout["contigs"][gi]["gstart"]) + go) // 128 * 128)
    total_g = max(int(layout["total"]), 128)
    nxt = total_g
    ranges = [None] * world
    for r in range(world - 1, -1, -1):
        b = starts[r] if starts[r] is not None else nxt
        ranges[r] = (b, nxt)
        nxt = b
    first_real = next((r for r in range(world) if starts[r] is not None), None)
    if first_real is not None:                   # the first rank that holds bases also owns everything in front of them
        ranges[first_real] = (0, ranges[first_real][1])
        for r in range(first_real):
            ranges[r] = (0, 0)
    out.owned_range = ranges[rank]
    return out


class FASTALoader:
    """Class for loading FASTA files (mirror of io/fasta.py:15)."""

    @staticmethod
    def load_file(filename: str, engine=None) -> List[FASTARecord]:
        """`engine` (a MerPCR bound to a CUDA device) enables the device-side ingest for files >= 1 MiB; the
        records then keep their sequences in HBM (`FASTARecord.sequence` still reads back as `str`)."""
        start = time.time()
        size = os.path.getsize(filename)
        if size == 0:
            logger.error(f"FASTA file '{filename}' is empty")
            return []
        logger.info(f"Reading FASTA file: {filename}")
        records = None
        min_bytes = int(os.environ.get("MPCR_DEVICE_INGEST_MIN_BYTES", DEVICE_INGEST_MIN_BYTES))
        if engine is not None and getattr(engine, "_ctx", None) and size >= min_bytes:
            from . import multi
            if multi.active_world(engine.shard) and os.environ.get("MPCR_RANK_LOCAL_INGEST", "1") not in ("0", ""):
                records = _ranked_ingest(filename, size, engine)       # None: not splittable, everybody reads it all
            if records is None:
                records = _device_ingest(filename, size, engine)
        if records is None:
            a = np.fromfile(filename, dtype=np.uint8)
            if a.size and int(a.max()) < 128:
                records = _parse_ascii(a)
            else:
                records = _parse_text(a.tobytes().decode(locale.getpreferredencoding(False)))
        logger.info(f"Loaded {len(records)} sequences in {time.time() - start:.2f} seconds")
        return records
