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


def _device_ingest(filename: str, size: int, engine) -> List[FASTARecord] | None:
    """FASTA text ingest on the GPU (mpcr_fasta_index / mpcr_fasta_compact): file bytes -> pinned host buffer -> HBM ->
    header table + filtered bases, which stay in device memory.  Returns None for a non-ASCII file (the host parser
    applies the locale rules)."""
    import ctypes as C

    import torch

    from . import _capi
    lib, be, ctx, dev = engine._be.lib, engine._be, engine._ctx, engine._tdev
    gpu = dev.type == "cuda"
    # file -> HBM in 64 MiB pieces: every piece is read by several pread streams at once (mpcr_file_read; one thread
    # copies out of the page cache at a few GB/s) into one of three pinned staging buffers and goes up on a copy stream
    # while the next piece is being read, so the wall clock is max(read, PCIe), not their sum
    chunk = min(1 << 26, max(1 << 16, size))
    text = torch.empty(size, dtype=torch.uint8, device=dev)
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
        r = int(lib.mpcr_file_read(path_b, off, n, stages[slot].data_ptr(), threads))
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
    size = got
    stream = engine._stream()
    cap = 1 << 16
    while True:
        ws = torch.empty(int(lib.mpcr_fasta_workspace_bytes(size, cap)), dtype=torch.uint8, device=dev)
        recs = np.zeros(cap, dtype=_capi.FASTA_RECORD_DTYPE)
        n_rec, flags = C.c_uint32(0), C.c_uint32(0)
        rc = lib.mpcr_fasta_index(ctx, text.data_ptr(), size, recs.ctypes.data, cap, C.byref(n_rec), C.byref(flags),
                                  ws.data_ptr(), ws.numel(), stream)
        if rc == _capi.MPCR_EOVERFLOW:
            # header lines are not blanked yet at this point, so the call can simply be repeated with room for all
            cap = int(n_rec.value) + 16
            continue
        be.check(rc)
        break
    if flags.value & 1:
        return None
    n = int(n_rec.value)
    recs = recs[:n]
    total = int(recs["seq_offset"][-1] + recs["seq_length"][-1]) if n else 0
    seq = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    be.check(lib.mpcr_fasta_compact(ctx, text.data_ptr(), size, ws.data_ptr(), seq.data_ptr(), stream))
    engine._sync()
    # deflines: a few header lines are read straight from the file; a file with very many records is read once more
    raw = np.fromfile(filename, dtype=np.uint8) if n > 4096 else None
    out = []
    with open(filename, "rb", buffering=0) as f:
        for r in recs:
            hb, he = int(r["header_begin"]), int(r["header_end"])
            line = raw[hb:he].tobytes() if raw is not None else os.pread(f.fileno(), he - hb, hb)
            defline = line.decode("ascii").strip()
            a, b = int(r["seq_offset"]), int(r["seq_offset"] + r["seq_length"])
            rec = FASTARecord(defline=defline, sequence=seq[a:b])
            rec._from_loader = True
            out.append(rec)
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
            records = _device_ingest(filename, size, engine)
        if records is None:
            a = np.fromfile(filename, dtype=np.uint8)
            if a.size and int(a.max()) < 128:
                records = _parse_ascii(a)
            else:
                records = _parse_text(a.tobytes().decode(locale.getpreferredencoding(False)))
        logger.info(f"Loaded {len(records)} sequences in {time.time() - start:.2f} seconds")
        return records
