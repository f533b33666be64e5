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


class FASTALoader:
    """Class for loading FASTA files (mirror of io/fasta.py:15)."""

    @staticmethod
    def load_file(filename: str) -> List[FASTARecord]:
        start = time.time()
        if os.path.getsize(filename) == 0:
            logger.error(f"FASTA file '{filename}' is empty")
            return []
        logger.info(f"Reading FASTA file: {filename}")
        a = np.fromfile(filename, dtype=np.uint8)
        if a.size and int(a.max()) < 128:
            records = _parse_ascii(a)
        else:
            records = _parse_text(a.tobytes().decode(locale.getpreferredencoding(False)))
        logger.info(f"Loaded {len(records)} sequences in {time.time() - start:.2f} seconds")
        return records
