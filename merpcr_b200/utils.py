"""Module-level helpers with the names of the reference's `merpcr/core/utils.py:43-113`
(`reverse_complement`, `hash_value`, `init_iupac_tables`) for code that imported them directly."""
from __future__ import annotations

from typing import Dict, Tuple

from .engine import _COMPL, _IUPAC

AMBIG = 100
_scode = [AMBIG] * 256
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    _scode[ord(_ch)] = _scode[ord(_ch.lower())] = _v
_compl: Dict[str, str] = {chr(k): chr(v) for k, v in _COMPL.items()}


def reverse_complement(sequence: str) -> str:
    """utils.py:43-45."""
    return sequence[::-1].translate(_COMPL)


def hash_value(primer: str, wordsize: int) -> Tuple[int, int]:
    """utils.py:48-82: (offset, hash) of the first W-mer free of ambiguity codes, (-1, 0) if none."""
    primer = primer.upper()
    if len(primer) < wordsize:
        return -1, 0
    run, h, mask = 0, 0, (1 << (2 * wordsize)) - 1
    for i, ch in enumerate(primer):
        c = _scode[ord(ch)] if ord(ch) < 256 else AMBIG
        if c == AMBIG:
            run = 0
            continue
        h = ((h << 2) | c) & mask
        run += 1
        if run >= wordsize:
            return i - wordsize + 1, h
    return -1, 0


def init_iupac_tables():
    """utils.py:85-113: (iupac_mapping, iupac_mismatch) -- letter -> set string, incl. lower-case keys."""
    mapping = dict(_IUPAC)
    mapping.update({k.lower(): v for k, v in _IUPAC.items()})
    return mapping, {}
