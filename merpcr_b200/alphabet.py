"""Byte -> code look-up tables shared by the host and the device kernels.

Reference semantics (paths relative to /root/reference/src/merpcr/):
  * hashing code `scode` A0 C1 G2 T3 U3, everything else ambiguous           core/engine.py:102-109
  * IUPAC letter sets                                                         core/engine.py:138-172
  * non-IUPAC compare = identical letter, IUPAC compare = set intersection,
    letters outside the map compare by identity                               core/engine.py:614-631
  * FASTA files only ever contain `ACGTBDHKMNRSVWXY` (either case)            io/fasta.py:60

The 4-bit IUPAC mask (A1 C2 G4 T8 ... N15) is a bijection on the 15 "real" letters, so one nibble carries
both identity (non-IUPAC equality == nibble equality) and the IUPAC sets (match == AND != 0).  Code 0 is the
single spare value: it stands for the one "exotic" letter (normally X) that matches only itself.
"""
from __future__ import annotations

import numpy as np

IUPAC_MASK = {
    "A": 1, "C": 2, "G": 4, "T": 8, "M": 3, "R": 5, "S": 6, "V": 7, "W": 9, "Y": 10, "H": 11, "K": 12,
    "D": 13, "B": 14, "N": 15,
}
FASTA_KEEP = "ACGTBDHKMNRSVWXY"  # io/fasta.py:60
NON_ASCII = 0x80                 # placeholder byte for non-ASCII primer characters (never matches)

_SCODE = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}


def real_letters(iupac_mode: int) -> str:
    """Letters that have a proper nibble code in this mode (U is T only under IUPAC rules)."""
    return "".join(IUPAC_MASK) + ("U" if iupac_mode else "")


def genome_lut(iupac_mode: int) -> np.ndarray:
    """byte -> nibble | code2 << 4 | clean << 6 for sequence characters (case-insensitive, engine.py:455)."""
    lut = np.zeros(256, dtype=np.uint8)
    for ch, m in IUPAC_MASK.items():
        for c in (ch, ch.lower()):
            lut[ord(c)] = m
    if iupac_mode:
        lut[ord("U")] = lut[ord("u")] = 8
    for ch, code in _SCODE.items():
        for c in (ch, ch.lower()):
            lut[ord(c)] |= (code << 4) | (1 << 6)
    return lut


def genome_exotics(seq_bytes: np.ndarray, iupac_mode: int) -> set:
    """Upper-cased characters of a sequence that have no nibble code of their own."""
    present = np.flatnonzero(np.bincount(seq_bytes, minlength=256))
    real = set(real_letters(iupac_mode))
    return {chr(b).upper() for b in present if chr(b).upper() not in real}


def primer_lut(iupac_mode: int, zero_char: str = "X") -> np.ndarray:
    """byte -> nibble | never_match << 4 | zero_code_char << 5 for (upper-cased) primer characters.

    `zero_char` is the one exotic letter the sequences may contain (code 0 on the genome side); every other
    exotic primer letter can match nothing and is flagged never_match.
    """
    lut = np.full(256, 1 << 4, dtype=np.uint8)
    for ch, m in IUPAC_MASK.items():
        lut[ord(ch)] = m
    if iupac_mode:
        lut[ord("U")] = 8
    if zero_char:
        lut[ord(zero_char)] = 1 << 5
    return lut


def primer_exotics(primers_upper, iupac_mode: int) -> set:
    real = set(real_letters(iupac_mode))
    out = set()
    for p in primers_upper:
        out.update(set(p) - real)
    return out
