"""Data models of the STS search -- same names, fields and defaults as the reference's
`merpcr/core/models.py:17-69` (STSRecord, FASTARecord, STSHit, ThreadData), so code written against the
reference keeps working.  `FASTARecord` can additionally carry its sequence as a byte array so that a
3 Gbp genome never has to exist as a Python `str`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional

import numpy as np


class SeqType(Enum):
    """Sequence type enumeration (models.py:10-14; unused by the engine, kept for import compatibility)."""

    AMINO_ACID = 1
    NUCLEOTIDE = 2


@dataclass
class STSRecord:
    """One strand of one STS line (models.py:17-29)."""

    id: str
    primer1: str
    primer2: str
    pcr_size: int
    alias: str = ""
    offset: int = 0  # line number in the STS file
    hash_offset: int = 0  # offset of the hash word inside primer1
    direct: str = "+"  # '+' or '-'
    ambig_primer: int = 0


class FASTARecord:
    """One FASTA record (models.py:32-49): `FASTARecord(defline, sequence, label="")`.

    `sequence` may be a `str` (reference behaviour) or a bytes-like / uint8 array (what our loader produces);
    `.sequence` always reads back as `str`, `.sequence_bytes` as a uint8 array, each converted lazily.
    """

    __slots__ = ("defline", "label", "_seq_str", "_seq_bytes", "_seq_dev", "_from_loader")

    def __init__(self, defline: str, sequence, label: str = ""):
        self.defline = defline
        self._from_loader = False
        self._seq_dev = None
        if hasattr(sequence, "data_ptr") and hasattr(sequence, "device"):
            # a torch uint8 tensor (the device-side FASTA ingest leaves the filtered bases in HBM); host views
            # are materialised lazily
            self._seq_str = None
            self._seq_bytes = None
            self._seq_dev = sequence
        elif isinstance(sequence, str):
            self._seq_str: Optional[str] = sequence
            self._seq_bytes: Optional[np.ndarray] = None
        else:
            self._seq_str = None
            self._seq_bytes = np.frombuffer(sequence, dtype=np.uint8) if not isinstance(sequence, np.ndarray) \
                else np.ascontiguousarray(sequence, dtype=np.uint8)
        self.label = label
        if not self.label:  # models.py:40-49
            if ">" in self.defline:
                d = self.defline.strip()[1:]
            else:
                d = self.defline.strip()
            self.label = d.split()[0]  # IndexError for a bare '>' header, like the reference

    @property
    def sequence(self) -> str:
        if self._seq_str is None:
            self._seq_str = self.sequence_bytes.tobytes().decode("latin-1")
        return self._seq_str

    @sequence.setter
    def sequence(self, value) -> None:
        self.__init__(self.defline, value, self.label)

    @property
    def sequence_bytes(self) -> np.ndarray:
        """uint8 view of the sequence; raises ValueError for non-ASCII text (not encodable on the device)."""
        if self._seq_bytes is None and self._seq_dev is not None:
            self._seq_bytes = self._seq_dev.cpu().numpy()
        if self._seq_bytes is None:
            try:
                self._seq_bytes = np.frombuffer(self._seq_str.encode("ascii"), dtype=np.uint8)
            except UnicodeEncodeError as e:
                raise ValueError(f"sequence '{self.label}' contains non-ASCII characters") from e
        return self._seq_bytes

    @property
    def sequence_device(self):
        """The torch uint8 tensor holding the sequence in device memory, or None."""
        return self._seq_dev

    def __len__(self) -> int:
        if self._seq_str is not None:
            return len(self._seq_str)
        if self._seq_bytes is not None:
            return int(self._seq_bytes.size)
        return int(self._seq_dev.numel())

    def __eq__(self, other) -> bool:
        if not isinstance(other, FASTARecord):
            return NotImplemented
        return (self.defline, self.sequence, self.label) == (other.defline, other.sequence, other.label)

    def __repr__(self) -> str:
        s = self.sequence
        shown = s if len(s) <= 60 else s[:57] + "..."
        return f"FASTARecord(defline={self.defline!r}, sequence={shown!r}, label={self.label!r})"


@dataclass
class STSHit:
    """One hit (models.py:52-58); positions 0-based inclusive."""

    pos1: int
    pos2: int
    sts: STSRecord


@dataclass
class ThreadData:
    """Kept for API compatibility (models.py:61-69); the device path does not chunk by thread."""

    thread_id: int
    sequence: str
    offset: int
    length: int
    hits: List[STSHit] = field(default_factory=list)
