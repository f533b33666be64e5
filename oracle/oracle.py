"""ctypes wrapper around oracle/merpcr_oracle.c (test infrastructure, not product code).

`Oracle` mirrors the reference engine's surface (`MerPCR`, /root/reference/src/merpcr/core/engine.py:44)
closely enough that parity tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmerpcr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds). Safe to call repeatedly."""
    src = os.path.join(_HERE, "merpcr_oracle.c")
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, cp, i32, i64, sz = C.c_void_p, C.c_char_p, C.c_int, C.c_longlong, C.c_size_t
        L.orc_new.restype = vp
        L.orc_new.argtypes = [i32, i32, i32, i32, i32, i64, C.c_char_p, sz]
        L.orc_free.argtypes = [vp]
        L.orc_load_sts_text.restype = i32
        L.orc_load_sts_text.argtypes = [vp, cp, sz]
        L.orc_load_sts_file.restype = i32
        L.orc_load_sts_file.argtypes = [vp, cp]
        L.orc_num_records.restype = i32
        L.orc_num_records.argtypes = [vp]
        L.orc_max_pcr_size.restype = i64
        L.orc_max_pcr_size.argtypes = [vp]
        for name in ("orc_rec_id", "orc_rec_alias", "orc_rec_primer1", "orc_rec_primer2"):
            getattr(L, name).restype = cp
            getattr(L, name).argtypes = [vp, i32]
        L.orc_rec_pcr_size.restype = i64
        L.orc_rec_pcr_size.argtypes = [vp, i32]
        L.orc_rec_hash_offset.restype = i32
        L.orc_rec_hash_offset.argtypes = [vp, i32]
        L.orc_rec_hash.restype = C.c_uint32
        L.orc_rec_hash.argtypes = [vp, i32]
        L.orc_rec_line.restype = i32
        L.orc_rec_line.argtypes = [vp, i32]
        L.orc_rec_direct.restype = C.c_char
        L.orc_rec_direct.argtypes = [vp, i32]
        L.orc_rec_lines.argtypes = [vp, vp, vp]
        L.orc_last_error.restype = cp
        L.orc_last_error.argtypes = [vp]
        L.orc_hash_value.restype = i32
        L.orc_hash_value.argtypes = [vp, cp, i32, C.POINTER(C.c_uint32)]
        L.orc_reverse_complement.argtypes = [vp, cp, i32, C.c_char_p]
        L.orc_compare_seqs.restype = i32
        L.orc_compare_seqs.argtypes = [vp, cp, i32, cp, i32, C.c_char]
        L.orc_load_fasta_text.restype = vp
        L.orc_load_fasta_text.argtypes = [cp, sz]
        L.orc_load_fasta_file.restype = vp
        L.orc_load_fasta_file.argtypes = [cp]
        L.orc_fasta_count.restype = i32
        L.orc_fasta_count.argtypes = [vp]
        L.orc_fasta_error.restype = i32
        L.orc_fasta_error.argtypes = [vp]
        for name in ("orc_fasta_label", "orc_fasta_defline"):
            getattr(L, name).restype = cp
            getattr(L, name).argtypes = [vp, i32]
        L.orc_fasta_seq.restype = vp
        L.orc_fasta_seq.argtypes = [vp, i32]
        L.orc_fasta_len.restype = sz
        L.orc_fasta_len.argtypes = [vp, i32]
        L.orc_fasta_free.argtypes = [vp]
        L.orc_search_seq_text.restype = vp
        L.orc_search_seq_text.argtypes = [vp, cp, cp, sz, i32, C.POINTER(i64)]
        L.orc_search_seq_hits.restype = i64
        L.orc_search_seq_hits.argtypes = [vp, cp, sz, i32, C.POINTER(C.POINTER(C.c_int64))]
        L.orc_search_seq_count.restype = i64
        L.orc_search_seq_count.argtypes = [vp, C.c_void_p, sz, i32]
        L.orc_run_files.restype = vp
        L.orc_run_files.argtypes = [vp, cp, cp, i32, C.POINTER(i64), C.POINTER(i32)]
        L.orc_free_text.argtypes = [vp]
        L.orc_now.restype = C.c_double
        _lib = L
    return _lib


class Oracle:
    """CPU restatement of MerPCR (engine.py:44-642). Constructor raises ValueError like engine.py:80-97."""

    def __init__(self, wordsize=11, margin=50, mismatches=0, three_prime_match=1, iupac_mode=0,
                 default_pcr_size=240, threads=1):
        L = lib()
        err = C.create_string_buffer(256)
        self._h = L.orc_new(wordsize, margin, mismatches, three_prime_match, 1 if iupac_mode else 0,
                            default_pcr_size, err, 256)
        if not self._h:
            raise ValueError(err.value.decode())
        self.wordsize, self.margin, self.mismatches = wordsize, margin, mismatches
        self.three_prime_match, self.iupac_mode, self.threads = three_prime_match, iupac_mode, threads

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.orc_free(h)

    # -- STS -------------------------------------------------------------
    def load_sts_text(self, text) -> bool:
        b = text.encode("latin-1") if isinstance(text, str) else text
        return lib().orc_load_sts_text(self._h, b, len(b)) == 1

    def load_sts_file(self, path: str) -> bool:
        r = lib().orc_load_sts_file(self._h, os.fsencode(path))
        if r < 0:
            raise FileNotFoundError(path)
        return r == 1

    @property
    def num_records(self) -> int:
        return lib().orc_num_records(self._h)

    @property
    def max_pcr_size(self) -> int:
        return lib().orc_max_pcr_size(self._h)

    def record(self, i: int) -> dict:
        L = lib()
        return dict(
            id=L.orc_rec_id(self._h, i).decode("latin-1"), alias=L.orc_rec_alias(self._h, i).decode("latin-1"),
            primer1=L.orc_rec_primer1(self._h, i).decode("latin-1"),
            primer2=L.orc_rec_primer2(self._h, i).decode("latin-1"),
            pcr_size=L.orc_rec_pcr_size(self._h, i), hash_offset=L.orc_rec_hash_offset(self._h, i),
            hash=L.orc_rec_hash(self._h, i), offset=L.orc_rec_line(self._h, i),
            direct=L.orc_rec_direct(self._h, i).decode(),
        )

    def records(self) -> List[dict]:
        return [self.record(i) for i in range(self.num_records)]

    def record_lines(self):
        """(source line number, is-minus-strand) of every record as two int64 arrays (insertion order)."""
        import numpy as np
        n = self.num_records
        lines = np.zeros(max(n, 1), dtype=np.int32)
        directs = np.zeros(max(n, 1), dtype=np.uint8)
        lib().orc_rec_lines(self._h, lines.ctypes.data, directs.ctypes.data)
        return lines[:n].astype(np.int64), (directs[:n] == ord("-")).astype(np.int64)

    # -- helpers mirrored from the reference's private API ------------------
    def hash_value(self, primer: str) -> Tuple[int, int]:
        h = C.c_uint32(0)
        b = primer.encode("latin-1")
        off = lib().orc_hash_value(self._h, b, len(b), C.byref(h))
        return off, h.value

    def reverse_complement(self, s: str) -> str:
        b = s.encode("latin-1")
        out = C.create_string_buffer(len(b) + 1)
        lib().orc_reverse_complement(self._h, b, len(b), out)
        return out.value.decode("latin-1")

    def compare_seqs(self, seq1: str, seq2: str, strand: str) -> bool:
        a, b = seq1.encode("latin-1"), seq2.encode("latin-1")
        return bool(lib().orc_compare_seqs(self._h, a, len(a), b, len(b), strand.encode()))

    # -- search ------------------------------------------------------------
    def search_text(self, label: str, sequence, threads: Optional[int] = None) -> Tuple[int, str]:
        """Output text for one record, exactly as engine.py:437-444 prints it."""
        s = sequence.encode("latin-1") if isinstance(sequence, str) else bytes(sequence)
        n = C.c_longlong(0)
        p = lib().orc_search_seq_text(self._h, label.encode("latin-1"), s, len(s),
                                      threads or self.threads, C.byref(n))
        try:
            return n.value, C.string_at(p).decode("latin-1")
        finally:
            lib().orc_free_text(p)

    def search_hits(self, sequence, threads: Optional[int] = None):
        """Sorted 0-based (pos1, pos2, record_index) triples for one record."""
        import numpy as np
        s = sequence.encode("latin-1") if isinstance(sequence, str) else bytes(sequence)
        ptr = C.POINTER(C.c_int64)()
        n = lib().orc_search_seq_hits(self._h, s, len(s), threads or self.threads, C.byref(ptr))
        try:
            arr = np.ctypeslib.as_array(ptr, shape=(max(n, 1) * 3,))[: n * 3].copy().reshape(-1, 3)
        finally:
            lib().orc_free_text(ptr)
        return arr

    def search_hits_array(self, arr, threads: int = 1):
        """`search_hits` over a contiguous uint8 numpy array, without copying it (the C call releases the GIL, so
        several contigs can be searched from Python threads at once; the engine is read-only while searching)."""
        import numpy as np
        assert arr.dtype == np.uint8 and arr.flags.c_contiguous
        ptr = C.POINTER(C.c_int64)()
        n = lib().orc_search_seq_hits(self._h, C.cast(C.c_void_p(arr.ctypes.data), C.c_char_p), int(arr.size), threads,
                                      C.byref(ptr))
        try:
            out = np.ctypeslib.as_array(ptr, shape=(max(n, 1) * 3,))[: n * 3].copy().reshape(-1, 3)
        finally:
            lib().orc_free_text(ptr)
        return out

    def search_count_buffer(self, addr: int, length: int, threads: int) -> int:
        """Count-only search over a raw ASCII buffer (timing legs; nothing formatted)."""
        return lib().orc_search_seq_count(self._h, addr, length, threads)

    def search(self, records, threads: Optional[int] = None) -> Tuple[int, str]:
        """records: iterable of (label, sequence). Returns (total_hits, output_text) like MerPCR.search."""
        total, parts = 0, []
        for label, seq in records:
            n, t = self.search_text(label, seq, threads)
            total += n
            parts.append(t)
        return total, "".join(parts)

    def run_files(self, sts_path: str, fasta_path: str, threads: Optional[int] = None) -> Tuple[int, int, str]:
        """(exit_status, hits, output_text) of `merpcr sts fa` (cli.py:232-255)."""
        n, st = C.c_longlong(0), C.c_int(0)
        p = lib().orc_run_files(self._h, os.fsencode(sts_path), os.fsencode(fasta_path),
                                threads or self.threads, C.byref(n), C.byref(st))
        if not p:
            return st.value, 0, ""
        try:
            return st.value, n.value, C.string_at(p).decode("latin-1")
        finally:
            lib().orc_free_text(p)


def load_fasta_text(text):
    """[(defline, label, sequence)] per io/fasta.py:19-71; raises IndexError for a bare '>' header (models.py:49)."""
    b = text.encode("latin-1") if isinstance(text, str) else text
    L = lib()
    fa = L.orc_load_fasta_text(b, len(b))
    try:
        if L.orc_fasta_error(fa):
            raise IndexError("list index out of range")
        out = []
        for i in range(L.orc_fasta_count(fa)):
            seq = C.string_at(L.orc_fasta_seq(fa, i), L.orc_fasta_len(fa, i)).decode("latin-1")
            out.append((L.orc_fasta_defline(fa, i).decode("latin-1"), L.orc_fasta_label(fa, i).decode("latin-1"), seq))
        return out
    finally:
        L.orc_fasta_free(fa)
