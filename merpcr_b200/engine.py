"""merpcr_b200 engine: the reference's `MerPCR` class re-hosted on hand-written sm_100a kernels.

Public surface = the reference's (`/root/reference/src/merpcr/core/engine.py:44-451`):

    MerPCR(wordsize=11, margin=50, mismatches=0, three_prime_match=1, iupac_mode=0,
           default_pcr_size=240, threads=1, max_sts_line_length=1022)
    .load_sts_file(path) -> bool        .load_fasta_file(path) -> List[FASTARecord]
    .search(records, output_file=None) -> int
    attributes: sts_records, sts_table, max_pcr_size, total_hits (+ the constructor arguments)
    helpers kept for the reference's tests: _hash_value, _reverse_complement, _compare_seqs

What runs where: STS text rules (engine.py:216-251) and output formatting (:437-444) stay in Python; hashing,
reverse complements, table build, sequence packing, scan, verification, hit compaction and ordering run on
the GPU through the C ABI in include/merpcr_b200.h.  PyTorch only owns device buffers and streams.
There is no CPU fallback; without the CUDA library or a device every entry point raises.
"""
from __future__ import annotations

import logging
import os
import sys
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _capi, multi
from .alphabet import genome_exotics, genome_lut, primer_exotics, primer_lut
from .fasta import FASTALoader
from .models import FASTARecord, STSHit, STSRecord, ThreadData  # noqa: F401  (re-exported like the reference)

# Constants (engine.py:17-39)
AMBIG = 100
MIN_FILESIZE_FOR_THREADING = 100000
DEFAULT_MARGIN = 50
DEFAULT_WORDSIZE = 11
DEFAULT_MISMATCHES = 0
DEFAULT_THREE_PRIME_MATCH = 1
DEFAULT_IUPAC_MODE = 0
DEFAULT_THREADS = 1
DEFAULT_PCR_SIZE = 240
MIN_WORDSIZE, MAX_WORDSIZE = 3, 16
MIN_MISMATCHES, MAX_MISMATCHES = 0, 10
MIN_MARGIN, MAX_MARGIN = 0, 10000
MIN_THREE_PRIME_MATCH = 0
MIN_PCR_SIZE, MAX_PCR_SIZE = 1, 10000

EXTENDED_WORDSIZE = 16        # key width of the extended tables of exact, candidate-heavy searches
STREAM_SCAN_BASES = 1 << 26    # upload_and_scan: scan a finished contig (group) once this many bases are packed
EXT_LINES_PER_TABLE = 125_000  # STS lines per extended table: 2.5*10^5 keys is what the scanner's filter holds at ~5 % f.p.
BLOCK_MIN_KEY = 12            # block tables (searches with mismatches): used when seed + block make at least this many letters
SAMPLE_WORDSIZE = 16          # window width of position-sampled tables (exact, candidate-heavy searches)
SAMPLE_STRIDE = 3             # ... and the stride: primers of >= hash_offset + 18 plain letters can be sampled
PCR_SIZE_CLAMP = 0x7FFFFFFF   # any expected size >= a contig length behaves identically (engine.py:531-533)
PLANE_SLACK_BASES = 1024      # read-ahead of the last strip / last primer window

logger = logging.getLogger("merpcr.core.engine")  # same logger name as the reference module


class _Complement(dict):
    def __missing__(self, key):  # engine.py:359 : unknown -> 'N'
        return ord("N")


_COMPL = _Complement()
for _a, _b in ("AT", "CG", "GC", "TA", "UA", "BV", "DH", "HD", "KM", "MK", "NN", "RY", "SS", "VB", "WW", "XX", "YR"):
    _COMPL[ord(_a)] = ord(_b)
    _COMPL[ord(_a.lower())] = ord(_b.lower())

_IUPAC = {  # engine.py:138-172
    "A": "A", "C": "C", "G": "G", "T": "TU", "U": "TU", "R": "AGR", "Y": "CTUY", "M": "ACM", "K": "GTUK",
    "S": "CGS", "W": "ATUW", "B": "CGTUYKSB", "D": "AGTURKWD", "H": "ACTUYMWH", "V": "ACGRMSV",
    "N": "ACGTURYMKSWBDHVN",
}


def _primer_bytes(p: str) -> bytes:
    if p.isascii():
        return p.encode("ascii")
    return bytes(ord(ch) if ord(ch) < 128 else 0x80 for ch in p)


class _STSLines:
    """The accepted STS lines of one file (after the text rules of engine.py:216-247), in file order.

    Two producers: the native parser (`mpcr_sts_parse`: field byte ranges into the file text, nothing decoded until
    somebody asks) and the Python line loop kept for non-ASCII files."""

    def __init__(self, n, sizes, line_nos, *, raw=None, lines=None, ids=None, aliases=None, p1s=None, p2s=None):
        self.n = n
        self.sizes = sizes          # adjusted expected sizes (Python ints or an int64 array)
        self.line_nos = line_nos
        self.raw, self.lines = raw, lines
        self._ids, self._aliases, self._p1s, self._p2s = ids, aliases, p1s, p2s

    def _field(self, off_key, len_key, upper=False):
        raw, L = self.raw, self.lines
        out = []
        for a, k in zip(L[off_key].tolist(), L[len_key].tolist()):
            t = raw[a: a + k].tobytes().decode("ascii")
            out.append(t.upper() if upper else t)
        return out

    @property
    def ids(self):
        if self._ids is None:
            self._ids = self._field("id_off", "id_len")
        return self._ids

    @property
    def aliases(self):
        if self._aliases is None:
            self._aliases = self._field("alias_off", "alias_len")
        return self._aliases

    @property
    def p1s(self):
        if self._p1s is None:
            self._p1s = self._field("p1_off", "p1_len", upper=True)
        return self._p1s

    @property
    def p2s(self):
        if self._p2s is None:
            self._p2s = self._field("p2_off", "p2_len", upper=True)
        return self._p2s


class _Piece:
    """A stretch of a contig that is available on this rank (rank-local FASTA ingest): `data` holds the bases from
    contig-local offset `start` on."""

    __slots__ = ("data", "start")

    def __init__(self, data, start: int):
        self.data, self.start = data, int(start)


class _Shard:
    """Device-resident packed genome of one shard (planes + the layout they were built for)."""

    __slots__ = ("device", "origin", "bases", "alloc", "plane2", "plane4", "valid", "begin", "end", "hits", "count",
                 "staging", "stage2", "contig_sig", "last_n")


class MerPCR:
    """Main merPCR class that handles all the e-PCR functionality (GPU-backed)."""

    def __init__(
        self,
        wordsize: int = DEFAULT_WORDSIZE,
        margin: int = DEFAULT_MARGIN,
        mismatches: int = DEFAULT_MISMATCHES,
        three_prime_match: int = DEFAULT_THREE_PRIME_MATCH,
        iupac_mode: int = DEFAULT_IUPAC_MODE,
        default_pcr_size: int = DEFAULT_PCR_SIZE,
        threads: int = DEFAULT_THREADS,
        max_sts_line_length: int = 1022,
        *,
        device: Optional[int] = None,
        shard: Optional[Sequence[int]] = None,
        true_strands: bool = False,
    ):
        """Reference parameters (engine.py:47-57) plus two device knobs:

        device : CUDA device index (default: $LOCAL_RANK or 0).
        shard  : (rank, world) -- scan only this rank's bp-balanced share of the genome (multi-GPU runs use
                 one process per GPU; hits are merged by the caller, no collective on the scan path).
        `threads` (-T) is accepted and stored for compatibility; the GPU path does not use host threads.
        true_strands : NOT reference behaviour (off by default).  "+" records look for primer1 ... revcomp(primer2),
                 a biologically normal forward amplicon as NCBI me-PCR reports it, instead of the reference's
                 primer1 ... primer2 (engine.py:267).
        """
        self.wordsize = wordsize
        self.margin = margin
        self.mismatches = mismatches
        self.three_prime_match = three_prime_match
        self.iupac_mode = iupac_mode
        self.default_pcr_size = default_pcr_size
        self.threads = threads
        self.max_sts_line_length = max_sts_line_length

        self._sts_records: Optional[List[STSRecord]] = []
        self._hash_offsets = np.zeros(0, dtype=np.int32)
        self._sts_table: Optional[Dict[int, List[STSRecord]]] = {}
        self.max_pcr_size = 0
        self.total_hits = 0

        self._validate_parameters()
        if threads > 1:
            # with -T > 1 the reference cuts sequences of >= 100 kbp into overlapping chunks and, because
            # its de-duplication compares an absolute position with the overlap length (engine.py:424-431), prints the
            # hits of every overlap twice.  One GPU pass owns every position exactly once.
            logger.warning("-T/threads=%d is accepted for compatibility only: the GPU path does not chunk, so the "
                           "duplicate overlap hits the reference prints with -T > 1 are not reproduced "
                           "(output equals the reference's -T 1)", threads)

        self._be = _capi.backend()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if self._be.device_kind == "cuda" else 0
        self.device = device
        self.shard = (int(shard[0]), int(shard[1])) if shard else (0, 1)
        self.true_strands = bool(true_strands)
        if self._be.device_kind == "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("merpcr_b200 needs a CUDA device (none visible); there is no CPU fallback")
            self._tdev = torch.device("cuda", device)
        else:
            self._tdev = torch.device("cpu")
        self._ctx = None
        self._pinned_hits = None  # D2H staging of the hit list
        self._copy_stream = None  # upload_and_scan: the H2D copies run beside pack + scan
        self._host_stage = None   # pinned staging buffers of the host-side nibble packer
        self._count_host = None   # pinned landing place of a step's hit count
        self._slots = None        # scan_device_async: two pipeline slots (hit buffer, count, pinned result, event)
        # host-resident sequence goes over PCIe as packed nibbles (0.5 byte/base) unless switched off
        self.host_pack = os.environ.get("MPCR_HOST_PACK", "1") not in ("0", "")
        try:
            cpus = len(os.sched_getaffinity(0))
        except AttributeError:  # pragma: no cover
            cpus = os.cpu_count() or 1
        self._pack_threads = max(1, cpus // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        if self._pack_threads >= 8:     # the packer shares the host's memory system with the DMA engine: 3/4 of the cores
            self._pack_threads = self._pack_threads * 3 // 4    # pack as fast as all of them (scripts/gpu/hybrid_probe.py)
        self.hybrid_wire = os.environ.get("MPCR_HYBRID_WIRE", "1") not in ("0", "")
        # a piece is packed on the host while the copy engine has at least this share of a pack's duration queued
        # (0.7 packed 70 % of a human genome: 42.9 ms; 1.4 - 2.0 pack ~60 %, which is where link and host memory balance: 40 ms)
        self.hybrid_backlog = float(os.environ.get("MPCR_HYBRID_BACKLOG", "1.7"))
        if os.environ.get("MPCR_PACK_THREADS"):
            self._pack_threads = max(1, int(os.environ["MPCR_PACK_THREADS"]))
        self._pack_rate = 50e9    # bases / s the host packer sustains (refined while it runs)
        self._h2d_rate = 50e9     # bytes / s of the host -> device link (PCIe Gen5 x16 moves ~55 GB/s)
        self._ctx_exts = []       # extended tables of exact, candidate-heavy searches (mpcr_ctx_set_seed_extension)
        self._ctx_samp = None     # position-sampled table of the same searches (mpcr_ctx_set_sampling)
        self._create_ctx()
        # parsed STS lines kept so the table can be re-encoded if a sequence brings an unusual alphabet
        self._sts_lines = None
        self._zero_char = "X"
        self._rec_to_idx = np.zeros(0, dtype=np.int64)
        self._hashes = np.zeros(0, dtype=np.uint32)
        self.last_scan_ms = 0.0
        self.last_h2d_bytes = 0
        self.last_timing: Dict[str, float] = {}

    # ------------------------------------------------------------------ plumbing
    def _new_ctx(self):
        import ctypes as C
        p = _capi.Params(self.wordsize, self.margin, self.mismatches, self.three_prime_match,
                         1 if self.iupac_mode else 0)
        h = C.c_void_p()
        self._be.check(self._be.lib.mpcr_ctx_create(self.device, C.byref(p), C.byref(h)))
        if self.true_strands:
            self._be.check(self._be.lib.mpcr_ctx_set_true_strands(h, 1))
        return h

    def _create_ctx(self):
        self._ctx = self._new_ctx()

    def close(self):
        ctxs = [getattr(self, "_ctx", None)] + list(getattr(self, "_ctx_exts", [])) + [getattr(self, "_ctx_samp", None)]
        self._ctx, self._ctx_exts, self._ctx_samp = None, [], None
        for ctx in ctxs:
            if ctx:
                self._be.lib.mpcr_ctx_destroy(ctx)

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _stream(self) -> int:
        if self._tdev.type == "cuda":
            return torch.cuda.current_stream(self._tdev).cuda_stream
        return 0

    def _sync(self):
        if self._tdev.type == "cuda":
            torch.cuda.current_stream(self._tdev).synchronize()

    @property
    def gpu_launches(self) -> int:
        return sum(int(self._be.lib.mpcr_launch_count(c)) for c in self._all_ctxs())

    def _all_ctxs(self):
        return [self._ctx] + list(self._ctx_exts) + ([self._ctx_samp] if self._ctx_samp else [])

    # ------------------------------------------------------------------ engine.py:80-97
    def _validate_parameters(self):
        if not (MIN_WORDSIZE <= self.wordsize <= MAX_WORDSIZE):
            raise ValueError(f"Word size must be between {MIN_WORDSIZE} and {MAX_WORDSIZE}")
        if not (MIN_MISMATCHES <= self.mismatches <= MAX_MISMATCHES):
            raise ValueError(f"Number of mismatches must be between {MIN_MISMATCHES} and {MAX_MISMATCHES}")
        if not (MIN_MARGIN <= self.margin <= MAX_MARGIN):
            raise ValueError(f"Margin must be between {MIN_MARGIN} and {MAX_MARGIN}")
        if self.three_prime_match < MIN_THREE_PRIME_MATCH:
            raise ValueError(f"Three prime match must be at least {MIN_THREE_PRIME_MATCH}")
        if not (MIN_PCR_SIZE <= self.default_pcr_size <= MAX_PCR_SIZE):
            raise ValueError(f"Default PCR size must be between {MIN_PCR_SIZE} and {MAX_PCR_SIZE}")

    # ------------------------------------------------------------------ STS loading (engine.py:193-322)
    def load_sts_file(self, filename: str) -> bool:
        """Load STS records from a tab-delimited file; the word-hash table is built on the device."""
        start_time = time.time()
        file_size = os.path.getsize(filename)
        if file_size == 0:
            logger.error(f"STS file '{filename}' is empty")
            return False
        logger.info(f"Reading STS file: {filename}")

        self.sts_records = []
        self._sts_table = {}
        self.max_pcr_size = 0
        self._sts_lines = None

        raw = np.fromfile(filename, dtype=np.uint8)
        parsed = self._parse_sts_native(raw)
        if parsed is None:                      # non-ASCII text: the Python line loop with the locale's decoding
            parsed = self._parse_sts_python(filename)
        if parsed is False:
            self.max_pcr_size = 0
            return False
        self._sts_lines, bad_primers_short, bad_pcr_size = parsed
        self.max_pcr_size = int(max(self._sts_lines.sizes)) if self._sts_lines.n else 0
        self._zero_char = "X"
        bad_primers_ambig = self._build_table()

        if bad_primers_short > 0:
            logger.warning(f"{bad_primers_short} STSs have primer shorter than word size ({self.wordsize}): "
                           "not included in search")
        if bad_primers_ambig > 0:
            logger.warning(f"{bad_primers_ambig} primers have ambiguities which prevent computation of a hash "
                           "value: not included in search")
        if bad_pcr_size > 0:
            logger.warning(f"{bad_pcr_size} STSs have a primer length sum greater than the pcr size: "
                           "expected pcr size adjusted")
        n_records = int(np.count_nonzero(self._hash_offsets >= 0))
        logger.info(f"Loaded {n_records} STS records in {time.time() - start_time:.2f} seconds")
        return True

    def _parse_sts_native(self, raw: np.ndarray):
        """engine.py:216-251 through mpcr_sts_parse.  Returns (lines, n_short, n_adjusted), False for a malformed file,
        None when the text is not ASCII."""
        import ctypes as C
        lib = self._be.lib
        cap = int(np.count_nonzero((raw == 10) | (raw == 13))) + 1
        lines = np.zeros(cap, dtype=_capi.STS_LINE_DTYPE)
        n, bad, short, flags = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        self._be.check(lib.mpcr_sts_parse(raw.ctypes.data, raw.size, self.wordsize, self.default_pcr_size,
                                          lines.ctypes.data, cap, C.byref(n), C.byref(bad), C.byref(short), C.byref(flags)))
        if flags.value & 1:
            return None
        if bad.value:
            logger.error(f"Bad STS file format at line {bad.value}. Expected at least 4 fields.")
            return False
        lines = lines[: n.value]
        sizes = lines["pcr_size"].astype(np.int64)
        slow = np.flatnonzero(sizes < 0)        # fields that are not a plain "123" / "100-200": Python's own int()
        if slow.size:
            sizes = sizes.astype(object)
            for i in slow.tolist():
                a, k = int(lines["size_off"][i]), int(lines["size_len"][i])
                sizes[i] = self._parse_pcr_size(raw[a: a + k].tobytes().decode("ascii"))
        l12 = lines["p1_len"].astype(np.int64) + lines["p2_len"].astype(np.int64)
        adjust = l12 > sizes                    # engine.py:245-247
        n_adjusted = int(np.count_nonzero(adjust))
        if n_adjusted:
            sizes = np.where(adjust, l12, sizes)
        src = _STSLines(int(n.value), sizes, lines["line_no"], raw=raw, lines=lines)
        return src, int(short.value), n_adjusted

    def _parse_sts_python(self, filename: str):
        """The reference's line loop (engine.py:216-251), for files the native parser hands back."""
        bad_primers_short = bad_pcr_size = 0
        ids, aliases, p1s, p2s, sizes, line_nos = [], [], [], [], [], []
        with open(filename, "r") as file:
            line_no = 0
            for line in file:
                line_no += 1
                line = line.strip()
                if not line or line.startswith("#"):
                    continue
                fields = line.split("\t")
                if len(fields) < 4:
                    logger.error(f"Bad STS file format at line {line_no}. Expected at least 4 fields.")
                    return False
                primer1 = fields[1].upper()
                primer2 = fields[2].upper()
                pcr_size = self._parse_pcr_size(fields[3])
                if len(primer1) < self.wordsize or len(primer2) < self.wordsize:
                    bad_primers_short += 1
                    continue
                if len(primer1) + len(primer2) > pcr_size:
                    bad_pcr_size += 1
                    pcr_size = len(primer1) + len(primer2)
                ids.append(fields[0])
                aliases.append(fields[4] if len(fields) > 4 else "")
                p1s.append(primer1)
                p2s.append(primer2)
                sizes.append(pcr_size)
                line_nos.append(line_no)
        src = _STSLines(len(ids), sizes, line_nos, ids=ids, aliases=aliases, p1s=p1s, p2s=p2s)
        return src, bad_primers_short, bad_pcr_size

    def _build_table(self) -> int:
        """Ship the accepted STS lines to the device, build the table there, read back hash offsets."""
        src = self._sts_lines
        n = src.n
        lib = self._be.lib
        if src.lines is not None:               # native source: upper-cased primer blob straight from the file text
            blob = np.zeros(int(src.lines["p1_len"].sum(dtype=np.int64) + src.lines["p2_len"].sum(dtype=np.int64)) + 16,
                            dtype=np.uint8)
            off = np.zeros(2 * n + 1, dtype=np.uint64)
            self._be.check(lib.mpcr_sts_blob(src.raw.ctypes.data, src.lines.ctypes.data, n, blob.ctypes.data,
                                             off.ctypes.data))
        else:
            lens = np.empty(2 * n + 1, dtype=np.uint64)
            lens[0] = 0
            parts = []
            for i in range(n):
                b1, b2 = _primer_bytes(src.p1s[i]), _primer_bytes(src.p2s[i])
                parts.append(b1)
                parts.append(b2)
                lens[2 * i + 1] = len(b1)
                lens[2 * i + 2] = len(b2)
            off = np.cumsum(lens, dtype=np.uint64)
            blob = np.frombuffer(b"".join(parts) + b"\0" * 16, dtype=np.uint8)
        pcr = np.array([min(int(x), PCR_SIZE_CLAMP) for x in src.sizes], dtype=np.uint32) \
            if isinstance(src.sizes, list) or getattr(src.sizes, "dtype", None) == object else \
            np.minimum(np.asarray(src.sizes, dtype=np.int64), PCR_SIZE_CLAMP).astype(np.uint32)
        plut = primer_lut(self.iupac_mode, self._zero_char)
        # Exact searches whose seed words cover a quarter or more of all 4^W words (small -W, many STS) are keyed on
        # 16-letter words instead: the records whose seed extends to 16 plain letters go to extended tables of at most
        # EXT_LINES_PER_TABLE lines each (what the scanner's shared-memory filter can hold), the rest stay in the
        # ordinary table; every table is scanned over the same planes and the hits are sorted once.
        w_ext = EXTENDED_WORDSIZE
        can_extend = self.mismatches == 0 and not self.iupac_mode and self.wordsize < w_ext
        extend = can_extend and 8 * n >= 4 ** self.wordsize
        env = os.environ.get("MPCR_SEED_EXTENSION")
        if env is not None and can_extend:
            extend = env not in ("0", "")
        # Position sampling on top of that: the records whose first primer offers SAMPLE_STRIDE clean 16-letter windows from
        # its hash offset go to ONE sampled table (the scanner then probes every SAMPLE_STRIDE-th position only, through a
        # filter in global memory); the extended / ordinary tables keep the rest.
        stride = SAMPLE_STRIDE if extend else 0
        env = os.environ.get("MPCR_SAMPLING")
        if env is not None and can_extend:
            stride = int(env) if env not in ("0", "") else 0
        rest_lines = n
        if stride >= 2:
            if self._ctx_samp is None:
                self._ctx_samp = self._new_ctx()
            self._be.check(lib.mpcr_ctx_set_sampling(self._ctx_samp, SAMPLE_WORDSIZE, stride, 1))
            self._be.check(lib.mpcr_table_build(self._ctx_samp, blob.ctypes.data, off.ctypes.data, pcr.ctypes.data, n,
                                                plut.ctypes.data, self._stream()))
            sampled_records = int(lib.mpcr_table_items(self._ctx_samp)) // stride
            rest_lines = max(0, n - sampled_records // 2)
        elif self._ctx_samp is not None:
            lib.mpcr_ctx_destroy(self._ctx_samp)
            self._ctx_samp = None
        role = 2 if stride >= 2 else 0
        parts = max(1, -(-rest_lines // EXT_LINES_PER_TABLE)) if extend else 0
        # Candidate-heavy searches that ALLOW mismatches (-N >= 1, no IUPAC mode): block tables.  At most N of the first
        # primer's letters behind the seed differ, so of N + 1 disjoint blocks right behind the seed one is identical:
        # every record whose W + (N + 1) * block letters are plain goes to N + 1 tables keyed on seed + one block each
        # (a site is reported by the table of its first identical block), the rest stay in the ordinary table.
        n_blocks = self.mismatches + 1
        block = (16 - self.wordsize) // n_blocks
        can_block = self.mismatches >= 1 and not self.iupac_mode and block >= 1
        blocks = can_block and self.wordsize + block >= BLOCK_MIN_KEY and 8 * n >= 4 ** self.wordsize
        env = os.environ.get("MPCR_SEED_BLOCKS")
        if env is not None and can_block:
            blocks = env not in ("0", "")
            if blocks and env != "1":      # tests: letters per block
                block = max(1, min(block, int(env)))
        if blocks:
            parts = max(1, -(-n // EXT_LINES_PER_TABLE))
        if (extend or blocks) and os.environ.get("MPCR_SEED_PARTS"):
            parts = max(1, int(os.environ["MPCR_SEED_PARTS"]))
        if blocks:
            self._be.check(lib.mpcr_ctx_set_seed_blocks(self._ctx, block, n_blocks, 1))
        else:
            self._be.check(lib.mpcr_ctx_set_seed_extension(self._ctx, w_ext if extend else 0, 1 if extend else 0))
        self._be.check(lib.mpcr_ctx_set_sampling(self._ctx, SAMPLE_WORDSIZE if role else 0, stride if role else 0, role))
        self._be.check(lib.mpcr_table_build(self._ctx, blob.ctypes.data, off.ctypes.data, pcr.ctypes.data, n,
                                            plut.ctypes.data, self._stream()))
        n_ext = parts * (n_blocks if blocks else 1)
        while len(self._ctx_exts) > n_ext:
            lib.mpcr_ctx_destroy(self._ctx_exts.pop())
        while len(self._ctx_exts) < n_ext:
            self._ctx_exts.append(self._new_ctx())
        for k, ctx in enumerate(self._ctx_exts):
            if blocks:
                self._be.check(lib.mpcr_ctx_set_seed_blocks(ctx, block, n_blocks, 2 + k // parts))
            else:
                self._be.check(lib.mpcr_ctx_set_seed_extension(ctx, w_ext, 2))
            self._be.check(lib.mpcr_ctx_set_sampling(ctx, SAMPLE_WORDSIZE if role else 0, stride if role else 0, role))
            self._be.check(lib.mpcr_ctx_set_table_part(ctx, k % parts, parts))
            self._be.check(lib.mpcr_table_build(ctx, blob.ctypes.data, off.ctypes.data, pcr.ctypes.data, n,
                                                plut.ctypes.data, self._stream()))
        ho = np.full(2 * n, -1, dtype=np.int32)
        hv = np.zeros(2 * n, dtype=np.uint32)
        self._be.check(lib.mpcr_table_records(self._ctx, ho.ctypes.data, hv.ctypes.data))
        # insertion order == record-slot order ("+" then "-" of each line, engine.py:265-281)
        inserted = ho >= 0
        rec_to_idx = np.full(2 * n, -1, dtype=np.int64)
        rec_to_idx[inserted] = np.arange(int(np.count_nonzero(inserted)), dtype=np.int64)
        self._rec_to_idx = rec_to_idx
        self._hash_offsets = ho
        self._hashes = hv
        self._sts_records = None                # host mirror of the records: built when somebody looks at it
        self._sts_table = None
        return int(2 * n - np.count_nonzero(inserted))

    @property
    def sts_records(self) -> List[STSRecord]:
        """The reference's record list (engine.py:253-281, insertion order), materialised lazily from the parsed lines
        and the hash offsets the device computed."""
        if self._sts_records is None:
            src, ho = self._sts_lines, self._hash_offsets
            recs: List[STSRecord] = []
            if src is not None and src.n:
                ids, aliases, p1s, p2s = src.ids, src.aliases, src.p1s, src.p2s
                sizes = [int(x) for x in src.sizes]
                line_nos = [int(x) for x in src.line_nos]
                hol = ho.tolist()
                for i in range(src.n):
                    if hol[2 * i] >= 0:
                        recs.append(STSRecord(id=ids[i], primer1=p1s[i],
                                              primer2=self._reverse_complement(p2s[i]) if self.true_strands else p2s[i],
                                              pcr_size=sizes[i],
                                              alias=aliases[i], offset=line_nos[i], hash_offset=hol[2 * i], direct="+"))
                    if hol[2 * i + 1] >= 0:
                        recs.append(STSRecord(id=ids[i], primer1=p2s[i], primer2=self._reverse_complement(p1s[i]),
                                              pcr_size=sizes[i], alias=aliases[i], offset=line_nos[i],
                                              hash_offset=hol[2 * i + 1], direct="-"))
            self._sts_records = recs
        return self._sts_records

    @sts_records.setter
    def sts_records(self, value):
        self._sts_records = value

    @property
    def sts_table(self) -> Dict[int, List[STSRecord]]:
        """hash value -> records in insertion order (engine.py:70,324-329), materialised from the device build."""
        if self._sts_table is None:
            table: Dict[int, List[STSRecord]] = {}
            for slot in np.flatnonzero(self._rec_to_idx >= 0).tolist():
                table.setdefault(int(self._hashes[slot]), []).append(self.sts_records[int(self._rec_to_idx[slot])])
            self._sts_table = table
        return self._sts_table

    @sts_table.setter
    def sts_table(self, value):
        self._sts_table = value

    def _parse_pcr_size(self, pcr_size_str: str) -> int:
        """engine.py:304-322."""
        if "-" in pcr_size_str:
            try:
                size_range = pcr_size_str.split("-")
                if len(size_range) == 2 and size_range[0] and size_range[1]:
                    return (int(size_range[0]) + int(size_range[1])) // 2
                return self.default_pcr_size
            except ValueError:
                return self.default_pcr_size
        try:
            pcr_size = int(pcr_size_str)
            return pcr_size if pcr_size > 0 else self.default_pcr_size
        except ValueError:
            return self.default_pcr_size

    # ------------------------------------------------------------------ helpers the reference's tests call
    def _hash_value(self, primer: str):
        """engine.py:331-355 (host mirror; the table itself is hashed on the device)."""
        primer = primer.upper()
        W = self.wordsize
        if len(primer) < W:
            return -1, 0
        code = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}
        run, h, mask = 0, 0, (1 << (2 * W)) - 1
        for i, ch in enumerate(primer):
            c = code.get(ch)
            if c is None:
                run = 0
                continue
            h = ((h << 2) | c) & mask
            run += 1
            if run >= W:
                return i - W + 1, h
        return -1, 0

    def _reverse_complement(self, sequence: str) -> str:
        """engine.py:357-359."""
        return sequence[::-1].translate(_COMPL)

    def _compare_seqs(self, seq1: str, seq2: str, strand: str) -> bool:
        """engine.py:599-642 (host mirror for API users; the search itself compares on the device)."""
        if len(seq1) != len(seq2):
            return False
        n, mism = len(seq1), 0
        for i in range(n):
            prot = (strand == "+" and i >= n - self.three_prime_match) or \
                   (strand == "-" and i < self.three_prime_match)
            c1, c2 = seq1[i].upper(), seq2[i].upper()
            if self.iupac_mode and c1 in _IUPAC and c2 in _IUPAC:
                match = bool(set(_IUPAC[c1]) & set(_IUPAC[c2]))
            else:
                match = c1 == c2
            if not match:
                if prot:
                    return False
                mism += 1
                if mism > self.mismatches:
                    return False
        return True

    def load_fasta_file(self, filename: str) -> List[FASTARecord]:
        """engine.py:361-363."""
        return FASTALoader.load_file(filename, engine=self)

    # ------------------------------------------------------------------ search (engine.py:365-451)
    def search(self, fasta_records: List[FASTARecord], output_file: str = None) -> int:
        """Search for STS markers in the provided FASTA sequences; prints one line per hit (engine.py:442)."""
        total_hits = 0
        # one rank of a multi-GPU run (multi.py): every rank scans its shard, rank 0 gathers, merges and writes
        gather = multi.active_world(self.shard)
        writer = not gather or self.shard[0] == 0
        if writer and output_file and output_file.lower() != "stdout":
            output = open(output_file, "w")
        else:
            output = sys.stdout
        try:
            hits = self.search_hits(fasta_records, copy=False) if fasta_records else np.zeros(0, dtype=_capi.HIT_DTYPE)
            if writer:
                for record in fasta_records:
                    logger.info(f"Processing sequence: {record.label} ({len(record)} bp)")
            total_hits = int(hits.size)
            if gather:
                hits, total_hits = multi.gather_hits(hits)
                if hits is None:
                    hits = np.zeros(0, dtype=_capi.HIT_DTYPE)
            src = self._sts_lines
            labels = [r.label for r in fasta_records]
            if hits.size and src is not None and src.lines is not None and all(lb.isascii() for lb in labels):
                # engine.py:437-444 in one native pass over the hit array (ids / aliases straight from the file text)
                lb = [x.encode("ascii") for x in labels]
                label_off = np.zeros(len(lb) + 1, dtype=np.uint64)
                label_off[1:] = np.cumsum([len(x) for x in lb])
                label_blob = np.frombuffer(b"".join(lb) + b"\0", dtype=np.uint8)
                hits = np.ascontiguousarray(hits)
                fmt = self._be.lib.mpcr_format_hits
                args = (hits.ctypes.data, hits.size, src.raw.ctypes.data, src.lines.ctypes.data, label_blob.ctypes.data,
                        label_off.ctypes.data)
                need = int(fmt(*args, None, 0))
                buf = np.empty(need, dtype=np.uint8)
                used = int(fmt(*args, buf.ctypes.data, need))
                output.write(buf[:used].tobytes().decode("ascii"))
            elif hits.size:
                bounds = np.searchsorted(hits["contig"], np.arange(len(fasta_records) + 1))
                recs = self.sts_records
                r2i = self._rec_to_idx
                for ci, record in enumerate(fasta_records):
                    h = hits[bounds[ci]: bounds[ci + 1]]
                    if h.size == 0:
                        continue
                    p1 = (h["pos1"].astype(np.int64) + 1).tolist()
                    p2 = (h["pos2"].astype(np.int64) + 1).tolist()
                    ri = r2i[h["rec"]].tolist()
                    lines = []
                    for a, b, r in zip(p1, p2, ri):
                        sts = recs[r]
                        lines.append(f"{record.label}\t{a}..{b}\t{sts.id}\t{sts.alias}\t({sts.direct})\n")
                    output.write("".join(lines))
        finally:
            if output is not sys.stdout:
                output.close()
        logger.info(f"Total hits found: {total_hits}")
        self.total_hits = total_hits
        return total_hits

    def search_hits(self, fasta_records: Sequence[FASTARecord], copy: bool = True) -> np.ndarray:
        """The device path of `search`: returns this shard's hits (structured array, _capi.HIT_DTYPE) in the
        reference's output order; `contig` indexes `fasta_records`, positions are 0-based inclusive."""
        t0 = time.perf_counter()
        if self._sts_lines is None:  # nothing loaded: an empty table, zero hits (like the reference)
            self._sts_lines = _STSLines(0, [], [], ids=[], aliases=[], p1s=[], p2s=[])
            self._build_table()
        # sequences the device-side ingest left in HBM are used in place; everything else is host bytes
        seqs = []
        for r in fasta_records:
            if hasattr(r, "piece"):             # rank-local ingest: only this rank's stretch of the record is here
                seqs.append(_Piece(*r.piece) if r.piece is not None else None)
            elif getattr(r, "sequence_device", None) is not None and r.sequence_device.device == self._tdev:
                seqs.append(r.sequence_device)
            else:
                seqs.append(r.sequence_bytes)
        self._check_alphabet(fasta_records, seqs)
        layout = self.make_layout([len(r) for r in fasta_records])
        owned = getattr(fasta_records, "owned_range", None)
        if owned is not None:                   # ... and the rank owns the positions its file range starts
            layout["begin"], layout["end"] = int(owned[0]), int(owned[1])
        _, hits_t, n = self.upload_and_scan(layout, seqs)
        hits = self._hits_to_host(hits_t, n, copy=copy)
        self.last_timing = dict(search_s=time.perf_counter() - t0)
        return hits

    # -- alphabet corner cases: sequences built through the API may hold letters FASTA files cannot
    def _check_alphabet(self, records, seqs):
        exotic = set()
        for r, s in zip(records, seqs):
            if not getattr(r, "_from_loader", False) and len(r):
                exotic |= genome_exotics(s if isinstance(s, np.ndarray) else r.sequence_bytes, self.iupac_mode)
        if not exotic or exotic == {"X"}:
            zero = "X"
        elif len(exotic) == 1:
            zero = next(iter(exotic))
        else:
            pex = primer_exotics(self._sts_lines.p1s + self._sts_lines.p2s, self.iupac_mode) if self._sts_lines else set()
            if pex & exotic:
                raise ValueError(
                    "sequences contain several non-IUPAC characters "
                    f"({''.join(sorted(exotic))}) that also occur in primers; the 4-bit device alphabet can "
                    "represent only one such character")
            zero = ""  # none of them can ever match a primer character
        if zero != self._zero_char and self._sts_lines is not None:
            self._zero_char = zero
            self._build_table()

    # -- layout: padded global coordinate (include/merpcr_b200.h)
    def make_layout(self, lengths: Sequence[int]) -> dict:
        n = len(lengths)
        contigs = np.zeros(n, dtype=_capi.CONTIG_DTYPE)
        g = 0
        for i, L in enumerate(lengths):
            if L >= (1 << 31):
                raise ValueError("sequences longer than 2^31-1 bases are not supported")
            contigs[i]["gstart"] = g
            contigs[i]["length"] = L
            g = (g + L + 1 + 127) // 128 * 128
        total = g
        rank, world = self.shard
        # bp-balanced shard boundaries on multiples of 128; a tile belongs to the shard holding its first base
        begin = (total * rank // world) // 128 * 128
        end = (total * (rank + 1) // world) // 128 * 128 if rank + 1 < world else max(total, 128)
        return dict(contigs=contigs, total=total, begin=begin, end=end, lengths=list(lengths))

    def _prepare_shard(self, layout: dict, shard: Optional[_Shard]) -> _Shard:
        """Plane buffers of this rank's range (+ halos), zeroed where that matters; re-uses `shard` when it fits."""
        lib = self._be.lib
        total, begin, end = layout["total"], layout["begin"], layout["end"]
        halo_l, halo_r = int(lib.mpcr_halo_left(self._ctx)), int(lib.mpcr_halo_right(self._ctx))
        origin = max(0, begin - halo_l) // 128 * 128
        stop = min(total, end + halo_r)
        bases = max(128, (stop - origin + 127) // 128 * 128)
        alloc = bases + int(lib.mpcr_tile_bases()) + PLANE_SLACK_BASES
        sh = shard
        contig_sig = layout["contigs"].tobytes()
        if sh is not None and (sh.origin, sh.bases, sh.begin, sh.end) == (origin, bases, begin, end):
            if sh.contig_sig != contig_sig:     # same extent, other contig boundaries: the gaps must be zero again
                sh.plane2.zero_()
                sh.plane4.zero_()
                sh.valid.zero_()
            # identical layout: every word that holds a base is overwritten below, the gaps are still zero
        else:
            sh = _Shard()
            sh.device, sh.origin, sh.bases, sh.begin, sh.end = self._tdev, origin, bases, begin, end
            sh.alloc, sh.last_n = alloc, 0     # alloc = bases the plane allocations hold (mpcr_scan checks it)
            sh.plane2 = torch.zeros(alloc // 4, dtype=torch.uint8, device=self._tdev)
            sh.plane4 = torch.zeros(alloc // 2, dtype=torch.uint8, device=self._tdev)
            sh.valid = torch.zeros(alloc // 8, dtype=torch.uint8, device=self._tdev)
            sh.hits, sh.count, sh.staging, sh.stage2 = None, None, None, None
        sh.contig_sig = contig_sig
        return sh

    @staticmethod
    def _pieces(layout: dict, seqs: Sequence, sh: _Shard, chunk: int):
        """(contig index, global begin, global end, source slice, last piece of its contig) for every stretch of bases
        this shard holds, in genome order."""
        for ci, s in enumerate(seqs):
            if s is None:
                continue
            g0 = int(layout["contigs"][ci]["gstart"])
            L = int(layout["contigs"][ci]["length"])
            lo, hi = max(g0, sh.origin), min(g0 + L, sh.origin + sh.bases)
            if isinstance(s, _Piece):           # a stretch of the contig: it must hold everything the shard touches
                if lo >= hi:
                    continue
                if lo < g0 + s.start or hi > g0 + s.start + s.data.numel():
                    raise RuntimeError(
                        f"rank-local FASTA ingest: contig {ci} is needed on [{lo - g0}, {hi - g0}) but only "
                        f"[{s.start}, {s.start + s.data.numel()}) was read here -- load the STS file before the FASTA "
                        "file (the halos depend on it) or set MPCR_RANK_LOCAL_INGEST=0")
                g0 += s.start                   # index the stretch by global coordinate
                s = s.data
            for a in range(lo, hi, chunk):
                b = min(hi, a + chunk)
                if isinstance(s, torch.Tensor):
                    src = s[a - g0: b - g0]
                else:
                    arr = s[a - g0: b - g0]
                    src = torch.from_numpy(arr if arr.flags.writeable else arr.copy())
                yield ci, a, b, src, b == hi

    def upload(self, layout: dict, seqs: Sequence, shard: Optional[_Shard] = None) -> _Shard:
        """FASTA ingest, device half: copy the bases this shard needs (its range + halos) to the GPU and pack them
        into the planes.  seqs[i] is the i-th contig as a uint8 numpy array, a torch uint8 tensor (pinned host or
        already on the device) or None for a contig this shard never touches.  Passing the previous `shard`
        re-uses its buffers (steady-state re-upload)."""
        lib = self._be.lib
        sh = self._prepare_shard(layout, shard)
        lut = genome_lut(self.iupac_mode)
        stream = self._stream()
        chunk = 1 << 28
        h2d = 0
        for ci, a, b, src, _ in self._pieces(layout, seqs, sh, chunk):
            if src.device != self._tdev:
                # one device staging buffer, reused: copy and pack are ordered on the same stream
                h2d += b - a
                if sh.staging is None or sh.staging.numel() < b - a:
                    sh.staging = torch.empty(min(chunk, max(b - a, 1 << 24)), dtype=torch.uint8, device=self._tdev)
                dst = sh.staging[: b - a]
                dst.copy_(src, non_blocking=True)
                src = dst
            self._be.check(lib.mpcr_pack_sequence(self._ctx, src.data_ptr(), b - a, a, sh.origin,
                                                  sh.plane2.data_ptr(), sh.plane4.data_ptr(), sh.valid.data_ptr(),
                                                  lut.ctypes.data, stream))
        self._sync()
        self.last_h2d_bytes = h2d
        return sh

    def upload_and_scan(self, layout: dict, seqs: Sequence, shard: Optional[_Shard] = None, sort: bool = True):
        """`upload` + `scan_device` as one pipeline for sequences that still sit in host memory: a copy stream moves
        64 MiB pieces into two staging buffers while the compute stream packs the previous piece, and every contig
        (group of small contigs) is scanned as soon as its bases are packed -- the tables append to one hit buffer
        (mpcr_ctx_set_append), the host reads the count once at the end, the hits are sorted once.  Host -> hits time
        is then the PCIe copy plus the last contig's scan.  Returns (shard, hit tensor, n_hits)."""
        if all(s is None or (isinstance(s, torch.Tensor) and s.device == self._tdev) or
               (isinstance(s, _Piece) and s.data.device == self._tdev) for s in seqs):
            sh = self.upload(layout, seqs, shard)
            hits, n = self.scan_device(layout, sh, sort=sort)
            return sh, hits, n
        lib = self._be.lib
        sh = self._prepare_shard(layout, shard)
        lut = genome_lut(self.iupac_mode)
        # (with the CPU test tier's emulated library everything below runs in program order: host memory IS device
        # memory there, so no piece is ever staged, and the range / append logic is exercised as is)
        gpu = self._tdev.type == "cuda"
        chunk = min(1 << 26, max(1 << 20, (sh.bases + 127) // 128 * 128))   # small genomes: small staging buffers
        compute = copy = None
        copied = packed = ()
        if gpu:
            compute = torch.cuda.current_stream(self._tdev)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(self._tdev)
            copy = self._copy_stream
            if sh.stage2 is None or sh.stage2[0].numel() < chunk:
                sh.stage2 = [torch.empty(chunk, dtype=torch.uint8, device=self._tdev) for _ in range(2)]
            copied = [torch.cuda.Event() for _ in range(2)]
            packed = [torch.cuda.Event() for _ in range(2)]
        stream = compute.cuda_stream if gpu else 0
        isz = _capi.HIT_DTYPE.itemsize
        if sh.count is None:
            sh.count = torch.zeros(1, dtype=torch.int64, device=self._tdev)
        if sh.hits is None:
            sh.hits = torch.empty((1 << 16) * isz, dtype=torch.uint8, device=self._tdev)
        cap = sh.hits.numel() // isz
        sh.count.zero_()
        if gpu:
            copy.wait_stream(compute)          # the staging buffers and the zeroed planes exist before the first copy
        contigs = layout["contigs"]
        ctxs = self._all_ctxs()
        nxt = [int(c["gstart"]) for c in contigs[1:]] + [max(layout["total"], sh.end)]

        def scan_range(lo: int, hi: int):
            for ctx in ctxs:
                self._be.check(lib.mpcr_scan(ctx, contigs.ctypes.data, len(contigs), sh.plane2.data_ptr(),
                                             sh.plane4.data_ptr(), sh.valid.data_ptr(), sh.origin, sh.alloc, lo, hi,
                                             sh.hits.data_ptr(), cap, sh.count.data_ptr(), stream))

        for ctx in ctxs:
            # work descriptors of the whole range go up now, while the copy engine is idle; the range scans below
            # are views into them
            self._be.check(lib.mpcr_scan_prepare(ctx, contigs.ctypes.data, len(contigs), sh.origin, sh.begin, sh.end,
                                                 stream))
            self._be.check(lib.mpcr_ctx_set_append(ctx, 1))
        # Host-resident pieces go over PCIe as packed nibbles (0.5 byte per base): the host cores pack them into pinned
        # staging buffers in plane4's own layout (mpcr_host_pack_nibbles), the copy lands directly in plane4, and the
        # device derives the other two planes (mpcr_derive_planes).  A piece the nibble path cannot carry (a 'U' outside
        # IUPAC mode) goes up as ASCII through the device staging buffers like before.
        use_nibbles = self.host_pack
        n_stage = 3
        hstage, hfree = None, ()
        if use_nibbles:
            if gpu:
                if self._host_stage is None or self._host_stage[0].numel() < chunk // 2:
                    self._host_stage = [torch.empty(chunk // 2, dtype=torch.uint8).pin_memory() for _ in range(n_stage)]
                hfree = [torch.cuda.Event() for _ in range(n_stage)]
            elif self._host_stage is None or self._host_stage[0].numel() < chunk // 2:
                self._host_stage = [torch.empty(chunk // 2, dtype=torch.uint8) for _ in range(n_stage)]
            hstage = self._host_stage
        try:
            k, kn, h2d, pending, done_to, deferred, pack_s = 0, 0, 0, 0, sh.begin, None, 0.0
            # Hybrid wire format (pinned sources only): the host cores pack at about the rate the link carries ASCII, so
            # neither route alone beats ~56 ms for a human genome.  Both together do: a piece is packed only while the copy
            # engine still has at least most of a pack's duration of work queued (packing then costs no wall clock and
            # halves the piece's bytes); when the engine is about to run dry the next piece goes up as ASCII at once.
            h2d_rate, busy_until = self._h2d_rate, time.perf_counter()
            for ci, a, b, src, last in self._pieces(layout, seqs, sh, chunk):
                on_host = src.device.type == "cpu"
                done = False
                pack_it = on_host and use_nibbles
                if pack_it and gpu and self.hybrid_wire and src.is_pinned():
                    backlog = busy_until - time.perf_counter()
                    pack_it = backlog >= self.hybrid_backlog * (b - a) / self._pack_rate
                if pack_it:
                    slot = kn % n_stage
                    if gpu and kn >= n_stage:
                        hfree[slot].synchronize()              # the copy that last read this staging buffer is done
                    nb = (b - a + 1) // 2
                    t_pack = time.perf_counter()
                    rc = int(lib.mpcr_host_pack_nibbles(src.data_ptr(), b - a, lut.ctypes.data, hstage[slot].data_ptr(),
                                                        self._pack_threads))
                    dt_pack = time.perf_counter() - t_pack
                    pack_s += dt_pack
                    if b - a >= (1 << 22):     # running estimate of the packer's rate (bases / s) on this host
                        self._pack_rate = 0.7 * self._pack_rate + 0.3 * (b - a) / max(dt_pack, 1e-6)
                    if rc < 0:
                        raise ValueError("mpcr_host_pack_nibbles: bad argument")
                    if rc == 0:
                        off = (a - sh.origin) // 2
                        if gpu:
                            with torch.cuda.stream(copy):
                                sh.plane4[off: off + nb].copy_(hstage[slot][:nb], non_blocking=True)
                            hfree[slot].record(copy)
                            busy_until = max(busy_until, time.perf_counter()) + nb / h2d_rate
                        else:
                            sh.plane4[off: off + nb].copy_(hstage[slot][:nb])
                        kn += 1
                        h2d += nb
                        if deferred:                           # the finished contig is scanned while this copy runs
                            scan_range(*deferred)
                            deferred = None
                        if gpu:
                            compute.wait_event(hfree[slot])
                        self._be.check(lib.mpcr_derive_planes(self._ctx, b - a, a, sh.origin, sh.plane4.data_ptr(),
                                                              sh.plane2.data_ptr(), sh.valid.data_ptr(), stream))
                        done = True
                if not done:
                    slot = -1
                    if src.device != self._tdev:
                        slot = k & 1
                        k += 1
                        buf = sh.stage2[slot][: b - a]
                        copy.wait_event(packed[slot])              # the pack that last read this buffer is done
                        if not src.is_pinned():
                            src = src.pin_memory()                 # keeps the copy asynchronous
                        with torch.cuda.stream(copy):
                            buf.copy_(src, non_blocking=True)
                        copied[slot].record(copy)
                        busy_until = max(busy_until, time.perf_counter()) + (b - a) / h2d_rate
                        h2d += b - a
                        src = buf
                    if deferred:                                   # the finished contig is scanned while this copy runs
                        scan_range(*deferred)
                        deferred = None
                    if slot >= 0:
                        compute.wait_event(copied[slot])
                    self._be.check(lib.mpcr_pack_sequence(self._ctx, src.data_ptr(), b - a, a, sh.origin,
                                                          sh.plane2.data_ptr(), sh.plane4.data_ptr(), sh.valid.data_ptr(),
                                                          lut.ctypes.data, stream))
                    if slot >= 0:
                        packed[slot].record(compute)
                pending += b - a
                # a finished contig (or run of small ones) whose right neighbourhood is complete can be scanned
                if last and pending >= STREAM_SCAN_BASES and done_to < nxt[ci] < sh.end:
                    deferred = (done_to, nxt[ci])
                    done_to, pending = nxt[ci], 0
            if deferred:
                scan_range(*deferred)
            if sh.end > done_to:
                scan_range(done_to, sh.end)
            if sort:    # queued behind the last scan: the count is read on the device
                self._be.check(lib.mpcr_sort_hits_dev(self._ctx, sh.hits.data_ptr(), sh.count.data_ptr(), cap,
                                                      sh.last_n, stream))
            need = int(sh.count.item())                        # the one host round trip
            self.last_scan_ms = float(lib.mpcr_last_scan_ms(self._ctx))
        finally:
            for ctx in ctxs:
                lib.mpcr_ctx_set_append(ctx, 0)
        self.last_h2d_bytes = h2d
        self.last_timing = dict(host_pack_s=pack_s, host_pack_threads=self._pack_threads)
        if need > cap:      # the hit list outgrew the buffer: the planes are resident now, scan them again with room
            hits, n = self.scan_device(layout, sh, sort=sort)
            return sh, hits, n
        sh.last_n = need
        return sh, sh.hits, need

    def scan(self, layout: dict, sh: _Shard, sort: bool = True) -> np.ndarray:
        """scanner + verifier + hit emitter + ordering on the resident planes; returns the hits on the host."""
        hits, n = self.scan_device(layout, sh, sort=sort)
        return self._hits_to_host(hits, n)

    def _hits_to_host(self, hits, n: int, copy: bool = True) -> np.ndarray:
        """The first n hit records as a host array.  copy=False returns a view of the engine's pinned staging buffer,
        valid until the next call (what `search` formats its output from)."""
        if n == 0:
            return np.zeros(0, dtype=_capi.HIT_DTYPE)
        nb = n * _capi.HIT_DTYPE.itemsize
        if hits.device.type != "cuda":
            return hits[:nb].numpy().view(_capi.HIT_DTYPE).copy()
        # through a pinned staging buffer: a pageable D2H of a few MB costs more than the copy itself
        if self._pinned_hits is None or self._pinned_hits.numel() < nb:
            self._pinned_hits = torch.empty(max(nb, 1 << 22), dtype=torch.uint8).pin_memory()
        stage = self._pinned_hits[:nb]
        stage.copy_(hits[:nb], non_blocking=True)
        self._sync()
        out = stage.numpy().view(_capi.HIT_DTYPE)
        return out.copy() if copy else out

    def scan_device_async(self, layout: dict, sh: _Shard, slot: int = 0, sort: bool = True):
        """One step (scan + verify + ordering) queued WITHOUT waiting for it: for callers that keep several steps in
        flight -- the host reads step k's result (`scan_finish`) while the next ones run.  Consecutive steps share the
        stream and the contexts' scratch; each slot (0 .. _capi.MAX_SLOTS - 1) has its own hit buffer, count and pinned
        result words.  Returns a handle for `scan_finish`."""
        import ctypes as C
        lib = self._be.lib
        contigs = layout["contigs"]
        isz = _capi.HIT_DTYPE.itemsize
        if self._slots is None:
            self._slots = [dict(hits=None, count=None, result=None, event=None, last_n=0) for _ in range(_capi.MAX_SLOTS)]
        st = self._slots[slot]
        gpu = self._tdev.type == "cuda"
        if st["count"] is None:
            st["count"] = torch.zeros(1, dtype=torch.int64, device=self._tdev)
            st["result"] = torch.zeros(2, dtype=torch.int64)
            if gpu:
                st["result"] = st["result"].pin_memory()
                st["event"] = torch.cuda.Event()
        if st["hits"] is None:
            cap0 = 1 << 16 if sh.hits is None else sh.hits.numel() // isz
            st["hits"] = torch.empty(cap0 * isz, dtype=torch.uint8, device=self._tdev)
        cap = st["hits"].numel() // isz
        ctxs = self._all_ctxs()
        arr = (C.c_void_p * len(ctxs))(*[c.value if hasattr(c, "value") else c for c in ctxs])
        stream = self._stream()
        self._be.check(lib.mpcr_scan_sorted_async(arr, len(ctxs), contigs.ctypes.data, len(contigs), sh.plane2.data_ptr(),
                                                  sh.plane4.data_ptr(), sh.valid.data_ptr(), sh.origin, sh.alloc, sh.begin,
                                                  sh.end, st["hits"].data_ptr(), cap, st["count"].data_ptr(),
                                                  st["result"].data_ptr(), st["last_n"], 1 if sort else 0, slot, stream))
        if gpu:
            st["event"].record(torch.cuda.current_stream(self._tdev))
        return (slot, cap, sort)

    def scan_finish(self, layout: dict, sh: _Shard, handle):
        """Wait for the step `scan_device_async` queued and return (hit tensor, n_hits) like `scan_device`.  A hit list
        that outgrew the slot's buffer is scanned again with room (synchronously; nothing is ever truncated)."""
        slot, cap, sort = handle
        lib = self._be.lib
        st = self._slots[slot]
        if st["event"] is not None:
            st["event"].synchronize()
        need = int(st["result"][0])
        if need > cap:
            isz = _capi.HIT_DTYPE.itemsize
            st["hits"] = torch.empty(max(need, 2 * cap) * isz, dtype=torch.uint8, device=self._tdev)
            return self.scan_finish(layout, sh, self.scan_device_async(layout, sh, slot, sort))
        if sort and int(st["result"][1]):
            self._be.check(lib.mpcr_sort_finish(self._ctx, st["hits"].data_ptr(), st["result"].data_ptr(), cap, self._stream()))
        st["last_n"] = need
        return st["hits"], need

    def scan_device(self, layout: dict, sh: _Shard, sort: bool = True):
        """Device-resident scan: returns (uint8 tensor holding mpcr_hit records, n_hits).  Re-runs with a larger
        buffer if the hit list outgrows it (nothing is ever truncated)."""
        lib = self._be.lib
        contigs = layout["contigs"]
        if sh.count is None:
            sh.count = torch.zeros(1, dtype=torch.int64, device=self._tdev)
        isz = _capi.HIT_DTYPE.itemsize
        cap = 1 << 16 if sh.hits is None else sh.hits.numel() // isz
        import ctypes as C
        ctxs = self._all_ctxs()
        stream = self._stream()
        arr = (C.c_void_p * len(ctxs))(*[c.value if hasattr(c, "value") else c for c in ctxs])
        if self._count_host is None:    # where the one 8-byte read-back of a step lands
            self._count_host = torch.zeros(1, dtype=torch.int64)
            if self._tdev.type == "cuda":
                self._count_host = self._count_host.pin_memory()
        while True:
            if sh.hits is None or sh.hits.numel() < cap * isz:
                sh.hits = torch.empty(cap * isz, dtype=torch.uint8, device=self._tdev)
            # one call per step: every table's scan + verify, the sort (count read on the device) and the read-back of
            # the count, queued back to back; returns when the stream has drained
            self._be.check(lib.mpcr_scan_sorted(arr, len(ctxs), contigs.ctypes.data, len(contigs), sh.plane2.data_ptr(),
                                                sh.plane4.data_ptr(), sh.valid.data_ptr(), sh.origin, sh.alloc, sh.begin,
                                                sh.end, sh.hits.data_ptr(), cap, sh.count.data_ptr(),
                                                self._count_host.data_ptr(), sh.last_n, 1 if sort else 0, stream))
            need = int(self._count_host[0])
            if need <= cap:
                break
            cap = max(need, 2 * cap)        # nothing is ever truncated: scan again with room
        self.last_scan_ms = sum(float(lib.mpcr_last_scan_ms(ctx)) for ctx in ctxs)
        sh.last_n = need
        return sh.hits, need
