"""Deterministic synthetic inputs for parity tests and the benchmark (test/bench infrastructure).

Everything is driven by a counter-based splitmix64 stream implemented with numpy uint64 wrap-around
arithmetic, so the bytes produced for a given seed do not depend on the numpy / Python version.

Planting follows the reference's strand convention (engine.py:267,276-278):
  "+" amplicon = primer1 + filler + primer2          (primer2 literal, NOT reverse-complemented)
  "-" amplicon = primer2 + filler + revcomp(primer1)
"""
from __future__ import annotations

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, start: int, n: int) -> np.ndarray:
    """n 64-bit outputs of splitmix64 seeded with `seed`, for counters start .. start+n-1."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + 1 + n, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


class Rng:
    """Small sequential façade over the counter-based stream."""

    def __init__(self, seed: int):
        self.seed = int(seed)
        self.ctr = 0

    def u64(self, n: int) -> np.ndarray:
        out = splitmix64(self.seed, self.ctr, n)
        self.ctr += n
        return out

    def randint(self, lo: int, hi: int) -> int:
        """uniform integer in [lo, hi] (inclusive)."""
        return lo + int(self.u64(1)[0] % np.uint64(hi - lo + 1))

    def ints(self, lo: int, hi: int, n: int) -> np.ndarray:
        return (self.u64(n) % np.uint64(hi - lo + 1)).astype(np.int64) + lo

    def chance(self, p: float) -> bool:
        return float(self.u64(1)[0] >> np.uint64(11)) / float(1 << 53) < p

    def choice(self, seq):
        return seq[self.randint(0, len(seq) - 1)]

    def dna(self, n: int) -> np.ndarray:
        """n iid uniform bases as ASCII bytes (uint8)."""
        if n <= 0:
            return np.zeros(0, dtype=np.uint8)
        words = self.u64((n + 31) // 32)
        shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
        codes = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n]
        return ACGT[codes]


ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in ("AT", "CG", "GC", "TA", "UA", "BV", "DH", "HD", "KM", "MK", "NN", "RY", "SS", "VB", "WW", "XX", "YR"):
    _COMP[ord(_a)] = ord(_b)
    _COMP[ord(_a.lower())] = ord(_b.lower())


def revcomp_bytes(b: np.ndarray) -> np.ndarray:
    return _COMP[b[::-1]]


def dna_chunked(seed: int, n: int, chunk: int = 1 << 24) -> np.ndarray:
    """n iid bases for large n without large temporaries; stream position = base index / 32."""
    out = np.empty(n, dtype=np.uint8)
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        w0, w1 = s // 32, (e + 31) // 32
        words = splitmix64(seed, w0, w1 - w0)
        codes = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)
        out[s:e] = ACGT[codes[s - w0 * 32: e - w0 * 32]]
    return out


# ---------------------------------------------------------------------------------------------
# Planted-amplicon workloads (BASELINE.json configs 2-5, scaled by the caller)
# ---------------------------------------------------------------------------------------------

GRCH38_LENGTHS = [
    248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
    133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
    58617616, 64444167, 46709983, 50818468, 156040895, 57227415,
]
GRCH38_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]


def make_sts_set(seed: int, n_sts: int, len_lo: int = 18, len_hi: int = 25, size_lo: int = 100,
                 size_hi: int = 1000):
    """Random STS table: returns dict of arrays (primer bytes are rows of a padded uint8 matrix)."""
    r = Rng(seed)
    l1 = r.ints(len_lo, len_hi, n_sts)
    l2 = r.ints(len_lo, len_hi, n_sts)
    size = r.ints(size_lo, size_hi, n_sts)
    p1 = r.dna(n_sts * len_hi).reshape(n_sts, len_hi)
    p2 = r.dna(n_sts * len_hi).reshape(n_sts, len_hi)
    return dict(l1=l1, l2=l2, size=size, p1=p1, p2=p2)


def sts_lines(sts, id_fmt: str = "STS%06d", ranged: bool = False) -> bytes:
    """STS file text for a make_sts_set() table."""
    out = []
    for i in range(len(sts["l1"])):
        a = sts["p1"][i, : sts["l1"][i]].tobytes().decode()
        b = sts["p2"][i, : sts["l2"][i]].tobytes().decode()
        s = int(sts["size"][i])
        size = f"{s - 20}-{s + 20}" if ranged else str(s)
        out.append(f"{id_fmt % i}\t{a}\t{b}\t{size}\tsyn {i}\n")
    return "".join(out).encode()


def plant_amplicons(seed: int, contigs, sts, margin: int, sub_mode: str = "none", x_protect: int = 1,
                    plant_count=None):
    """Plant each STS once, in place, at non-overlapping slots spread over the contigs.

    contigs: list of writable uint8 arrays.  sub_mode:
      "none"  : exact primers; product = size + d, d in [-margin, margin] (90%) or +-[margin+1, margin+20] (10%, negatives)
      "cfg3"  : d as above for all; primers 45% exact, 45% one substitution outside the protected 3' base,
                5% one substitution ON the protected base (negative), 5% two substitutions (negative at N=1)
    If `contigs` is a list of LENGTHS the genome is not touched and (expected, writes) is returned, writes being
    (contig_index, offset, bytes) triples to apply in order (used to plant into device-resident genomes).
    Returns a list of expected hits (contig_index, pos1, pos2, sts_index, strand) for plants that must be
    found at the given margin (and N>=1 for cfg3), as a cross-check independent of any implementation.
    """
    writes = None
    if contigs and not hasattr(contigs[0], "__len__"):
        # planning mode: `contigs` is a list of lengths; the writes are returned instead of applied
        lengths, writes = [int(x) for x in contigs], []
    else:
        lengths = [len(c) for c in contigs]
    r = Rng(seed)
    n = len(sts["l1"]) if plant_count is None else min(int(plant_count), len(sts["l1"]))   # plant the first n STS
    total = sum(lengths)
    slot = total // max(n, 1)
    max_span = int(sts["size"][:n].max()) + margin + 64 if n else 0
    if slot < max_span + 8:
        raise ValueError("genome too small for non-overlapping plants")
    # global slot start -> (contig, offset); skip slots that straddle a contig end
    starts = np.cumsum([0] + lengths)
    jitter = r.ints(0, slot - max_span - 1, n)
    strand = r.u64(n) & np.uint64(1)
    d_in = r.ints(-margin, margin, n)
    d_out = r.ints(margin + 1, margin + 20, n) * (1 - 2 * (r.u64(n) & np.uint64(1)).astype(np.int64))
    neg = (r.u64(n) % np.uint64(10)) == 0
    kind = r.u64(n) % np.uint64(100)
    subpos = r.u64(n)
    subwhich = r.u64(n)
    expected = []
    for i in range(n):
        g = i * slot + int(jitter[i])
        ci = int(np.searchsorted(starts, g, side="right") - 1)
        off = g - int(starts[ci])
        L1, L2, size = int(sts["l1"][i]), int(sts["l2"][i]), int(sts["size"][i])
        size = max(size, L1 + L2)
        d = int(d_out[i]) if neg[i] else int(d_in[i])
        prod = size + d
        if prod < L1 + L2:
            prod, d = L1 + L2, L1 + L2 - size
        if off + prod > lengths[ci]:
            continue
        a = sts["p1"][i, :L1].copy()
        b = sts["p2"][i, :L2].copy()
        if strand[i] == 0:
            left, right, left_is_p1 = a, b, True            # p1 ... p2          -> "+"
        else:
            left, right, left_is_p1 = b, revcomp_bytes(a), False   # p2 ... rc(p1)  -> "-"
        found = abs(d) <= margin
        if sub_mode == "cfg3":
            k = int(kind[i])
            tgt_left = bool(subwhich[i] & np.uint64(1))
            arr = left if tgt_left else right
            ln = len(arr)
            # protected zone: left primer (strand "+" compare) = last x bases; right primer ("-") = first x
            free_idx = list(range(0, ln - x_protect)) if tgt_left else list(range(x_protect, ln))
            prot_idx = list(range(ln - x_protect, ln)) if tgt_left else list(range(0, x_protect))
            # never touch the reference's hash word (first W-mer of the LEFT primer) -- callers use W<=11, primers>=18,
            # so restrict left-primer substitutions to index >= 11
            if tgt_left:
                free_idx = [j for j in free_idx if j >= 11]
            def sub(j):
                arr[j] = ACGT[(int(np.searchsorted(ACGT, arr[j])) + 1 + int(subpos[i] % np.uint64(3))) % 4]
            if 45 <= k < 90 and free_idx:
                sub(free_idx[int(subpos[i] >> np.uint64(8)) % len(free_idx)])
            elif 90 <= k < 95 and prot_idx:
                sub(prot_idx[0]); found = False
            elif k >= 95 and len(free_idx) >= 2:
                j0 = int(subpos[i] >> np.uint64(8)) % len(free_idx)
                j1 = (j0 + 1 + int(subpos[i] >> np.uint64(20)) % (len(free_idx) - 1)) % len(free_idx)
                sub(free_idx[j0]); sub(free_idx[j1]); found = False
        if writes is None:
            seq = contigs[ci]
            seq[off: off + len(left)] = left
            seq[off + prod - len(right): off + prod] = right
        else:
            writes.append((ci, off, left))
            writes.append((ci, off + prod - len(right), right))
        if found:
            expected.append((ci, off, off + prod - 1, i, "+" if left_is_p1 else "-"))
    return expected if writes is None else (expected, writes)


# ---------------------------------------------------------------------------------------------
# The same generator on a torch device (bench.py builds the 3.1 Gbp genome directly in HBM)
# ---------------------------------------------------------------------------------------------

def _i64(x: int) -> int:
    x &= 0xFFFFFFFFFFFFFFFF
    return x - (1 << 64) if x >= (1 << 63) else x


def dna_torch(seed: int, start: int, n: int, device, chunk: int = 1 << 26):
    """Bases [start, start+n) of the stream dna_chunked(seed, ...) as a uint8 tensor on `device`
    (bit-identical to the numpy generator; int64 arithmetic wraps like uint64)."""
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    shifts = (torch.arange(32, dtype=torch.int64, device=device) * 2)[None, :]

    def lsr(z, s):
        return (z >> s) & ((1 << (64 - s)) - 1)

    for s in range(start, start + n, chunk):
        e = min(start + n, s + chunk)
        w0, w1 = s // 32, (e + 31) // 32
        idx = torch.arange(w0 + 1, w1 + 1, dtype=torch.int64, device=device)
        z = idx * _i64(0x9E3779B97F4A7C15) + _i64(seed)
        z = (z ^ lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * _i64(0x94D049BB133111EB)
        z = z ^ lsr(z, 31)
        codes = ((z[:, None] >> shifts) & 3).reshape(-1)
        out[s - start: e - start] = lut[codes[s - w0 * 32: e - w0 * 32]]
    return out


def block_table_case(seed: int, wordsize: int, mismatches: int, n_sts: int = 3000, contig_lens=(70000, 30000, 11),
                     margin: int = 30, plant_count: int = 90):
    """A workload for the block tables of searches that allow mismatches (mpcr_ctx_set_seed_blocks): exact plants whose
    LEFT primer site is then damaged at chosen offsets behind the seed -- one substitution walking through the first
    block, the second block and the letters behind them (two for -N 2), a genome N inside a block, plus STS lines that
    cannot be keyed by blocks (short primers, an ambiguity code in a block) and a duplicated line.
    Returns (contigs, sts_text, expected_planted_hits)."""
    rng = Rng(seed)
    contigs = [rng.dna(n) for n in contig_lens]
    sts = make_sts_set(seed + 1, n_sts, wordsize + 1, 26, 80, 400)
    sts["p1"][::11, wordsize + 2] = ord("N")          # ambiguity code right behind the seed: not blockable
    sts["p2"][5::13, wordsize + 1] = ord("R")
    expected = plant_amplicons(seed + 2, contigs[:2], sts, margin, plant_count=plant_count)
    for n, (ci, pos1, _pos2, i, strand) in enumerate(expected):
        seq = contigs[ci]
        ln = int(sts["l1"][i] if strand == "+" else sts["l2"][i])
        kind = n % 4
        offs = [wordsize + (n // 4) % 10]
        if mismatches >= 2 and kind == 1:
            offs.append(wordsize + (n // 4 + 3) % 10)
        if kind == 3:
            continue                                   # left exact
        for o in offs:
            if o < ln:
                if kind == 2:
                    seq[pos1 + o] = ord("N")           # a genome base that equals nothing
                else:
                    seq[pos1 + o] = ACGT[(int(np.searchsorted(ACGT, seq[pos1 + o])) + 1 + n % 3) % 4]
    text = sts_lines(sts)
    lines = text.split(b"\n")
    text = b"\n".join(lines[:-1] + [lines[3].replace(b"STS000003", b"DUP000003"), b""])   # same primers, another id
    return contigs, text, expected
