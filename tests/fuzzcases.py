"""Seeded small-case generator for golden vectors (test infrastructure).

make_case(seed) -> dict(params=..., sts_text=str, fasta_text=str).  The golden file stores the generated
inputs inline next to the reference's output, so parity tests never depend on regenerating them; the
generator is committed so the vectors can be re-made (tests/golden/make_golden.py).

The cases are built to hit every rule of SURVEY.md Appendix A: both strand conventions (Q1), end-of-sequence
clamp (Q5), ambiguity letters in sequence and primers (Q6/Q7), hash_offset > 0, duplicate STS lines, range
sizes (Q8), unparsable sizes, short primers, comments / blank lines / CRLF, junk characters and lower case in
FASTA text (Q2), contigs with len <= W (Q3), multi-delta hits, 3' protection on both strands.
"""
from __future__ import annotations

from synth import Rng, revcomp_bytes, ACGT
import numpy as np

IUPAC_AMBIG = "RYMKSWBDHVN"


def _mutate(r: Rng, s: str, nsub: int, lo: int, hi: int) -> str:
    """substitute nsub positions in s[lo:hi] with a different base"""
    b = list(s)
    for _ in range(nsub):
        if hi <= lo:
            break
        j = r.randint(lo, hi - 1)
        b[j] = r.choice([c for c in "ACGT" if c != b[j].upper()])
    return "".join(b)


def _rc(s: str) -> str:
    return revcomp_bytes(np.frombuffer(s.encode(), dtype=np.uint8)).tobytes().decode()


def make_case(seed: int) -> dict:
    r = Rng(seed * 7919 + 13)
    W = r.choice([3, 4, 5, 8, 8, 11, 11, 11, 12, 16])
    N = r.choice([0, 0, 1, 1, 2, 3])
    X = r.choice([0, 1, 1, 1, 2, 5, 30])
    M = r.choice([0, 3, 10, 50, 50])
    I = r.choice([0, 0, 1])
    Z = r.choice([240, 240, 100])
    small_w = W <= 5
    n_contigs = r.randint(1, 3)
    contigs = []
    for ci in range(n_contigs):
        if r.chance(0.12):
            L = r.randint(0, W + 2)            # Q3 territory
        elif small_w:
            L = r.randint(40, 400)
        else:
            L = r.randint(200, 5000)
        contigs.append(bytearray(r.dna(L).tobytes()))

    n_sts = r.randint(1, 6 if small_w else 30)
    lines = []
    sts_defs = []
    for si in range(n_sts):
        lmin = max(W, 8) if r.chance(0.9) else max(3, W - 3)
        l1 = r.randint(lmin, lmin + 14)
        l2 = r.randint(lmin, lmin + 14)
        p1 = r.dna(l1).tobytes().decode()
        p2 = r.dna(l2).tobytes().decode()
        size = r.randint(l1 + l2 - 4, 60 if small_w else 400)
        sts_defs.append([f"S{si}", p1, p2, size, f"alias {si}" if r.chance(0.7) else ""])

    # plant amplicons into the contigs (before decorating primers with ambiguity codes)
    for si, (sid, p1, p2, size, alias) in enumerate(sts_defs):
        for _ in range(r.choice([0, 1, 1, 1, 2])):
            ci = r.randint(0, n_contigs - 1)
            seq = contigs[ci]
            eff = max(size, len(p1) + len(p2))
            d = r.randint(-M - 3, M + 3)
            prod = max(len(p1) + len(p2), eff + d)
            if r.chance(0.5):
                left, right = p1, p2
            else:
                left, right = p2, _rc(p1)
            nl = r.choice([0, 0, 0, 1, 1, 2])
            nr = r.choice([0, 0, 0, 1, 1, 2])
            left_m = _mutate(r, left, nl, 0, len(left))
            right_m = _mutate(r, right, nr, 0, len(right))
            if len(seq) < prod + 2:
                # place a truncated product against the end (Q5 clamp)
                prod2 = len(p1) + len(p2) + r.randint(0, 30)
                if len(seq) < prod2:
                    continue
                off = len(seq) - prod2 - (r.randint(0, 40) if r.chance(0.5) else 0)
                if off < 0:
                    continue
                prod = prod2
            elif r.chance(0.15):
                off = len(seq) - prod - r.randint(0, min(M + 5, len(seq) - prod))   # near the end
            else:
                off = r.randint(0, len(seq) - prod)
            seq[off: off + len(left_m)] = left_m.encode()
            seq[off + prod - len(right_m): off + prod] = right_m.encode()
            if r.chance(0.1) and off + prod + 3 + len(right_m) <= len(seq):
                # a second copy of the right primer a few bases later -> multi-delta hits
                seq[off + prod + 3: off + prod + 3 + len(right_m)] = right_m.encode()

    # decorate sequence: N runs, IUPAC letters, X, lower case
    for seq in contigs:
        L = len(seq)
        if L < 10:
            continue
        if r.chance(0.4):
            for _ in range(r.randint(1, 3)):
                a = r.randint(0, L - 1)
                n = r.randint(1, min(40, L - a))
                seq[a: a + n] = b"N" * n
        if r.chance(0.35):
            for _ in range(r.randint(1, 6)):
                seq[r.randint(0, L - 1)] = ord(r.choice(IUPAC_AMBIG + "X"))
        if r.chance(0.3):
            a = r.randint(0, L - 1)
            n = r.randint(1, L - a)
            seq[a: a + n] = bytes(seq[a: a + n]).lower()

    # decorate primers / STS lines
    for si, d in enumerate(sts_defs):
        sid, p1, p2, size, alias = d
        if r.chance(0.2):
            j = r.randint(0, len(p1) - 1)
            p1 = p1[:j] + r.choice(IUPAC_AMBIG) + p1[j + 1:]
        if r.chance(0.2):
            j = r.randint(0, len(p2) - 1)
            p2 = p2[:j] + r.choice(IUPAC_AMBIG) + p2[j + 1:]
        if r.chance(0.05):
            j = r.randint(0, len(p1) - 1)
            p1 = p1[:j] + r.choice("XUZ-") + p1[j + 1:]
        if r.chance(0.05):
            j = r.randint(0, len(p2) - 1)
            p2 = p2[:j] + r.choice("XUZ*") + p2[j + 1:]
        if r.chance(0.1):
            p1 = p1.lower()
        t = r.randint(0, 19)
        if t == 0:
            size_s = f"{size - 10}-{size + 11}"
        elif t == 1:
            size_s = r.choice(["abc", "-100", "0", "", "12-", "1-2-3", " 150 ", "1_50", "+200", "1e3"])
        else:
            size_s = str(size)
        fields = [sid, p1, p2, size_s]
        if alias or r.chance(0.3):
            fields.append(alias)
            if r.chance(0.1):
                fields.append("extra field")
        lines.append("\t".join(fields))
        if r.chance(0.08):
            lines.append("\t".join(fields))                    # duplicate line
        if r.chance(0.05):
            lines.append("\t".join([sid + "b", p1, p2, str(size + 7), "same primers"]))
    if r.chance(0.2):
        lines.insert(r.randint(0, len(lines)), "# a comment line")
    if r.chance(0.2):
        lines.insert(r.randint(0, len(lines)), "")
    if r.chance(0.03):
        lines.insert(r.randint(0, len(lines)), "bad\tline")    # < 4 fields -> load fails (Q10)
    nl = "\r\n" if r.chance(0.1) else "\n"
    sts_text = nl.join(lines) + (nl if r.chance(0.8) else "")

    fa = []
    if r.chance(0.1):
        fa.append("ACGTACGT")                                  # data before the first header is dropped
    for ci, seq in enumerate(contigs):
        fa.append(r.choice([f">c{ci} description text", f">c{ci}", f">  c{ci}\tmore", f" >c{ci} x"]))
        s = bytes(seq).decode()
        width = r.choice([60, 70, 13, 1000000])
        chunks = [s[i: i + width] for i in range(0, len(s), width)]
        for ch in chunks:
            if r.chance(0.05):
                j = r.randint(0, len(ch))
                ch = ch[:j] + r.choice(["1", " ", "*", "U", "u", "-", "Z"]) + ch[j:]   # filtered out (Q2)
            fa.append(ch)
            if r.chance(0.03):
                fa.append("")
    fnl = "\r\n" if r.chance(0.1) else "\n"
    fasta_text = fnl.join(fa) + fnl
    return dict(
        seed=seed,
        params=dict(wordsize=W, margin=M, mismatches=N, three_prime_match=X, iupac_mode=I, default_pcr_size=Z),
        sts_text=sts_text,
        fasta_text=fasta_text,
    )
