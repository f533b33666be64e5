"""merpcr_b200 -- B200-native STS search (electronic PCR) with the API surface of FOI-Bioinformatics/merpcr.

    from merpcr_b200 import MerPCR
    eng = MerPCR(wordsize=11, margin=50, mismatches=1)
    eng.load_sts_file("markers.sts"); recs = eng.load_fasta_file("genome.fa"); eng.search(recs, "hits.txt")

The hot path (FASTA packing, primer word-hash table, scanner, verifier, hit emitter) is hand-written CUDA for
sm_100a behind the C ABI in include/merpcr_b200.h; there is no CPU fallback.
"""
__version__ = "1.0.0"
__author__ = "merpcr_b200 contributors"
__license__ = "GPL-3.0"

from .engine import MerPCR  # noqa: E402
from .models import FASTARecord, STSHit, STSRecord  # noqa: E402

__all__ = ["MerPCR", "STSRecord", "FASTARecord", "STSHit"]
