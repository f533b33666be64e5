"""Loaders for the committed golden vectors (tests/golden/, made by tests/golden/make_golden.py)."""
import gzip
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURE_STS = os.path.join(GOLDEN_DIR, "fixture", "test.sts")
FIXTURE_FA = os.path.join(GOLDEN_DIR, "fixture", "test.fa")
# the reference's own end-to-end golden line (reference tests/test_comprehensive.py:65-95,193-220)
FIXTURE_LINE = "L78833\t75823..76023\tAFM248yg9\t(D17S932)  Chr.17, 63.7 cM\t(-)\n"

_fuzz = None


def fuzz_cases():
    global _fuzz
    if _fuzz is None:
        with gzip.open(os.path.join(GOLDEN_DIR, "fuzz_golden.json.gz"), "rb") as f:
            _fuzz = json.loads(f.read().decode())
    return _fuzz


def threaded_cases():
    with open(os.path.join(GOLDEN_DIR, "threaded_golden.json")) as f:
        return json.load(f)
