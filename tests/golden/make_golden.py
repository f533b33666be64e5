#!/usr/bin/env python3
"""Regenerate the golden vectors by running the UNMODIFIED reference (pure Python) in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (absent on the GPU box)

Writes
  tests/golden/fuzz_golden.json.gz   seeded small cases: inputs inline + the reference's output text, load
                                     status, STS record summary and FASTA record summary
  tests/golden/threaded_golden.json  >=100 kbp cases run with threads in {1,2,4}: regenerated from seeds by
                                     tests/synth.py (sha256 of the inputs recorded), reference output text inline
The reference is imported from /root/reference/src; nothing from it is copied into this repository.
"""
import gzip
import hashlib
import io
import json
import logging
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/src")

from merpcr import MerPCR  # noqa: E402  (the reference)
import fuzzcases  # noqa: E402
import synth  # noqa: E402

logging.disable(logging.CRITICAL)
N_FUZZ = 600


def run_reference(params, sts_text, fasta_text, threads=1):
    with tempfile.TemporaryDirectory() as d:
        sp, fp, op = os.path.join(d, "in.sts"), os.path.join(d, "in.fa"), os.path.join(d, "out.txt")
        with open(sp, "w", newline="") as f:
            f.write(sts_text)
        with open(fp, "w", newline="") as f:
            f.write(fasta_text)
        eng = MerPCR(threads=threads, **params)
        res = dict(load_ok=None, error=None, records=[], max_pcr_size=0, fasta=[], hits=0, output="")
        try:
            res["load_ok"] = bool(eng.load_sts_file(sp))
        except Exception as e:  # noqa: BLE001
            res["error"] = "sts:" + type(e).__name__
            return res
        res["records"] = [[r.id, r.direct, r.hash_offset, r.pcr_size, r.offset, r.primer1, r.primer2, r.alias]
                          for r in eng.sts_records]
        res["max_pcr_size"] = eng.max_pcr_size
        if not res["load_ok"]:
            return res
        try:
            recs = eng.load_fasta_file(fp)
        except Exception as e:  # noqa: BLE001
            res["error"] = "fasta:" + type(e).__name__
            return res
        res["fasta"] = [[r.label, len(r.sequence), hashlib.sha256(r.sequence.encode()).hexdigest()[:16]] for r in recs]
        res["hits"] = eng.search(recs, op)
        with open(op, newline="") as f:
            res["output"] = f.read()
        return res


def main():
    cases = []
    for seed in range(N_FUZZ):
        c = fuzzcases.make_case(seed)
        c["expect"] = run_reference(c["params"], c["sts_text"], c["fasta_text"])
        cases.append(c)
    with gzip.GzipFile(os.path.join(HERE, "fuzz_golden.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(cases, separators=(",", ":")).encode())
    nh = sum(c["expect"]["hits"] for c in cases)
    print(f"fuzz: {len(cases)} cases, {nh} hits, {sum(1 for c in cases if c['expect']['hits'])} with hits, "
          f"{sum(1 for c in cases if not c['expect']['load_ok'])} load failures")

    # threaded cases (>= 100 kbp so the reference really forks workers, engine.py:381-419)
    tcases = []
    for i, (seed, length, n_sts, threads, M) in enumerate(
            [(501, 130000, 40, 2, 50), (502, 220000, 60, 4, 50), (503, 150001, 30, 3, 10), (504, 100000, 25, 2, 50),
             (505, 260000, 80, 1, 50)]):
        genome = synth.dna_chunked(seed, length)
        sts = synth.make_sts_set(seed + 1000, n_sts, 18, 25, 100, 600)
        synth.plant_amplicons(seed + 2000, [genome], sts, M, sub_mode="cfg3")
        sts_text = synth.sts_lines(sts).decode()
        fasta_text = ">big%d synthetic\n" % i + "\n".join(
            genome[j: j + 60].tobytes().decode() for j in range(0, length, 60)) + "\n"
        params = dict(wordsize=11, margin=M, mismatches=1, three_prime_match=1, iupac_mode=0, default_pcr_size=240)
        exp = run_reference(params, sts_text, fasta_text, threads=threads)
        tcases.append(dict(seed=seed, length=length, n_sts=n_sts, threads=threads, params=params,
                           sha=hashlib.sha256((sts_text + fasta_text).encode()).hexdigest(),
                           expect=dict(hits=exp["hits"], output=exp["output"])))
        print(f"threaded case {i}: T={threads} hits={exp['hits']}")
    with open(os.path.join(HERE, "threaded_golden.json"), "w") as f:
        json.dump(tcases, f, indent=1)


if __name__ == "__main__":
    main()
