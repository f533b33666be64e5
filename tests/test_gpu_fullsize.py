"""GPU tier, BASELINE.json sizes: configs 2, 3 and 4 at full size with the COMPLETE ordered hit list of the whole genome
compared bit for bit with the oracle (one contig per host thread, each single-threaded == reference -T 1), config 5 (a
quarter of its 10^6-line STS table, which takes a minute of host-side generation at full size) on 16 slices of 2 Mbp
spread over the contigs including both genome ends, and the same with one mismatch allowed (cfg5n1: block tables);
plus the size-independent properties of the domain -- every planted amplicon found (truth known by construction),
output in the reference's order, rescan idempotent -- and bp-balanced shards merging to the whole."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _real_backend():
    import torch
    from merpcr_b200 import _capi
    _capi._inject_backend_for_tests("", "cpu")
    assert _capi.backend().device_kind == "cuda" and torch.cuda.is_available()
    yield


@pytest.mark.parametrize("name,scale", [("cfg2", 1.0), ("cfg3", 1.0), ("cfg4", 1.0), ("cfg5", 0.25),
                                        ("cfg5n1", 0.25)])
def test_baseline_config_properties(name, scale):
    import torch
    import fullsize
    out = fullsize.run(*fullsize.configs(scale)[name], torch.device("cuda", 0),
                       oracle=("slices", 16, 2_000_000) if name.startswith("cfg5") else "whole", verbose=False)
    assert out["planted"] > 1000 and out["planted_found"], out
    assert out["sorted"] and out["idempotent"], out
    assert out["oracle_bit_exact"] and out["oracle_hits"] > 1000, out
    if not name.startswith("cfg5"):
        assert out["oracle_bp"] == out["bp"] and out["oracle_hits"] == out["hits"], out


def test_two_shards_merge_to_the_whole_at_full_size():
    """cfg3-sized genome: the hits of two bp-balanced shards (cut inside a chromosome, halos) == the unsharded hits."""
    import tempfile, os
    import torch
    import fullsize, synth
    from merpcr_b200 import MerPCR, _capi
    dev = torch.device("cuda", 0)
    _, lengths, n_sts, params, sub_mode, ranged, seed, decorate = fullsize.configs(1.0)["cfg3"]
    sts = synth.make_sts_set(seed + 1, n_sts, 18, 25, 100, 1000)
    contigs = fullsize.build_genome(seed, lengths, dev)
    fullsize.plant(contigs, lengths, sts, params["margin"], seed + 2, sub_mode)
    with tempfile.NamedTemporaryFile("wb", suffix=".sts", delete=False) as f:
        f.write(synth.sts_lines(sts))
    parts = []
    try:
        for shard in (None, (0, 2), (1, 2)):
            eng = MerPCR(**params, device=0, shard=shard)
            assert eng.load_sts_file(f.name)
            layout = eng.make_layout(lengths)
            sh = eng.upload(layout, contigs)
            parts.append(eng.scan(layout, sh))
            eng.close()
            del sh
            torch.cuda.empty_cache()
    finally:
        os.unlink(f.name)
    whole, a, b = parts
    merged = np.concatenate([a, b])
    order = np.lexsort((merged["rank"], merged["rec"], merged["hash_off"], merged["pos1"], merged["contig"]))
    assert len(whole) > 50000 and np.array_equal(merged[order], whole)
    assert len(a) > 10000 and len(b) > 10000
