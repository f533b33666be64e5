"""GPU tier, needs >= 2 GPUs (skipped otherwise): `python -m merpcr_b200 --gpus N` -- the launcher spawns one process
per GPU, every rank scans its shard (NCCL only for the final hit gather), rank 0 writes the merged list; the text must
equal the single-GPU run byte for byte."""
import os
import subprocess
import sys

import pytest

import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("extra", [[], ["-I", "1", "-N", "2"]])
def test_cli_gpus_launcher_equals_single_gpu(tmp_path, extra):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    rng = synth.Rng(97)
    contigs = [rng.dna(m) for m in (3_000_000, 700, 1_200_000, 2_500_001)]
    sts = synth.make_sts_set(98, 2000, 18, 25, 100, 900)
    synth.plant_amplicons(99, contigs, sts, 50, sub_mode="cfg3")
    sts_path, fa_path = str(tmp_path / "m.sts"), str(tmp_path / "m.fa")
    with open(sts_path, "wb") as f:
        f.write(synth.sts_lines(sts))
    with open(fa_path, "wb") as f:
        for i, c in enumerate(contigs):
            f.write(b">ctg%d\n" % i)
            body = c[: len(c) // 60 * 60].reshape(-1, 60)
            f.write(b"\n".join(r.tobytes() for r in body) + b"\n" + c[len(c) // 60 * 60:].tobytes() + b"\n")
    env = dict(os.environ, PYTHONPATH=ROOT)
    outs = {}
    for g in (1, min(n, 4)):
        out = str(tmp_path / f"out{g}.txt")
        rc = subprocess.run([sys.executable, "-m", "merpcr_b200", "--gpus", str(g), "-N", "1", "-O", out] + extra +
                            [sts_path, fa_path], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
        assert rc.returncode == 0, rc.stderr[-2000:]
        outs[g] = open(out).read()
    a, b = outs.values()
    assert a == b and a.count("\n") > 1000
