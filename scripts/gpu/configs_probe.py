#!/usr/bin/env python3
"""BASELINE.json configs 2..5 at (or near) full size on one B200: timings plus the size-independent checks of
tests/fullsize.py.   usage: configs_probe.py [cfg2 cfg3 cfg4 cfg5] [--scale=S] [--timing-only]
Writes one JSON line per config to stdout."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import fullsize  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 1.0
    for a in sys.argv[1:]:
        if a.startswith("--scale="):
            scale = float(a.split("=", 1)[1])
    which = args or ["cfg2", "cfg4", "cfg5"]
    dev = torch.device("cuda", 0)
    cfgs = fullsize.configs(scale)
    for k in which:
        fullsize.run(*cfgs[k], dev, oracle=("slices", 16, 2_000_000) if k.startswith("cfg5") else "whole")


if __name__ == "__main__":
    main()
