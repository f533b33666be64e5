"""Build libmerpcr_b200.so in-tree with nvcc for sm_100a:  python -m merpcr_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repository snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "mpcr_kernels.cu")]
HOST_SRC = os.path.join(HERE, "csrc", "mpcr_hostpack.cpp")    # host-side packer: g++ (AVX2 / AVX-512 paths), linked in
DEPS = SRC + [HOST_SRC] + [os.path.join(HERE, "csrc", f) for f in ("mpcr_core.cuh", "mpcr_sort.cuh", "mpcr_fasta.cuh", "mpcr_hostio.h")] + \
    [os.path.join(os.path.dirname(HERE), "include", "merpcr_b200.h")]
OUT = os.path.join(HERE, "lib", "libmerpcr_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550"]


def find_nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmerpcr_b200.so")


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """variant/defines: tuning builds (lib/libmerpcr_b200_<variant>.so with -D overrides), selected at run time with
    $MPCR_B200_LIB; the product is the default build."""
    out = OUT if not variant else OUT.replace(".so", f"_{variant}.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    host_obj = out.replace(".so", "_hostpack.o")
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-Wno-psabi", "-c", HOST_SRC,
                           "-o", host_obj])
    cmd = [find_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", out] + SRC + [host_obj]
    subprocess.check_call(cmd)
    os.unlink(host_obj)
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-D") and not a.startswith("--variant=")]
    variant = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), "")
    print(build(force="--force" in args or bool(variant), verbose="-v" in args, variant=variant,
                defines=[a[2:] for a in sys.argv[1:] if a.startswith("-D")]))
