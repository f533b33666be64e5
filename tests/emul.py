"""Builds and injects tests/host_emul (serial CPU emulation of the C ABI) -- CPU test tier only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "host_emul.cpp")
OUT = os.path.join(HERE, "host_emul", "_build", "libmerpcr_emul.so")
DEPS = [SRC, os.path.join(HERE, "..", "merpcr_b200", "csrc", "mpcr_core.cuh"),
        os.path.join(HERE, "..", "merpcr_b200", "csrc", "mpcr_hostio.h"),
        os.path.join(HERE, "..", "merpcr_b200", "csrc", "mpcr_hostpack.cpp"),
        os.path.join(HERE, "..", "include", "merpcr_b200.h")]


def build() -> str:
    if not os.path.exists(OUT) or any(os.path.getmtime(OUT) < os.path.getmtime(d) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-Wno-psabi", "-shared", "-o", OUT, SRC, "-lpthread"])
    return OUT


def inject():
    from merpcr_b200 import _capi
    _capi._inject_backend_for_tests(build(), "cpu")


def restore():
    from merpcr_b200 import _capi
    _capi._inject_backend_for_tests("", "cpu")
