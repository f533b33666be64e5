"""The built CUDA library loads on a CPU-only box and exports every symbol include/merpcr_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re


def test_library_exports_declared_symbols():
    from merpcr_b200 import _capi, build
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(os.path.dirname(path), "..", "..", "include", "merpcr_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mpcr_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.mpcr_abi_version.restype = ctypes.c_int
    assert lib.mpcr_abi_version() == _capi.ABI_VERSION


def test_product_fails_loudly_without_a_device():
    import pytest
    import torch
    from merpcr_b200 import MerPCR, _capi
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    _capi._inject_backend_for_tests("", "cpu")     # make sure the real library is the active backend
    with pytest.raises(RuntimeError):
        MerPCR()
