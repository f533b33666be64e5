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


def _real_lib():
    from merpcr_b200 import _capi, build
    return _capi.Backend(ctypes.CDLL(build.build()), "cuda").lib


def test_host_text_entry_points_of_the_real_library():
    """mpcr_sts_parse / mpcr_sts_blob / mpcr_format_hits / size queries are host C++ inside the CUDA library: they
    run (and are checked) here without a GPU, straight through the C ABI."""
    import numpy as np
    from merpcr_b200 import _capi
    lib = _real_lib()
    text = (b"# comment\r\nSTS1\tacgtacgtacgtA\tTTTTGGGGCCCCAAAA\t100-200\talias one\n"
            b"short\tACGT\tACGTACGTACGTACGT\t50\n\n"
            b"  STS2\tACGTNACGTACGTACGT\tGGGGCCCCAAAATTTT\t+7\r"
            b"STS3\tACGTACGTACGTACGT\tGGGGCCCCAAAATTTT\t0\talias\textra\n")
    raw = np.frombuffer(text, dtype=np.uint8)
    lines = np.zeros(8, dtype=_capi.STS_LINE_DTYPE)
    n, bad, short, flags = (ctypes.c_uint32(0) for _ in range(4))
    rc = lib.mpcr_sts_parse(raw.ctypes.data, raw.size, 11, 240, lines.ctypes.data, 8, ctypes.byref(n), ctypes.byref(bad),
                            ctypes.byref(short), ctypes.byref(flags))
    assert (rc, n.value, bad.value, short.value, flags.value) == (0, 3, 0, 1, 0)
    assert lines["line_no"][:3].tolist() == [2, 5, 6]
    assert lines["pcr_size"][:3].tolist() == [150, -1, 240]          # "+7" is left to Python's int(); "0" -> default
    field = lambda i, k: text[int(lines[k + "_off"][i]): int(lines[k + "_off"][i] + lines[k + "_len"][i])]
    assert (field(0, "id"), field(0, "alias"), field(1, "id"), field(1, "alias"), field(2, "alias")) == \
        (b"STS1", b"alias one", b"STS2", b"", b"alias")
    blob = np.zeros(256, dtype=np.uint8)
    off = np.zeros(7, dtype=np.uint64)
    assert lib.mpcr_sts_blob(raw.ctypes.data, lines.ctypes.data, 3, blob.ctypes.data, off.ctypes.data) == 0
    assert blob[: int(off[1])].tobytes() == b"ACGTACGTACGTA" and int(off[6]) == 13 + 16 + 17 + 16 + 16 + 16
    # overflow protocol and malformed line
    rc = lib.mpcr_sts_parse(raw.ctypes.data, raw.size, 11, 240, lines.ctypes.data, 1, ctypes.byref(n), ctypes.byref(bad),
                            ctypes.byref(short), ctypes.byref(flags))
    assert rc == _capi.MPCR_EOVERFLOW and n.value == 3
    bad_text = np.frombuffer(b"a\tb\tc\n", dtype=np.uint8)
    lib.mpcr_sts_parse(bad_text.ctypes.data, bad_text.size, 3, 240, lines.ctypes.data, 8, ctypes.byref(n),
                       ctypes.byref(bad), ctypes.byref(short), ctypes.byref(flags))
    assert bad.value == 1
    # engine.py:442 formatting
    hits = np.zeros(2, dtype=_capi.HIT_DTYPE)
    hits[0] = (0, 9, 208, 0, 0, 0)
    hits[1] = (1, 4294967294, 5, 5, 2, 0)
    labels = np.frombuffer(b"chr1L78833\0", dtype=np.uint8)
    loff = np.array([0, 4, 10], dtype=np.uint64)
    args = (hits.ctypes.data, 2, raw.ctypes.data, lines.ctypes.data, labels.ctypes.data, loff.ctypes.data)
    need = lib.mpcr_format_hits(*args, None, 0)
    out = np.zeros(need, dtype=np.uint8)
    used = lib.mpcr_format_hits(*args, out.ctypes.data, need)
    assert out[:used].tobytes() == b"chr1\t10..209\tSTS1\talias one\t(+)\nL78833\t4294967295..6\tSTS3\talias\t(-)\n"
    assert lib.mpcr_tile_bases() % 2048 == 0 and lib.mpcr_fasta_workspace_bytes(1 << 20, 16) > (1 << 20) // 4096 * 12
