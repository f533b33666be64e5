"""ctypes binding of libmerpcr_b200.so (include/merpcr_b200.h) -- the only door between the Python host
and the CUDA kernels.  There is no CPU fallback: if the library or a CUDA device is missing every entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# $MPCR_B200_LIB selects a tuning build of the same CUDA library (merpcr_b200/build.py --variant=...)
LIB_PATH = os.environ.get("MPCR_B200_LIB") or os.path.join(_HERE, "lib", "libmerpcr_b200.so")

MPCR_OK, MPCR_EINVAL, MPCR_ECUDA, MPCR_ENOMEM, MPCR_ESTATE, MPCR_EOVERFLOW = 0, -1, -2, -3, -4, -5
ABI_VERSION = 15
MAX_SLOTS = 8

# every symbol include/merpcr_b200.h declares (tests check the built library exports all of them)
SYMBOLS = [
    "mpcr_abi_version", "mpcr_last_error", "mpcr_ctx_create", "mpcr_ctx_destroy", "mpcr_ctx_set_seed_extension", "mpcr_ctx_set_seed_blocks", "mpcr_ctx_set_table_part", "mpcr_ctx_set_sampling", "mpcr_table_items", "mpcr_ctx_set_append", "mpcr_scan_prepare", "mpcr_ctx_set_true_strands", "mpcr_ctx_sm_count",
    "mpcr_pack_sequence", "mpcr_host_pack_nibbles", "mpcr_derive_planes", "mpcr_file_read", "mpcr_fasta_workspace_bytes", "mpcr_fasta_index", "mpcr_fasta_index_ex", "mpcr_fasta_offsets_at", "mpcr_fasta_compact",
    "mpcr_sts_parse", "mpcr_sts_blob", "mpcr_format_hits", "mpcr_table_build", "mpcr_table_records", "mpcr_table_primer_words", "mpcr_scan",
    "mpcr_halo_left", "mpcr_halo_right", "mpcr_tile_bases", "mpcr_sort_hits", "mpcr_sort_hits_dev", "mpcr_scan_sorted", "mpcr_scan_sorted_async", "mpcr_sort_finish", "mpcr_slot_scan_ms", "mpcr_slot_verify_ms", "mpcr_launch_count", "mpcr_last_scan_ms", "mpcr_last_verify_ms",
]

HIT_DTYPE = np.dtype([("contig", "<u4"), ("pos1", "<u4"), ("pos2", "<u4"), ("rec", "<u4"), ("rank", "<u4"),
                      ("hash_off", "<u4")])
CONTIG_DTYPE = np.dtype([("gstart", "<u8"), ("length", "<u4"), ("reserved", "<u4")])
STS_LINE_DTYPE = np.dtype([(k, "<u4") for k in ("line_no", "id_off", "id_len", "p1_off", "p1_len", "p2_off", "p2_len",
                                                "size_off", "size_len", "alias_off", "alias_len")] + [("pcr_size", "<i4")])
FASTA_RECORD_DTYPE = np.dtype([("header_begin", "<u8"), ("header_end", "<u8"), ("seq_offset", "<u8"),
                               ("seq_length", "<u8")])


class Params(C.Structure):
    _fields_ = [("wordsize", C.c_int32), ("margin", C.c_int32), ("mismatches", C.c_int32),
                ("three_prime_match", C.c_int32), ("iupac_mode", C.c_int32)]


class Backend:
    """A loaded library + the torch device kind its 'device pointers' live on."""

    def __init__(self, lib: C.CDLL, device_kind: str):
        self.lib = lib
        self.device_kind = device_kind  # "cuda" for the product; "cpu" only for the injected test emulation
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
        lib.mpcr_abi_version.restype = i32
        lib.mpcr_last_error.restype = C.c_char_p
        lib.mpcr_ctx_create.restype = i32
        lib.mpcr_ctx_create.argtypes = [i32, C.POINTER(Params), C.POINTER(vp)]
        lib.mpcr_ctx_destroy.argtypes = [vp]
        lib.mpcr_ctx_set_seed_extension.restype = i32
        lib.mpcr_ctx_set_seed_extension.argtypes = [vp, i32, i32]
        lib.mpcr_ctx_set_seed_blocks.restype = i32
        lib.mpcr_ctx_set_seed_blocks.argtypes = [vp, i32, i32, i32]
        lib.mpcr_ctx_set_sampling.restype = i32
        lib.mpcr_ctx_set_sampling.argtypes = [vp, i32, i32, i32]
        lib.mpcr_table_items.restype = u32
        lib.mpcr_table_items.argtypes = [vp]
        lib.mpcr_ctx_set_append.restype = i32
        lib.mpcr_ctx_set_append.argtypes = [vp, i32]
        lib.mpcr_ctx_set_table_part.restype = i32
        lib.mpcr_ctx_set_table_part.argtypes = [vp, C.c_uint32, C.c_uint32]
        lib.mpcr_ctx_set_true_strands.restype = i32
        lib.mpcr_ctx_set_true_strands.argtypes = [vp, i32]
        lib.mpcr_ctx_sm_count.restype = i32
        lib.mpcr_ctx_sm_count.argtypes = [vp]
        lib.mpcr_pack_sequence.restype = i32
        lib.mpcr_pack_sequence.argtypes = [vp, vp, u64, u64, u64, vp, vp, vp, vp, vp]
        lib.mpcr_host_pack_nibbles.restype = i32
        lib.mpcr_host_pack_nibbles.argtypes = [vp, u64, vp, vp, i32]
        lib.mpcr_file_read.restype = C.c_longlong
        lib.mpcr_file_read.argtypes = [C.c_char_p, u64, u64, vp, i32]
        lib.mpcr_derive_planes.restype = i32
        lib.mpcr_derive_planes.argtypes = [vp, u64, u64, u64, vp, vp, vp, vp]
        lib.mpcr_fasta_workspace_bytes.restype = u64
        lib.mpcr_fasta_workspace_bytes.argtypes = [u64, u32]
        lib.mpcr_fasta_index.restype = i32
        lib.mpcr_fasta_index.argtypes = [vp, vp, u64, vp, u32, C.POINTER(u32), C.POINTER(u32), vp, u64, vp]
        lib.mpcr_fasta_index_ex.restype = i32
        lib.mpcr_fasta_index_ex.argtypes = [vp, vp, u64, u32, vp, u32, C.POINTER(u32), C.POINTER(u32), vp, u64, vp]
        lib.mpcr_fasta_offsets_at.restype = i32
        lib.mpcr_fasta_offsets_at.argtypes = [vp, vp, u64, vp, vp, u32, vp, vp]
        lib.mpcr_fasta_compact.restype = i32
        lib.mpcr_fasta_compact.argtypes = [vp, vp, u64, vp, vp, vp]
        lib.mpcr_sts_parse.restype = i32
        lib.mpcr_sts_parse.argtypes = [vp, u64, C.c_int32, C.c_int32, vp, u32, C.POINTER(u32), C.POINTER(u32),
                                       C.POINTER(u32), C.POINTER(u32)]
        lib.mpcr_sts_blob.restype = i32
        lib.mpcr_sts_blob.argtypes = [vp, vp, u32, vp, vp]
        lib.mpcr_format_hits.restype = u64
        lib.mpcr_format_hits.argtypes = [vp, u64, vp, vp, vp, vp, vp, u64]
        lib.mpcr_table_build.restype = i32
        lib.mpcr_table_build.argtypes = [vp, vp, vp, vp, u32, vp, vp]
        lib.mpcr_table_records.restype = i32
        lib.mpcr_table_records.argtypes = [vp, vp, vp]
        lib.mpcr_table_primer_words.restype = i32
        lib.mpcr_table_primer_words.argtypes = [vp, u32, i32, vp, u32, C.POINTER(u32)]
        lib.mpcr_scan_prepare.restype = i32
        lib.mpcr_scan_prepare.argtypes = [vp, vp, u32, u64, u64, u64, vp]
        lib.mpcr_scan.restype = i32
        lib.mpcr_scan.argtypes = [vp, vp, u32, vp, vp, vp, u64, u64, u64, u64, vp, u64, vp, vp]
        lib.mpcr_halo_left.restype = u64
        lib.mpcr_halo_left.argtypes = [vp]
        lib.mpcr_halo_right.restype = u64
        lib.mpcr_halo_right.argtypes = [vp]
        lib.mpcr_tile_bases.restype = u64
        lib.mpcr_tile_bases.argtypes = []
        lib.mpcr_sort_hits.restype = i32
        lib.mpcr_sort_hits.argtypes = [vp, vp, u64, vp]
        lib.mpcr_sort_hits_dev.restype = i32
        lib.mpcr_sort_hits_dev.argtypes = [vp, vp, vp, u64, u64, vp]
        lib.mpcr_scan_sorted.restype = i32
        lib.mpcr_scan_sorted.argtypes = [vp, u32, vp, u32, vp, vp, vp, u64, u64, u64, u64, vp, u64, vp, vp, u64, i32, vp]
        lib.mpcr_scan_sorted_async.restype = i32
        lib.mpcr_scan_sorted_async.argtypes = [vp, u32, vp, u32, vp, vp, vp, u64, u64, u64, u64, vp, u64, vp, vp, u64, i32, i32,
                                               vp]
        lib.mpcr_sort_finish.restype = i32
        lib.mpcr_sort_finish.argtypes = [vp, vp, vp, u64, vp]
        lib.mpcr_slot_scan_ms.restype = C.c_float
        lib.mpcr_slot_scan_ms.argtypes = [vp, i32]
        lib.mpcr_slot_verify_ms.restype = C.c_float
        lib.mpcr_slot_verify_ms.argtypes = [vp, i32]
        lib.mpcr_launch_count.restype = u64
        lib.mpcr_launch_count.argtypes = [vp]
        lib.mpcr_last_scan_ms.restype = C.c_float
        lib.mpcr_last_scan_ms.argtypes = [vp]
        lib.mpcr_last_verify_ms.restype = C.c_float
        lib.mpcr_last_verify_ms.argtypes = [vp]
        if lib.mpcr_abi_version() != ABI_VERSION:
            raise RuntimeError("libmerpcr_b200.so ABI version mismatch; rebuild with `python -m merpcr_b200.build`")

    def check(self, rc: int) -> None:
        if rc == MPCR_OK:
            return
        msg = (self.lib.mpcr_last_error() or b"").decode("utf-8", "replace")
        if rc == MPCR_EINVAL:
            raise ValueError(msg)
        if rc == MPCR_ENOMEM:
            raise MemoryError(msg)
        raise RuntimeError(f"merpcr_b200: {msg} (code {rc})")


_backend = None


def backend() -> Backend:
    """The CUDA library.  Fails loudly when it has not been built -- there is no other implementation."""
    global _backend
    if _backend is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m merpcr_b200.build` "
                "(needs nvcc). merpcr_b200 has no CPU fallback.")
        _backend = Backend(C.CDLL(LIB_PATH), "cuda")
    return _backend


def _inject_backend_for_tests(path: str, device_kind: str = "cpu") -> None:
    """TEST HOOK ONLY: swap in tests/host_emul's serial emulation of the C ABI so CPU-only CI can exercise the
    host logic.  Never called by product code; GPU parity tests never use it."""
    global _backend
    _backend = Backend(C.CDLL(path), device_kind) if path else None
