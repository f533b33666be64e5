"""Host-side nibble packer alone: GB/s of ASCII consumed for 1, 4, 8, 16 ... threads on this box's cores."""
import ctypes as C, os, sys, time
sys.path[:0] = ['.', 'tests']
import numpy as np
import synth
from merpcr_b200 import _capi
from merpcr_b200.alphabet import genome_lut
lib = C.CDLL(_capi.LIB_PATH)
lib.mpcr_host_pack_nibbles.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
n = 1 << 29
a = synth.dna_chunked(5, n)
out = np.ones(n // 2, dtype=np.uint8)
lut = genome_lut(0)
ncpu = len(os.sched_getaffinity(0))
print("cpus", ncpu, flush=True)
for th in sorted({1, 2, 4, 8, 16, 32, ncpu}):
    if th > ncpu:
        continue
    best = 1e9
    for _ in range(4):
        t = time.perf_counter(); rc = lib.mpcr_host_pack_nibbles(a.ctypes.data, n, lut.ctypes.data, out.ctypes.data, th)
        best = min(best, time.perf_counter() - t)
    print(f"threads {th:3d}: {n / best / 1e9:6.1f} GB/s of ASCII in, rc {rc}", flush=True)
