"""Scratch: host FASTA reader (swga_fasta_scan) throughput, no GPU needed."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swga_b200 import _lib, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
path = "/dev/shm/swga_scan.fa"
if not os.path.isfile(path) or os.path.getsize(path) < n:
    synth.write_fasta(path, synth.bases(4, 0, n, 0.41), synth.equal_records(n, 4))
lib = _lib.load()
raw = np.memmap(path, dtype=np.uint8, mode="r")
nb, nr = ctypes.c_uint64(0), ctypes.c_uint64(0)
seq = np.empty(n, dtype=np.uint8); seq[:] = 0
rs = np.zeros(16, dtype=np.uint64); hdr = np.zeros(32, dtype=np.uint64)
for it in range(3):
    t0 = time.perf_counter()
    _lib.check(lib.swga_fasta_scan(_lib.ptr(raw), len(raw), None, None, None, ctypes.byref(nb), ctypes.byref(nr)))
    t1 = time.perf_counter()
    _lib.check(lib.swga_fasta_scan(_lib.ptr(raw), len(raw), _lib.ptr(seq), _lib.ptr(rs), _lib.ptr(hdr), ctypes.byref(nb), ctypes.byref(nr)))
    t2 = time.perf_counter()
    print("sizes %.1f ms (%.1f GB/s)  fill %.1f ms (%.1f GB/s)  bases %d" % ((t1 - t0) * 1e3, len(raw) / (t1 - t0) / 1e9, (t2 - t1) * 1e3, len(raw) / (t2 - t1) / 1e9, nb.value), flush=True)
import hashlib
print("md5", hashlib.md5(seq.tobytes()).hexdigest())
