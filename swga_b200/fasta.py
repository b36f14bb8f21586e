"""FASTA input for the hot path (host side).

`read_fasta` hands the file image to the library's native reader (swga_fasta_scan), which follows
DSK's record rules (ext/dsk/minia/Bank.cpp:150-209) so that counts match `dsk` byte for byte, and
returns the joined bases, the record boundaries and the record names.  Names follow pyfaidx's rule
-- first whitespace-delimited token of the header -- which swga/locate.py:42 and
swga/tests/test_locate.py:12-22 rely on ('> record1' -> 'record1').
"""
import ctypes
import os

import numpy as np

from . import _lib


class Genome(object):
    """Joined bases of all records (no separators) + record boundaries + record names."""

    def __init__(self, seq, rec_starts, names, path=None):
        self.seq = seq                      # np.uint8[N] (ASCII)
        self.rec_starts = rec_starts        # np.uint64[R+1], last = N
        self.names = names                  # list[str], pyfaidx keys, file order
        self.path = path

    @property
    def n_bases(self):
        return int(self.rec_starts[-1]) if len(self.rec_starts) else 0

    @property
    def n_records(self):
        return max(0, len(self.rec_starts) - 1)

    @property
    def dsk_mode_b(self):
        """True when DSK would read this file raw ('N' -> G): some early record > 1 Mbp."""
        rs = np.ascontiguousarray(self.rec_starts, dtype=np.uint64)
        return bool(_lib.load().swga_dsk_mode(_lib.ptr(rs), self.n_records))

    def record_lengths(self):
        return np.diff(self.rec_starts.astype(np.int64))


def _alloc_u8(n, pin):
    if pin:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True)
            return t.numpy()[:n]            # the numpy view keeps the pinned tensor alive
    return np.empty(n, dtype=np.uint8)


def parse_fasta_bytes(data, path=None, pin=False):
    lib = _lib.load()
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), dtype=np.uint8)
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    n_bases, n_rec = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _lib.check(lib.swga_fasta_scan(_lib.ptr(buf), len(buf), None, None, None,
                                   ctypes.byref(n_bases), ctypes.byref(n_rec)))
    seq = _alloc_u8(n_bases.value, pin)
    rs = np.zeros(n_rec.value + 1, dtype=np.uint64)
    hdr = np.zeros(2 * n_rec.value + 2, dtype=np.uint64)
    _lib.check(lib.swga_fasta_scan(_lib.ptr(buf), len(buf), _lib.ptr(seq) if n_bases.value else None,
                                   _lib.ptr(rs), _lib.ptr(hdr), ctypes.byref(n_bases), ctypes.byref(n_rec)))
    names = []
    raw = buf
    for r in range(n_rec.value):
        o, l = int(hdr[2 * r]), int(hdr[2 * r + 1])
        tok = raw[o:o + l].tobytes().decode("latin-1").split()
        names.append(tok[0] if tok else "")
    return Genome(seq, rs, names, path)


def read_image(path):
    """The bytes of a FASTA file, plain or gzip-compressed (DSK opens its input through zlib's gzopen, which
    reads both: ext/dsk/minia/Bank.cpp:39,275), as a uint8 array."""
    with open(path, "rb") as f:
        magic = f.read(2)
    if magic == b"\x1f\x8b":
        import gzip
        with gzip.open(path, "rb") as f:
            return np.frombuffer(f.read(), dtype=np.uint8)
    if os.path.getsize(path) > 0:
        return np.memmap(path, dtype=np.uint8, mode="r")          # the reader's threads read the page cache directly
    return np.zeros(0, dtype=np.uint8)


def read_fasta(path, pin=False):
    """Read a FASTA file into a Genome (joined bases on the host)."""
    return parse_fasta_bytes(read_image(path), path=os.path.abspath(path), pin=pin)
