"""All-k k-mer counting on the B200 -- the drop-in for the DSK step behind swga/kmers.py.

Reference seams reproduced here:
  * `count_kmers(k, genome_fp, cwd, threshold=1)`   swga/kmers.py:23-56  (same name, arguments,
    cache-file side effect and error behaviour; the `dsk` subprocess is replaced by the CUDA path)
  * `_parse_kmer_binary(fp)`                        swga/kmers.py:94-115
  * `parse_kmer_file(lines)`, `max_sequential_nt`   swga/kmers.py:14-20,59-91 (host helpers)

`count_all` is the native form: one pass over the genome yields the dense canonical tables for every
k in [kmin, kmax] (include/swga_b200.h: swga_count_genome_host / swga_count_genome_device).
"""
import os
import re
import struct

import numpy as np

from . import _lib, fasta

DEFAULT_KMIN, DEFAULT_KMAX = 5, 12          # swga/commands/specfiles/count.yaml:5-12

_ALPHABET = np.frombuffer(b"ACTG", dtype=np.uint8)     # ext/dsk/minia/Kmer.cpp:152 bin2NT


def message(msg):                                       # swga/__init__.py:22-30 (stderr)
    import sys
    sys.stderr.write(msg + "\n")


# --------------------------------------------------------------------------- code <-> string

def codes_to_kmers(codes, k):
    """Vectorised Kmer.cpp:148-162 code2seq: uint codes -> list[str] (first base most significant)."""
    codes = np.asarray(codes, dtype=np.uint64)
    out = np.empty((len(codes), k), dtype=np.uint8)
    for j in range(k):
        out[:, j] = _ALPHABET[((codes >> np.uint64(2 * (k - 1 - j))) & np.uint64(3)).astype(np.intp)]
    return [s.decode("ascii") for s in out.view("S%d" % k).ravel()] if k > 0 else []


def kmer_to_code(s):
    """Kmer.cpp:65-76 codeSeed with NT2int = (c>>1)&3."""
    x = 0
    for ch in s:
        x = x * 4 + ((ord(ch) >> 1) & 3)
    return x


def revcomp_code(code, k):
    """Kmer.cpp:172-186: reverse the base-4 digits, complement = digit ^ 2 (lut.h:8-10)."""
    r = 0
    for _ in range(k):
        r = (r << 2) | ((code & 3) ^ 2)
        code >>= 2
    return r


# --------------------------------------------------------------------------- dense tables

class CountTables(object):
    """Dense canonical count tables for k = kmin..kmax (uint32, indexed by DSK's canonical code;
    non-canonical slots are 0).  `block` is the concatenation the library returns."""

    def __init__(self, block, kmin, kmax):
        self.block, self.kmin, self.kmax = block, kmin, kmax
        lib = _lib.load()
        self._off = {k: int(lib.swga_table_offset(kmin, k)) for k in range(kmin, kmax + 2)}

    def table(self, k):
        if not self.kmin <= k <= self.kmax:
            raise KeyError(k)
        return self.block[self._off[k]:self._off[k + 1]]

    def as_dict(self, k, threshold=1):
        """{canonical k-mer: count} for counts >= threshold -- what swga.kmers.count_kmers returns."""
        t = self.table(k)
        if hasattr(t, "cpu"):
            t = t.cpu().numpy()
        nz = np.nonzero(t >= max(1, int(threshold)))[0]
        return dict(zip(codes_to_kmers(nz, k), t[nz].tolist()))


def count_all(genome, kmin=DEFAULT_KMIN, kmax=DEFAULT_KMAX, device=None):
    """Count every canonical k-mer, k in [kmin, kmax], of a FASTA file (path) or `fasta.Genome`
    through the host-buffer C-ABI calls.  Returns CountTables on the host (numpy).  A path goes through
    swga_count_fasta_host: the file image is parsed straight into the library's pinned staging blocks, no host
    copy of the joined bases is made (`n_bases`, `n_records`, `mode_b` of the file are set on the result)."""
    import ctypes
    lib = _lib.load()
    ctx = _lib.context(device)
    words = int(lib.swga_tables_words(kmin, kmax))
    out = np.empty(words, dtype=np.uint32)
    if not isinstance(genome, fasta.Genome):
        image = fasta.read_image(genome)
        nb, nr, mode = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_int(0)
        _lib.check(lib.swga_count_fasta_host(ctx.h, _lib.ptr(image) if len(image) else None, len(image), -1, kmin, kmax,
                                             _lib.ptr(out), ctypes.byref(nb), ctypes.byref(nr), ctypes.byref(mode)))
        tables = CountTables(out, kmin, kmax)
        tables.n_bases, tables.n_records, tables.mode_b = nb.value, nr.value, bool(mode.value)
        return tables
    rs = np.ascontiguousarray(genome.rec_starts, dtype=np.uint64)
    _lib.check(lib.swga_count_genome_host(ctx.h, _lib.ptr(genome.seq) if genome.n_bases else None,
                                          genome.n_bases, _lib.ptr(rs), genome.n_records,
                                          int(genome.dsk_mode_b), kmin, kmax, _lib.ptr(out)))
    return CountTables(out, kmin, kmax)


def count_all_device(d_ascii, rec_starts, mode_b, kmin=DEFAULT_KMIN, kmax=DEFAULT_KMAX, n_starts=None,
                     finalize=True, d_fwd=None, d_canon=None, d_rec_starts=None):
    """Device-resident form: `d_ascii` is a CUDA uint8 tensor of the joined bases (16-byte aligned).
    Returns (d_fwd, d_canon) CUDA uint32 tensors; with finalize=False d_fwd holds the additive
    "top-k + tails" form that can be allreduced across ranks (see swga_b200/dist.py)."""
    import torch
    lib = _lib.load()
    ctx = _lib.context(d_ascii.device.index)
    n = d_ascii.numel()
    words = int(lib.swga_tables_words(kmin, kmax))
    if d_fwd is None:
        d_fwd = torch.empty(words, dtype=torch.int32, device=d_ascii.device)
    if d_canon is None and finalize:
        d_canon = torch.empty(words, dtype=torch.int32, device=d_ascii.device)
    n_rec = len(rec_starts) - 1
    if d_rec_starts is None and n_rec > 1:
        d_rec_starts = torch.from_numpy(np.ascontiguousarray(rec_starts, dtype=np.uint64).view(np.int64)).to(
            d_ascii.device)
    stream = torch.cuda.current_stream(d_ascii.device).cuda_stream
    _lib.check(lib.swga_count_genome_device(
        ctx.h, _lib.ptr(d_ascii), n, _lib.ptr(d_rec_starts) if n_rec > 1 else None, n_rec,
        0 if mode_b else 1, kmin, kmax, n if n_starts is None else n_starts, int(bool(finalize)),
        _lib.ptr(d_fwd), _lib.ptr(d_canon) if finalize else None, stream))
    return d_fwd, d_canon


# --------------------------------------------------------------------------- reference seam

_TABLE_CACHE = {}      # (abspath, mtime, size) -> CountTables for DEFAULT range; one genome pass serves all k


def _tables_for(genome_fp, k):
    st = os.stat(genome_fp)
    key = (os.path.abspath(genome_fp), st.st_mtime_ns, st.st_size)
    kmin, kmax = min(k, DEFAULT_KMIN), max(k, DEFAULT_KMAX)
    hit = _TABLE_CACHE.get(key)
    if hit is None or not (hit.kmin <= k <= hit.kmax):
        hit = count_all(genome_fp, kmin, kmax)
        _TABLE_CACHE.clear()
        _TABLE_CACHE[key] = hit
    return hit


def count_kmers(k, genome_fp, cwd, threshold=1):
    """Counts k-mers in the specified genome (same contract as swga/kmers.py:23-56).

    Returns {canonical k-mer: both-strand count} for counts >= threshold.  Side effect kept from the
    reference: `<cwd>/<basename(genome_fp)>-<k>mers.solid_kmers_binary` is written in DSK's result
    format (SortingCount.cpp:149-156,649-653) and re-used when present (keyed by name and k only,
    kmers.py:34-40); on failure the partial file is removed and the exception re-raised.
    """
    assert isinstance(threshold, int)
    genome = genome_fp.split(os.sep).pop()
    out = "%s-%dmers" % (genome, k)
    outfile = os.path.join(cwd, out + ".solid_kmers_binary")
    if os.path.isfile(outfile):
        message("Binary kmer file already exists at %s; parsing..." % outfile)
    else:
        message("In {cwd}:\n> libswga_b200 count {fp} k={k} -o {out} -t {t}".format(
            cwd=cwd, fp=genome_fp, k=k, out=out, t=threshold))
        try:
            tables = _tables_for(genome_fp, k)
            t = np.ascontiguousarray(tables.table(k))
            _lib.check(_lib.load().swga_write_dsk_file(outfile.encode(), k, _lib.ptr(t), int(threshold), None))
        except BaseException:
            if os.path.isfile(outfile):
                os.remove(outfile)
            raise
    return dict((kmer, freq) for kmer, freq in _parse_kmer_binary(outfile) if freq >= threshold)


def _parse_kmer_binary(fp):
    """swga/kmers.py:94-115: [i32 nbits][i32 k] then ([nbits/8 B LE code][u32 count])*; a truncated
    header removes the file and raises, a short trailing record is ignored."""
    with open(fp, "rb") as f:
        head = f.read(8)
        if len(head) < 8:
            if os.path.isfile(fp):
                os.remove(fp)
            raise struct.error("unpack requires a buffer of 4 bytes")
        kmer_nbits, k = struct.unpack("ii", head)
        nb = kmer_nbits // 8
        body = np.frombuffer(f.read(), dtype=np.uint8)
    rec = nb + 4
    n = len(body) // rec
    body = body[: n * rec].reshape(n, rec)
    codes = np.zeros(n, dtype=np.uint64)
    for i in range(min(nb, 8)):
        codes |= body[:, i].astype(np.uint64) << np.uint64(8 * i)
    freqs = np.ascontiguousarray(body[:, nb:nb + 4]).view("<u4").ravel()
    return zip(codes_to_kmers(codes, k), freqs.tolist())


def parse_kmer_file(lines):
    """swga/kmers.py:14-20: first whitespace-delimited field of each line."""
    return [re.split(r"[ \t]+", line.strip("\n"))[0] for line in lines]


def max_sequential_nt(mer1, mer2):
    """swga/kmers.py:59-91: longest run of consecutive complementary bases between two k-mers, over
    all offsets of mer1 against reversed mer2 (host helper; the homodimer filter of `swga count`)."""
    binding = {"A": "T", "T": "A", "C": "G", "G": "C", "_": False}
    if len(mer2) > len(mer1):
        mer1, mer2 = mer2, mer1
    n1 = len(mer1)
    mer2 = mer2[::-1]
    mer1 = mer1.ljust(n1 + len(mer2), "_")
    best = 0
    for offset in range(n1):
        run = 0
        for x in range(len(mer2)):
            if binding[mer1[offset + x]] == mer2[x]:
                run += 1
                best = max(best, run)
            else:
                run = 0
    return best
