"""Binding-site location on the B200 -- the drop-in for swga/locate.py.

Reference seams (same names and argument meaning):
  * `binding_sites(kmer, genome_fp)`            swga/locate.py:39-51
  * `chromosome_ends(genome_fp)`                swga/locate.py:59-73
  * `linearize_binding_sites(primers, ends)`    swga/locate.py:22-36
  * `revcomp(s)`, `substr_indices(sub, s)`      swga/locate.py:12-19,76-90

Instead of two `str.find` scans of the whole genome per primer, the foreground genome is packed once
and a CSR position index (counting sort, csrc/index.cu) is built per primer length; a primer's sites
are then two O(1) bucket look-ups (its code and its reverse complement's).
"""
import os

import numpy as np

from . import _lib, fasta, kmers

_RC = {"A": "T", "T": "A", "G": "C", "C": "G"}


def revcomp(s):
    """Reverse complement of a DNA string (KeyError on anything but upper-case ACGT, as the reference)."""
    return "".join([_RC[i] for i in s][::-1])


def substr_indices(substring, string):
    """Overlapping occurrences of `substring` (upper-cased) in `string` (swga/locate.py:76-90).  Host
    helper kept for API compatibility; `binding_sites` does not use it."""
    locations = []
    start = 0
    substring = substring.upper()
    while True:
        start = string.find(substring, start) + 1
        if start > 0:
            locations.append(start - 1)
        else:
            return locations


class GenomeIndex(object):
    """A genome resident on the device in locate form: 2-bit packed bases, the strict break mask
    (only upper-case A/C/G/T inside one record can match, locate.py:83-86) and lazily built per-k CSR
    indexes."""

    def __init__(self, genome, device=None):
        import torch
        self.lib = _lib.load()
        self.ctx = _lib.context(device)
        self.genome = genome if isinstance(genome, fasta.Genome) else fasta.read_fasta(genome)
        g = self.genome
        self.dev = torch.device("cuda", self.ctx.device)
        self.n = g.n_bases
        if self.n >= (1 << 32) - 1:
            raise ValueError("locate index supports genomes below 2^32-1 bases")
        self.rec_starts = np.ascontiguousarray(g.rec_starts, dtype=np.uint64)
        P = _lib.ptr
        with torch.cuda.device(self.dev):
            d_ascii = torch.from_numpy(np.ascontiguousarray(g.seq)).to(self.dev) if self.n else torch.empty(
                16, dtype=torch.uint8, device=self.dev)
            self.packed = torch.empty(int(self.lib.swga_packed_words(self.n)), dtype=torch.int32, device=self.dev)
            self.brk = torch.empty(int(self.lib.swga_brk_words(self.n)), dtype=torch.int32, device=self.dev)
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(self.lib.swga_pack(self.ctx.h, P(d_ascii), self.n, 2, P(self.packed), P(self.brk), stream))
            if g.n_records > 1:
                d_rs = torch.from_numpy(self.rec_starts.view(np.int64).copy()).to(self.dev)
                _lib.check(self.lib.swga_mark_records(self.ctx.h, P(d_rs[1:]), g.n_records - 1, self.n, P(self.brk),
                                                      stream))
            torch.cuda.synchronize()
        self._idx = {}

    # ---- per-k index -------------------------------------------------------------------------
    def index(self, k):
        """(offsets int32[4^k+2], positions int32[n_valid]) on the device, built on first use."""
        import ctypes
        import torch
        if k not in self._idx:
            if not 2 <= k <= 15:
                raise NotImplementedError("locate index supports 2 <= k <= 15 (got %d)" % k)
            P = _lib.ptr
            with torch.cuda.device(self.dev):
                off = torch.empty(int(self.lib.swga_index_offsets_words(k)), dtype=torch.int32, device=self.dev)
                pos = torch.empty(max(self.n, 1), dtype=torch.int32, device=self.dev)
                nv = ctypes.c_uint64(0)
                stream = torch.cuda.current_stream().cuda_stream
                _lib.check(self.lib.swga_build_index(self.ctx.h, P(self.packed), P(self.brk), self.n, k, P(off), P(pos),
                                                     ctypes.byref(nv), stream))
                torch.cuda.synchronize()
            self._idx[k] = (off, pos[: nv.value].clone() if nv.value < self.n else pos, nv.value)
        return self._idx[k]

    def locate_codes(self, k, codes):
        """Sites of each code (ascending linear positions).  Returns (hits int32 tensor on device,
        hit_off np.uint64[Q+1])."""
        import torch
        P = _lib.ptr
        off, pos, _ = self.index(k)
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        q = len(codes)
        with torch.cuda.device(self.dev):
            stream = torch.cuda.current_stream().cuda_stream
            d_codes = torch.from_numpy(codes.view(np.int32)).to(self.dev)
            d_counts = torch.empty(max(q, 1), dtype=torch.int32, device=self.dev)
            _lib.check(self.lib.swga_locate_counts(self.ctx.h, k, P(off), P(d_codes), q, P(d_counts), stream))
            counts = d_counts[:q].cpu().numpy().view(np.uint32).astype(np.uint64)
            hit_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
            hits = torch.empty(max(int(hit_off[-1]), 1), dtype=torch.int32, device=self.dev)
            d_off = torch.from_numpy(hit_off.view(np.int64)).to(self.dev)
            _lib.check(self.lib.swga_locate_gather(self.ctx.h, k, P(off), P(pos), P(d_codes), q, P(d_off), P(hits),
                                                   stream))
        return hits, hit_off

    def binding_sites(self, kmer):
        """{record: [0-based starts of kmer ascending] + [starts of revcomp(kmer) ascending]} for every
        record, in file order (swga/locate.py:39-51)."""
        kmer = str(kmer)
        g = self.genome
        if g.n_records == 0:
            raise ValueError("No locations for {} found in fg genome!".format(kmer))
        fwd = kmer.upper()
        rc = revcomp(kmer).upper()                 # KeyError for non-ACGT / lower case, like the reference
        if any(c not in "ACGT" for c in fwd):      # unreachable after revcomp(); kept for clarity
            raise KeyError(fwd)
        k = len(fwd)
        hits, hit_off = self.locate_codes(k, [kmers.kmer_to_code(fwd), kmers.kmer_to_code(rc)])
        h = hits[: int(hit_off[-1])].cpu().numpy().view(np.uint32).astype(np.int64)
        f, r = h[: int(hit_off[1])], h[int(hit_off[1]):]
        rs = self.rec_starts.astype(np.int64)
        locations = {}
        fi = np.searchsorted(f, rs)
        ri = np.searchsorted(r, rs)
        for rec, name in enumerate(g.names):
            a = f[fi[rec]:fi[rec + 1]] - rs[rec]
            b = r[ri[rec]:ri[rec + 1]] - rs[rec]
            locations[name] = a.tolist() + b.tolist()
        return locations

    def chromosome_ends(self):
        ends = {}
        rs = self.rec_starts.astype(np.int64)
        for rec, name in enumerate(self.genome.names):
            ends[name] = [int(rs[rec]), int(rs[rec + 1]) - 1]
        return ends


_INDEX_CACHE = {}


def genome_index(genome_fp, device=None):
    """GenomeIndex of a FASTA file, cached per (path, mtime, size): `Primer._update_locations`
    (swga/workspace.py:150-152) calls binding_sites once per primer."""
    st = os.stat(genome_fp)
    key = (os.path.abspath(genome_fp), st.st_mtime_ns, st.st_size)
    if key not in _INDEX_CACHE:
        _INDEX_CACHE.clear()
        _INDEX_CACHE[key] = GenomeIndex(genome_fp, device)
    return _INDEX_CACHE[key]


def binding_sites(kmer, genome_fp):
    return genome_index(genome_fp).binding_sites(kmer)


def chromosome_ends(genome_fp):
    """Starts/ends of each record on the concatenation of all records, 0-based (swga/locate.py:59-73).
    Host-only: reads the FASTA, no GPU needed."""
    g = fasta.read_fasta(genome_fp)
    ends = {}
    rs = g.rec_starts.astype(np.int64)
    for rec, name in enumerate(g.names):
        ends[name] = [int(rs[rec]), int(rs[rec + 1]) - 1]
    return ends


def linearize_binding_sites(primers, chr_ends):
    """swga/locate.py:22-36 for objects carrying `.locations` ({record: [pos]}): the de-duplicated
    union of every site shifted to linear coordinates plus both ends of every record the primers
    list.  Host helper for single sets; the batched device form is swga_b200.score.SiteLists."""
    new_locs = []
    for primer in primers:
        locations = primer.locations if hasattr(primer, "locations") else primer
        for rec, locs in locations.items():
            chr_start, chr_end = chr_ends[rec]
            new_locs += [l + chr_start for l in locs] + [chr_start, chr_end]
    new_locs = list(set(new_locs))
    if new_locs == []:
        raise ValueError("Binding sites for primers not found!")
    return new_locs
