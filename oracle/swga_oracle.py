"""oracle/swga_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (Python 3, Python-2 arithmetic preserved) of the reference's hot path:

* counting   : ctypes binding of oracle/dsk_oracle.c (restates ext/dsk, see that file) plus
               `_parse_kmer_binary` (swga/kmers.py:94-115) and `primer_dict`
               (swga/commands/count.py:117-135)
* locate     : swga/locate.py:12-90 with the pyfaidx record semantics the reference relies on
* statistics : swga/stats.py:10-54 (gini's `fair_area` is Python-2 integer floor division)
* scoring    : swga/score.py:16-98

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  The product (swga_b200/) never does.

Parity status: PINNED -- against the reference's own known-answer tests
(swga/tests/test_locate.py:12-41, swga/tests/test_score.py:4-29, swga/tests/test_swga.py:132-154),
against the real DSK binary (oracle/_ref/dsk) via tests/golden/*, and SURVEY.md App. B.
"""
import ctypes
import math
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all", "ref"])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.isfile(so):
            build()
        L = ctypes.CDLL(so)
        u8p, u64p, u32p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_uint64, ctypes.c_uint32))
        L.orc_count_k.argtypes = [u8p, u64p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, u32p]
        L.orc_count_k.restype = None
        L.orc_count_allk.argtypes = [u8p, u64p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, u32p]
        L.orc_count_allk.restype = None
        L.orc_fasta_parse.argtypes = [u8p, ctypes.c_uint64, u8p, u64p, ctypes.c_uint64]
        L.orc_fasta_parse.restype = ctypes.c_int64
        L.orc_dsk_mode.argtypes = [u64p, ctypes.c_uint64]
        L.orc_dsk_mode.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# --------------------------------------------------------------------------- counting

def fasta_parse_dsk(data):
    """DSK's view of a FASTA image: (joined bases uint8[N], rec_starts uint64[R+1]).
    Restates ext/dsk/minia/Bank.cpp:150-209 (see dsk_oracle.c)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    buf = np.ascontiguousarray(buf)
    seq = np.empty(max(len(buf), 1), dtype=np.uint8)
    max_rec = int(np.count_nonzero(buf == ord(">")) + np.count_nonzero(buf == ord("@"))) + 1
    starts = np.zeros(max_rec + 1, dtype=np.uint64)
    n = lib().orc_fasta_parse(_p(buf, ctypes.c_uint8), len(buf), _p(seq, ctypes.c_uint8),
                              _p(starts, ctypes.c_uint64), max_rec)
    if n < 0:
        raise ValueError("orc_fasta_parse failed: %d" % n)
    starts = starts[: n + 1].copy()
    return seq[: int(starts[-1])].copy(), starts


def dsk_mode(rec_starts):
    """1 = mode B (raw, N->G), 0 = mode A (split at 'N').  SortingCount.cpp:81-82."""
    rs = np.ascontiguousarray(rec_starts, dtype=np.uint64)
    return lib().orc_dsk_mode(_p(rs, ctypes.c_uint64), len(rs) - 1)


def count_dense(seq, rec_starts, k, mode_b=None):
    """Dense canonical table (uint32[4^k]) exactly as DSK would count these records."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    rs = np.ascontiguousarray(rec_starts, dtype=np.uint64)
    if mode_b is None:
        mode_b = dsk_mode(rs)
    table = np.zeros(4 ** k, dtype=np.uint32)
    if len(seq) == 0:
        seq = np.zeros(1, dtype=np.uint8)
    lib().orc_count_k(_p(seq, ctypes.c_uint8), _p(rs, ctypes.c_uint64), len(rs) - 1, int(mode_b), k,
                      _p(table, ctypes.c_uint32))
    return table


def count_dense_allk(seq, rec_starts, kmin, kmax, mode_b=None):
    """Concatenated dense canonical tables for k = kmin..kmax."""
    return np.concatenate([count_dense(seq, rec_starts, k, mode_b) for k in range(kmin, kmax + 1)])


def count_dense_allk_parallel(seq, rec_starts, kmin, kmax, mode_b, threads=None):
    """count_dense_allk for a genome too large for one core: the records are counted independently (a k-mer
    never spans two records, Bank.cpp:889-902) by `threads` host threads -- the C calls release the GIL -- and
    the per-record tables are summed.  Same function, same answer; used for the 3.1 Gbp parity gate."""
    from concurrent.futures import ThreadPoolExecutor
    import threading
    rs = np.ascontiguousarray(rec_starts, dtype=np.uint64)
    threads = threads or os.cpu_count() or 1
    words = sum(4 ** k for k in range(kmin, kmax + 1))
    total = np.zeros(words, dtype=np.uint32)
    lock = threading.Lock()
    lib()

    def one(r):
        a, b = int(rs[r]), int(rs[r + 1])
        t = count_dense_allk(seq[a:b], np.array([0, b - a], dtype=np.uint64), kmin, kmax, mode_b)
        with lock:
            np.add(total, t, out=total)

    order = sorted(range(len(rs) - 1), key=lambda r: -(int(rs[r + 1]) - int(rs[r])))
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, order))
    return total


def code_to_kmer(code, k):
    """ext/dsk/minia/Kmer.cpp:148-162 code2seq with bin2NT = A,C,T,G; first base most significant."""
    return "".join("ACTG"[(int(code) >> (2 * (k - 1 - j))) & 3] for j in range(k))


def kmer_to_code(s):
    """ext/dsk/minia/Kmer.cpp:65-76 codeSeed."""
    x = 0
    for ch in s:
        x = x * 4 + ((ord(ch) >> 1) & 3)
    return x


def count_kmers_dict(data, k, threshold=1):
    """What swga.kmers.count_kmers returns (kmers.py:52-54) for a FASTA image."""
    seq, rs = fasta_parse_dsk(data)
    t = count_dense(seq, rs, k)
    nz = np.nonzero(t >= max(1, threshold))[0]
    return {code_to_kmer(c, k): int(t[c]) for c in nz}


def parse_kmer_binary(fp):
    """swga/kmers.py:94-115 restated: yields (kmer, freq) from a DSK result file."""
    with open(fp, "rb") as f:
        kmer_nbits = struct.unpack("i", f.read(4))[0]
        k = struct.unpack("i", f.read(4))[0]
        nb = kmer_nbits // 8
        while True:
            kb = f.read(nb)
            fb = f.read(4)
            if len(kb) < nb or len(fb) < 4:
                return
            freq = struct.unpack("I", fb)[0]
            kmer = ""
            for i in range(k):
                kmer = "ACTG"[(kb[i // 4] >> (2 * (i % 4))) % 4] + kmer
            yield kmer, freq


def run_dsk(fasta_fp, k, cwd, threshold=1):
    """Run the real DSK binary (oracle/_ref/dsk) the way swga/kmers.py:42-46 does."""
    dsk = os.path.join(_HERE, "_ref", "dsk")
    out = "%s-%dmers" % (os.path.basename(fasta_fp), k)
    subprocess.check_call([dsk, os.path.abspath(fasta_fp), str(k), "-o", out, "-t", str(threshold)],
                          cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.join(cwd, out + ".solid_kmers_binary")


def max_sequential_nt(mer1, mer2):
    """swga/kmers.py:59-91."""
    binding = {"A": "T", "T": "A", "C": "G", "G": "C", "_": False}
    if len(mer2) > len(mer1):
        mer1, mer2 = mer2, mer1
    mer1_len = len(mer1)
    mer2 = mer2[::-1]
    mer1 = mer1.ljust(mer1_len + len(mer2), "_")
    max_bind = 0
    for offset in range(mer1_len):
        consecutive = 0
        for x in range(len(mer2)):
            if binding[mer1[offset + x]] == mer2[x]:
                consecutive += 1
                if consecutive > max_bind:
                    max_bind = consecutive
            else:
                consecutive = 0
    return max_bind


def primer_dict(seq, fg, bg, min_fg_bind, max_bg_bind, max_dimer_bp):
    """swga/commands/count.py:117-135."""
    fg_freq = fg[seq]
    bg_freq = bg.get(seq, 0)
    ratio = fg_freq / float(bg_freq) if bg_freq > 0 else float("inf")
    max_bg_bind = float("inf") if max_bg_bind < 0 else max_bg_bind
    if fg_freq >= min_fg_bind and bg_freq <= max_bg_bind and max_sequential_nt(seq, seq) <= max_dimer_bp:
        return {"seq": seq, "fg_freq": fg_freq, "bg_freq": bg_freq, "ratio": ratio}
    return {}


# --------------------------------------------------------------------------- locate

def fasta_records_pyfaidx(data):
    """The records locate.py sees through pyfaidx (not vendored; > 0.4.5.2, requirements.txt:1):
    key = first whitespace-delimited token of the header (tests/test_locate.py:12-22 rely on
    '> record1' -> 'record1'), sequence = the record's lines joined, case preserved, file order."""
    if isinstance(data, np.ndarray):
        data = data.tobytes()
    text = bytes(data).decode("latin-1")
    recs = []
    name, parts = None, []
    for line in text.split("\n"):
        line = line.rstrip("\r")
        if line.startswith(">"):
            if name is not None:
                recs.append((name, "".join(parts)))
            tok = line[1:].split()
            name = tok[0] if tok else ""
            parts = []
        elif name is not None:
            parts.append(line.strip())
    if name is not None:
        recs.append((name, "".join(parts)))
    return recs


def revcomp(s):
    """swga/locate.py:12-19 (KeyError on non-ACGT)."""
    r = {"A": "T", "T": "A", "G": "C", "C": "G"}
    return "".join([r[i] for i in s][::-1])


def substr_indices(substring, string):
    """swga/locate.py:76-90: overlapping find; pattern upper-cased, genome not."""
    locations = []
    start = 0
    substring = substring.upper()
    while True:
        start = string.find(substring, start) + 1
        if start > 0:
            locations.append(start - 1)
        else:
            return locations


def binding_sites(kmer, records):
    """swga/locate.py:39-51 on pre-parsed records [(name, seq)]."""
    locations = {}
    kmer = str(kmer)
    for name, seq in records:
        locations[name] = substr_indices(kmer, seq)
        locations[name] += substr_indices(revcomp(kmer), seq)
    if locations == {}:
        raise ValueError("No locations for {} found in fg genome!".format(kmer))
    return locations


def chromosome_ends(records):
    """swga/locate.py:59-73."""
    len_so_far = 0
    chr_ends = {}
    for name, seq in records:
        chr_len = len(seq)
        chr_ends[name] = [len_so_far, chr_len + len_so_far - 1]
        len_so_far += chr_len
    return chr_ends


def linearize_binding_sites(primer_locations, chr_ends):
    """swga/locate.py:22-36; primer_locations = list of {record: [pos]} (Primer.locations)."""
    new_locs = []
    for locations in primer_locations:
        for rec, locs in locations.items():
            chr_start, chr_end = chr_ends[rec]
            new_locs += [l + chr_start for l in locs] + [chr_start, chr_end]
    new_locs = list(set(new_locs))
    if new_locs == []:
        raise ValueError("Binding sites for primers not found!")
    return list(set(new_locs))


# --------------------------------------------------------------------------- stats (Python 2!)

def mean(values):
    """swga/stats.py:10-12."""
    return float(sum(values)) / len(values) if len(values) > 0 else float("nan")


def stdev(values):
    """swga/stats.py:15-19 (n == 1 divides by zero, as in the reference)."""
    n = float(len(values))
    mu = sum(values) / n
    return math.sqrt(sum((x - mu) ** 2 for x in values) / (n - 1))


def seq_diff(seq):
    """swga/stats.py:22-33 (sorts in place)."""
    seq.sort()
    diffs = []
    for i in range(len(seq) - 1):
        diff = seq[i + 1] - seq[i]
        assert diff >= 0
        diffs.append(diff)
    return diffs


def gini(distances):
    """swga/stats.py:36-54.  `height * len(distances) / 2` is int/int under Python 2 (no
    `from __future__ import division` in stats.py:1-8) => floor division."""
    distances.sort()
    height = area = 0
    for value in distances:
        height += value
        area += height - value / 2.0
    fair_area = height * len(distances) // 2
    return (fair_area - area) / fair_area


# --------------------------------------------------------------------------- scoring

def read_set_finder_line(line):
    """swga/score.py:16-24."""
    primer_set, weight = line.strip("\n").split(" ")
    return [int(_) for _ in primer_set.split(",")], float(weight)


def default_score_set(expression, primer_set, primer_locs, max_dist, bg_dist_mean):
    """swga/score.py:27-53."""
    binding_distances = seq_diff(primer_locs)
    namespace = {
        "set_size": len(primer_set),
        "fg_dist_mean": mean(binding_distances),
        "fg_dist_std": stdev(binding_distances),
        "fg_dist_gini": gini(binding_distances),
        "bg_dist_mean": bg_dist_mean,
        "fg_max_dist": max_dist,
    }
    score = eval(expression, {"__builtins__": None}, dict(namespace))
    return score, namespace


def calculate_bg_dist_mean(bg_freqs, bg_length):
    """swga/score.py:56-70."""
    total = sum(bg_freqs)
    return float("inf") if total == 0 else float(bg_length) / total


DEFAULT_EXPRESSION = "(fg_dist_mean * fg_dist_gini) / (bg_dist_mean)"   # find_sets.yaml:39-40


def score_set(primer_locations, max_fg_bind_dist, bg_dist_mean, chr_ends,
              expression=DEFAULT_EXPRESSION, interactive=False):
    """swga/score.py:73-98; primer_locations = list of {record: [pos]} for the set's primers."""
    binding_locations = linearize_binding_sites(primer_locations, chr_ends)
    max_dist = max(seq_diff(binding_locations))
    if not interactive and max_dist > max_fg_bind_dist:
        return False, {}, max_dist
    set_score, variables = default_score_set(expression, primer_locations, binding_locations,
                                             max_dist, bg_dist_mean)
    return set_score, variables, max_dist


# --------------------------------------------------------------------------- exports (cold path)

def lorenz(distances):
    """swga/stats.py:57-81."""
    distances.sort()
    height = 0
    lz = []
    for dist in distances:
        height += dist
        lz.append(height)
    _max = float(max(lz))
    return [l / _max for l in lz]


def bedgraph_hits_per_record(primer_locations, chr_ends, window_size, step_size=None):
    """BedGraph._hits_per_record, swga/export.py:91-133, loop for loop (Python-2 `/` on ints is `//`);
    primer_locations = list of {record: [pos]}.  Returns [(record, midpoint, hits)]."""
    from collections import Counter
    window_size = int(window_size)
    step_size = int(step_size) if step_size else window_size // 5
    out = []
    for record_name, ends in chr_ends.items():
        record_length = ends[1] + 1
        this_window_size = window_size
        if this_window_size > record_length:
            this_window_size = record_length
        this_step_size = step_size
        if this_step_size > this_window_size:
            this_step_size = this_window_size
        counter = Counter()
        for locations in primer_locations:
            counter.update(Counter(locations[record_name]))
        for start in range(0, int(record_length - this_window_size), this_step_size):
            end = start + this_window_size
            midpoint = (end + start) // 2
            hits = sum([counter[i] for i in range(start, end)])
            out.append((record_name, midpoint, hits))
    return out


def py2_float_str(x):
    """Python 2's str(float) (floatobject.c float_str: PyFloat_STR_PRECISION = 12, '.0' appended to integers)."""
    s = "%.12g" % x
    if "." not in s and "e" not in s and "n" not in s and "i" not in s:
        s += ".0"
    return s
