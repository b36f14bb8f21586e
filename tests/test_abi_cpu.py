"""CPU-only checks of the boundary: the shared library loads, exports every symbol include/swga_b200.h
declares (and the ctypes table lists exactly those), host-only entry points work without a GPU, and
compute entry points fail loudly without one (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "swga_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(swga_[a-z0-9_]+)\s*\(", txt))


@pytest.fixture(scope="module")
def lib():
    from swga_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from swga_b200 import _lib
    declared = header_symbols()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libswga_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_only_sm100a_code_is_embedded():
    import subprocess
    from swga_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], stdout=subprocess.PIPE).stdout.decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_sizes(lib):
    assert lib.swga_tables_words(5, 12) == 22369280               # 89,477,120 B
    assert lib.swga_table_offset(5, 5) == 0 and lib.swga_table_offset(5, 7) == 4 ** 5 + 4 ** 6
    assert lib.swga_packed_words(32) == 4 and lib.swga_brk_words(33) == 3
    assert lib.swga_index_offsets_words(12) == 4 ** 12 + 2


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from swga_b200 import _lib, kmers, fasta
    with pytest.raises(_lib.SwgaError):
        _lib.Context(0)
    g = fasta.parse_fasta_bytes(b">x\nACGTACGTACGT\n")
    with pytest.raises(_lib.SwgaError):
        kmers.count_all(g, 3, 5)


def test_fasta_reader_matches_oracle_reader(lib, oracle):
    from swga_b200 import fasta
    from helpers import golden_case_bytes
    extra = [b"", b">a", b">a\n", b">a\r\nAC\r\n\r\nGT\r\n", b">a b c\nACGT\n@q\nTT\n>z\n\n\nAC", b">x\n\rACGT\n",
             b">x\nAC\xffGT\n\xffhdr\nAAA\n", b">only\nA\r\n"]
    fastq = [b"@r1\nACGT\n+\nIIII\n", b"@r1 c\nACGT\n+r1\nIIII", b"@r1\nACGT\n+", b"@r1\nAC\nGT\n+\n@@\n>>\n@r2\nTT\n+\nII\n",
             b">f\nACGT\n+\nII\nIIXX@h\nGG\n", b"@a 0123456789\nACGTACGTAC\n+\nI\r\n\r\n\n>x\nAC\n", b"@\n+\n\n@b\nA\r\n+\nI"]
    for data in [golden_case_bytes("edge_a"), golden_case_bytes("edge_q"), golden_case_bytes("synth_mode_a")] + extra + fastq:
        g = fasta.parse_fasta_bytes(data)
        seq, rs = oracle.fasta_parse_dsk(data)
        assert np.array_equal(g.seq, seq) and np.array_equal(g.rec_starts, rs), data[:40]
    with pytest.raises(Exception):
        fasta.parse_fasta_bytes(b"not fasta\n>x\nACGT\n")         # DSK reads this as a list of file names
    g = fasta.parse_fasta_bytes(b"@r1 x\nACGT\n+\nIIII\n@r2\nGG\n+\n@I\n")       # FASTQ, Bank.cpp:192-201
    assert g.names == ["r1", "r2"] and bytes(g.seq) == b"ACGTGG"
    g = fasta.parse_fasta_bytes(b"> record1 desc\nAC\n>r2\tx\nGT\n")
    assert g.names == ["record1", "r2"]


def test_threaded_fasta_reader_equals_oracle_reader(lib, oracle, monkeypatch):
    """The reader cuts the image at line starts into ranges scanned by up to 16 threads; with 1-byte ranges forced,
    random images (headers, empty lines, CRLF, lone CR lines, long lines) must still parse like DSK's reader."""
    from swga_b200 import fasta
    rng = np.random.RandomState(99)
    pieces = [b">h\n", b">h x y\n", b"@q\n", b"\n", b"\r\n", b"ACGT\n", b"ACGTNNAC\r\n", b"A\n", b"\r", b"acgtRYK\n",
              b"G" * 300 + b"\n", b"TTTT", b">\n"]
    monkeypatch.setenv("SWGA_FASTA_CHUNK", "1")
    for it in range(300):
        body = b"".join(pieces[i] for i in rng.randint(0, len(pieces), int(rng.randint(0, 40))))
        data = b">first\n" + body
        g = fasta.parse_fasta_bytes(data)
        seq, rs = oracle.fasta_parse_dsk(data)
        assert np.array_equal(g.seq, seq) and np.array_equal(g.rec_starts, rs), (it, data)
    # FASTQ records ('+' at a line start) anywhere in the image: the byte-level reader takes over
    qpieces = pieces + [b"+\n", b"+c\r\n", b"II@I\n", b"I>II\r\n", b"@q some words\n", b"IIIIIIII\n"]
    for it in range(400):
        body = b"".join(qpieces[i] for i in rng.randint(0, len(qpieces), int(rng.randint(0, 40))))
        data = (b">first\n", b"@first c\n")[it & 1] + body
        g = fasta.parse_fasta_bytes(data)
        seq, rs = oracle.fasta_parse_dsk(data)
        assert np.array_equal(g.seq, seq) and np.array_equal(g.rec_starts, rs), (it, data)
    monkeypatch.setenv("SWGA_FASTA_CHUNK", "64")
    big = b"".join([b">r%d d\n" % i + b"ACGTTGCA" * int(rng.randint(0, 40)) + b"\n" + b"GG\r\n" * int(rng.randint(0, 3))
                    for i in range(400)])
    g = fasta.parse_fasta_bytes(big)
    seq, rs = oracle.fasta_parse_dsk(big)
    assert np.array_equal(g.seq, seq) and np.array_equal(g.rec_starts, rs) and g.names[:2] == ["r0", "r1"]


def test_gzip_fasta_is_read_like_plain(tmp_path):
    """DSK opens its input with gzopen (Bank.cpp:39), which reads plain and gzip files alike."""
    import gzip
    from swga_b200 import fasta
    data = b">a x\nACGTN\nacgt\n>b\nGG\n"
    (tmp_path / "p.fa").write_bytes(data)
    with gzip.open(str(tmp_path / "z.fa.gz"), "wb") as f:
        f.write(data)
    p, z = fasta.read_fasta(str(tmp_path / "p.fa")), fasta.read_fasta(str(tmp_path / "z.fa.gz"))
    assert np.array_equal(p.seq, z.seq) and np.array_equal(p.rec_starts, z.rec_starts) and p.names == z.names == ["a", "b"]


def test_dsk_file_roundtrip(lib, oracle, tmp_path):
    from swga_b200 import _lib, kmers
    k = 6
    rng = np.random.RandomState(1)
    table = np.where(rng.rand(4 ** k) < 0.1, rng.randint(1, 1000, 4 ** k), 0).astype(np.uint32)
    fp = str(tmp_path / "g.fa-6mers.solid_kmers_binary")
    n = ctypes.c_uint64()
    _lib.check(lib.swga_write_dsk_file(fp.encode(), k, _lib.ptr(table), 3, ctypes.byref(n)))
    got = dict(kmers._parse_kmer_binary(fp))
    want = {oracle.code_to_kmer(c, k): int(table[c]) for c in np.nonzero(table >= 3)[0]}
    assert got == want and n.value == len(want)
    assert got == dict(oracle.parse_kmer_binary(fp))             # swga/kmers.py:94-115 reads it the same way
    # count_kmers re-uses an existing cache file without touching the GPU (kmers.py:34-40)
    genome = tmp_path / "g.fa"
    genome.write_bytes(b">x\nACGT\n")
    assert kmers.count_kmers(6, str(genome), str(tmp_path), threshold=5) == {s: c for s, c in want.items() if c >= 5}


def test_host_helpers():
    from swga_b200 import kmers, locate, score
    assert kmers.max_sequential_nt("ATGC", "GCAT") == 4 and kmers.max_sequential_nt("AAAA", "AAAA") == 0
    assert kmers.parse_kmer_file(["ACGT 3\n", "TTTT\t1\n"]) == ["ACGT", "TTTT"]
    assert locate.revcomp("GCAT") == "ATGC" and locate.substr_indices("aa", "AAAA") == [0, 1, 2]
    assert score.read_set_finder_line("1,2,3 2.500000\n") == ([1, 2, 3], 2.5)
    assert score.shard_sets(10, 3, 4) == (9, 10) and score.shard_sets(10, 0, 4) == (0, 3)
