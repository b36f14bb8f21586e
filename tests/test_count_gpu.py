"""GPU parity for kernels (a)+(b): the CUDA path through the C ABI vs the CPU oracle, bit-exact,
on the golden cases (real-DSK answers), seeded synthetic genomes, and size-independent properties."""
import numpy as np
import pytest

from helpers import golden_case_bytes, table_digest

pytestmark = pytest.mark.gpu

# "ref:<file>" = the reference's own swga/tests/data fixtures (copies under tests/golden/ref/), whose real-DSK
# answers are in dsk_golden.json: bburgdorferi_wgs.fa has IUPAC letters and two 'N' (mode A)
REF_CASES = ["ref:test.fasta", "ref:bburgdorferi_1k.fa", "ref:ecoli_2k.fa", "ref:bburgdorferi_wgs.fa",
             "ref:simple_fg_genome.fa"]
CASES = ["edge_a", "edge_q", "synth_mode_a", "synth_mode_b"] + REF_CASES


def k_range(case):
    return (3, 12) if case in ("edge_a", "edge_q") else (4, 12) if case.startswith("ref:") else (5, 12)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


def u32(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("case", CASES)
def test_host_call_matches_oracle_and_real_dsk(case, oracle, dsk_golden, torch_cuda):
    from swga_b200 import fasta, kmers
    data = golden_case_bytes(case)
    g = fasta.parse_fasta_bytes(data)
    seq, rs = oracle.fasta_parse_dsk(data)
    assert np.array_equal(g.seq, seq) and np.array_equal(g.rec_starts, rs)
    kmin, kmax = k_range(case)
    tabs = kmers.count_all(g, kmin, kmax)
    for k in range(kmin, kmax + 1):
        want = oracle.count_dense(seq, rs, k)
        got = tabs.table(k)
        assert np.array_equal(got, want), (case, k, int(np.count_nonzero(got != want)))
        gold = dsk_golden[case].get(str(k))
        if gold:
            assert table_digest(got, k) == (gold["distinct"], gold["sum"], gold["md5"])


@pytest.mark.parametrize("kmin,kmax", [(5, 12), (2, 4), (7, 7), (6, 13), (12, 12)])
def test_k_ranges(kmin, kmax, oracle, torch_cuda):
    from swga_b200 import fasta, kmers
    data = golden_case_bytes("synth_mode_a")
    g = fasta.parse_fasta_bytes(data)
    tabs = kmers.count_all(g, kmin, kmax)
    for k in range(kmin, kmax + 1):
        assert np.array_equal(tabs.table(k), oracle.count_dense(g.seq, g.rec_starts, k)), k


def test_empty_and_tiny_inputs(oracle, torch_cuda):
    from swga_b200 import fasta, kmers
    for fa in (b"", b">only_header\n", b">x\nACG\n", b">x\nACGTA\n", b">a\nAC\n>b\nGT\n>c\nACGTAC\n", b">n\nNNNNNNNN\n"):
        g = fasta.parse_fasta_bytes(fa)
        tabs = kmers.count_all(g, 3, 6)
        for k in range(3, 7):
            assert np.array_equal(tabs.table(k), oracle.count_dense(g.seq, g.rec_starts, k)), (fa, k)


def test_device_path_synthetic_c1(oracle, torch_cuda):
    """BASELINE config 1 sizes (1 Mbp fg, 10 Mbp bg), generated on the device by k_synth and on the
    CPU by swga_b200.synth -- the generators must agree bit for bit, then the counts must."""
    torch = torch_cuda
    from swga_b200 import _lib, kmers, synth
    lib, ctx = _lib.load(), _lib.context(0)
    for name in ("C1_fg", "C1_bg"):
        seed, n, gc, n_rec = synth.CONFIGS[name]
        ta, tc, tg = synth.thresholds(gc)
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        _lib.check(lib.swga_synth_bases(ctx.h, seed, 0, n, ta, tc, tg, _lib.ptr(d), None))
        torch.cuda.synchronize()
        h = synth.bases(seed, 0, n, gc)
        assert np.array_equal(d.cpu().numpy(), h)
        rs = synth.equal_records(n, n_rec)
        fwd, canon = kmers.count_all_device(d, rs, mode_b=n > 1000000)
        torch.cuda.synchronize()
        got = u32(canon)
        want = oracle.count_dense_allk(h, rs, 5, 12, mode_b=int(n > 1000000))
        assert np.array_equal(got, want)


def test_chunked_pipeline_crosses_chunk_boundaries(oracle, torch_cuda):
    """> 64 Mi bases so swga_count_genome_host runs 2 chunks; records and N runs straddle the cut."""
    from swga_b200 import fasta, kmers, synth
    n = (64 << 20) + 5_000_000
    seq = synth.bases(77, 0, n, 0.45)
    cut = 64 << 20
    seq[cut - 3:cut + 2] = ord("N")
    rs = np.array([0, 1000, cut - 40, cut + 5, cut + 6, n - 7, n], dtype=np.uint64)
    for mode_a in (True, False):
        starts = rs if not mode_a else np.array(sorted(set(list(range(0, n, 900_000)) + [int(x) for x in rs])),
                                                 dtype=np.uint64)
        g = fasta.Genome(seq, starts, ["r%d" % i for i in range(len(starts) - 1)])
        assert g.dsk_mode_b == (not mode_a)
        tabs = kmers.count_all(g, 5, 12)
        for k in (5, 8, 11, 12):
            assert np.array_equal(tabs.table(k), oracle.count_dense(seq, starts, k)), (mode_a, k)


def test_sharded_accumulate_equals_whole(oracle, torch_cuda):
    """Chunk + right-halo sharding as swga_b200.dist does it, emulated on one GPU: accumulating the
    shards of 2, 4 and 8 'ranks' into one table block equals the unsharded count."""
    torch = torch_cuda
    from swga_b200 import dist, kmers, synth
    n = 3_000_017
    seq = synth.bases(5, 0, n, 0.194)
    seq[1_500_000:1_500_010] = ord("N")
    rs = np.array([0, 17, 750_001, 750_002, 2_250_000, n], dtype=np.uint64)
    want = oracle.count_dense_allk(seq, rs, 5, 12, mode_b=0)
    for world in (1, 2, 4, 8):
        total = None
        for rank in range(world):
            sh = dist.shard_genome(seq, rs, rank, world, kmax=12)
            d = torch.from_numpy(sh.seq.copy()).cuda()
            fwd, _ = kmers.count_all_device(d, sh.rec_starts, mode_b=False, n_starts=sh.n_starts, finalize=False)
            total = fwd.clone() if total is None else total + fwd
        canon = dist.finalize_tables(total, 5, 12)
        torch.cuda.synchronize()
        assert np.array_equal(u32(canon), want), world


def test_properties_at_scale(torch_cuda):
    """Size-independent checks on a 400 Mbp device-generated genome (no CPU oracle at this size):
    sum of counts per k == sum over records of max(0, len-k+1); every non-canonical slot is 0;
    table k is the marginal of table k+1 up to the per-record tails."""
    torch = torch_cuda
    from swga_b200 import _lib, kmers, synth
    lib, ctx = _lib.load(), _lib.context(0)
    n, n_rec, gc = 400_000_000, 7, 0.41
    ta, tc, tg = synth.thresholds(gc)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    _lib.check(lib.swga_synth_bases(ctx.h, 4, 0, n, ta, tc, tg, _lib.ptr(d), None))
    rs = synth.equal_records(n, n_rec)
    fwd, canon = kmers.count_all_device(d, rs, mode_b=True)
    torch.cuda.synchronize()
    fwd, canon = u32(fwd), u32(canon)
    lens = np.diff(rs.astype(np.int64))
    for k in range(5, 13):
        o0, o1 = int(lib.swga_table_offset(5, k)), int(lib.swga_table_offset(5, k + 1))
        expect = int(np.maximum(0, lens - k + 1).sum())
        assert int(fwd[o0:o1].astype(np.int64).sum()) == expect
        assert int(canon[o0:o1].astype(np.int64).sum()) == expect
        codes = np.arange(4 ** k, dtype=np.uint64)
        rc = np.zeros_like(codes)
        c = codes.copy()
        for _ in range(k):
            rc = (rc << np.uint64(2)) | ((c & np.uint64(3)) ^ np.uint64(2))
            c >>= np.uint64(2)
        assert not canon[o0:o1][codes > rc].any()
        if k < 12:
            up = fwd[o1:o1 + 4 ** (k + 1)].astype(np.int64).reshape(-1, 4).sum(1)
            diff = fwd[o0:o1].astype(np.int64) - up
            assert (diff >= 0).all() and int(diff.sum()) == n_rec      # one tail k-mer per record per level


@pytest.mark.parametrize("kmin,kmax", [(5, 12), (8, 11), (11, 11), (12, 12)])
def test_partitioned_kernel_equals_direct_and_oracle(kmin, kmax, oracle, torch_cuda):
    """The partitioned shared-memory counting path (csrc/count_part.cu) against the direct red.global
    kernel and the oracle, on a genome with N runs, IUPAC bytes and short records (mode A)."""
    torch = torch_cuda
    from swga_b200 import _lib, kmers, synth
    lib, ctx = _lib.load(), _lib.context(0)
    n = 6_000_011
    seq = synth.bases(41, 0, n, 0.25)
    runs = synth.splitmix64(np.arange(300, dtype=np.uint64) + np.uint64(13 << 40))
    for r in runs:
        p = int(r % np.uint64(n))
        seq[p:p + int((r >> np.uint64(40)) % np.uint64(40)) + 1] = ord("N")
    seq[123_456] = ord("R")
    rs = np.array(sorted({0, 5, 999_999, 1_000_000, 3_000_017, 3_000_020, n}), dtype=np.uint64)
    d = torch.from_numpy(seq).cuda()
    out = {}
    try:
        # 1 = direct kernel on k_pack's output; 2 = partitioned kernels reading the ASCII bases themselves;
        # 3 = partitioned kernels on k_pack's output (round-1 data flow, kept for A/B timing)
        for impl in (1, 2, 3):
            _lib.check(lib.swga_ctx_set_count_impl(ctx.h, min(impl, 2)))
            _lib.check(lib.swga_ctx_set_count_via_pack(ctx.h, int(impl == 3)))
            for mode_b in (False, True):
                _, canon = kmers.count_all_device(d, rs, mode_b=mode_b, kmin=kmin, kmax=kmax)
                torch.cuda.synchronize()
                out[impl, mode_b] = u32(canon)
    finally:
        _lib.check(lib.swga_ctx_set_count_impl(ctx.h, 0))
        _lib.check(lib.swga_ctx_set_count_via_pack(ctx.h, 0))
    for mode_b in (False, True):
        assert np.array_equal(out[1, mode_b], out[2, mode_b]) and np.array_equal(out[3, mode_b], out[2, mode_b])
        want = oracle.count_dense_allk(seq, rs, kmin, kmax, mode_b=int(mode_b))
        assert np.array_equal(out[2, mode_b], want)


@pytest.mark.parametrize("mode_b", [False, True])
def test_partitioned_kernel_on_hostile_composition(mode_b, oracle, torch_cuda):
    """The partitioned path sizes its buckets and its per-tile staging ranges from a strided SAMPLE of the
    genome (one 1024-start block in every `stride`, stride <= 64).  This genome is built to make the sample wrong in every
    way the kernel has an exit for: composition that changes along the genome (staging ranges overflow
    into the per-tile list), a long two-base repeat (the list overflows too), poly-A / N runs (uniform
    groups; N -> G in mode B), and a repeat placed only where the sample does not look (global buckets
    overflow into direct table updates).  Counts must still equal the direct kernel's and the oracle's."""
    torch = torch_cuda
    from swga_b200 import _lib, kmers, synth
    lib, ctx = _lib.load(), _lib.context(0)
    n = 24_000_000
    seq = np.concatenate([synth.bases(51, 0, n // 3, 0.15), synth.bases(52, 0, n // 3, 0.80),
                          synth.bases(53, 0, n - 2 * (n // 3), 0.45)])
    seq[2_000_000:2_300_000] = np.frombuffer(b"AC" * 150_000, dtype=np.uint8)
    seq[5_000_000:5_200_000] = ord("A")
    seq[9_000_000:9_150_000] = ord("N")
    unit = np.frombuffer(b"ACGTTGCAAGGCTT" * 1200, dtype=np.uint8)
    pos = np.arange(16_000_000, 23_000_000)
    stride = min(64, max(1, ((n + 31) // 32) >> 15))           # the sampler's stride for this size (count_part.cu)
    hidden = pos[(pos % (1024 * stride)) >= 1100]              # never inside a sampled block
    seq[hidden] = unit[hidden % len(unit)]
    rs = np.array([0, 7_999_999, 8_000_000, 8_000_010, n], dtype=np.uint64)
    d = torch.from_numpy(seq).cuda()
    out = {}
    try:
        for impl in (1, 2):
            _lib.check(lib.swga_ctx_set_count_impl(ctx.h, impl))
            _, canon = kmers.count_all_device(d, rs, mode_b=mode_b, kmin=5, kmax=12)
            torch.cuda.synchronize()
            out[impl] = u32(canon)
        stats = np.zeros(4, dtype=np.uint64)
        _lib.check(lib.swga_count_partition_stats(ctx.h, stats.ctypes.data_as(_lib.c_u64p)))
    finally:
        _lib.check(lib.swga_ctx_set_count_impl(ctx.h, 0))
    # every exit of the kernel was taken: staging range full, global bucket full, uniform groups, break groups
    assert all(int(x) > 0 for x in stats), stats
    assert np.array_equal(out[1], out[2])
    want = oracle.count_dense_allk(seq, rs, 5, 12, mode_b=int(mode_b))
    assert np.array_equal(out[2], want)


def _file_tables(path, kmin, kmax):
    from swga_b200 import kmers
    return kmers.count_all(str(path), kmin, kmax)


@pytest.mark.parametrize("case", CASES)
def test_file_image_call_matches_real_dsk(case, oracle, dsk_golden, torch_cuda, tmp_path):
    """swga_count_fasta_host (FASTA file image in, canonical tables out) on the golden cases: same tables as the
    host-buffer call, the oracle and the real dsk; DSK's read mode is chosen inside the library."""
    from swga_b200 import fasta
    data = golden_case_bytes(case)
    fp = tmp_path / "g.fa"
    fp.write_bytes(data)
    seq, rs = oracle.fasta_parse_dsk(data)
    kmin, kmax = k_range(case)
    tabs = _file_tables(fp, kmin, kmax)
    g = fasta.parse_fasta_bytes(data)
    assert (tabs.n_bases, tabs.n_records, tabs.mode_b) == (len(seq), len(rs) - 1, g.dsk_mode_b)
    for k in range(kmin, kmax + 1):
        got = tabs.table(k)
        assert np.array_equal(got, oracle.count_dense(seq, rs, k)), (case, k)
        gold = dsk_golden[case].get(str(k))
        if gold:
            assert table_digest(got, k) == (gold["distinct"], gold["sum"], gold["md5"])


def test_file_image_chunking(oracle, torch_cuda, tmp_path, monkeypatch):
    """Ranges of a few bytes and chunks of a few dozen bases: every chunk hands its last kmax-1 bases to the next
    one, records / 'N' runs / CRLF / empty lines fall on range and chunk boundaries, some chunks are shorter
    than the halo.  Also the degenerate images."""
    from swga_b200 import fasta, kmers
    rng = np.random.RandomState(5)
    pieces = [b">h\n", b">h x y\n", b"\n", b"\r\n", b"ACGT\n", b"ACGTNNAC\r\n", b"A\n", b"\r", b"acgtRYK\n",
              b"G" * 300 + b"\n", b"TTTT", b">\n", b"ACGTTGCATGCAAGTC" * 9 + b"\n", b"N" * 40 + b"\n"]
    images = [b"", b">only\n", b">x\nACG\n", b">x\nACGTA", b">a\nAC\n>b\nGT\n>c\nACGTAC\n"]
    for it in range(12):
        images.append(b">first\n" + b"".join(pieces[i] for i in rng.randint(0, len(pieces), int(rng.randint(5, 120)))))
    for chunk, stream in (("1", "64"), ("7", "100"), ("64", "1000")):
        monkeypatch.setenv("SWGA_FASTA_CHUNK", chunk)
        monkeypatch.setenv("SWGA_STREAM_CHUNK", stream)
        for i, data in enumerate(images):
            fp = tmp_path / "c.fa"
            fp.write_bytes(data)
            seq, rs = oracle.fasta_parse_dsk(data)
            for mode_b in (0, 1):
                import ctypes
                from swga_b200 import _lib
                lib, ctx = _lib.load(), _lib.context(0)
                image = np.frombuffer(data, dtype=np.uint8)
                out = np.empty(int(lib.swga_tables_words(4, 9)), dtype=np.uint32)
                nb = ctypes.c_uint64(0)
                _lib.check(lib.swga_count_fasta_host(ctx.h, _lib.ptr(image) if len(image) else None, len(image), mode_b,
                                                     4, 9, _lib.ptr(out), ctypes.byref(nb), None, None))
                assert nb.value == len(seq)
                want = oracle.count_dense_allk(seq, rs, 4, 9, mode_b=mode_b)
                assert np.array_equal(out, want), (chunk, stream, i, mode_b, data[:60])


def test_file_image_fastq_record_deep_inside_fasta(oracle, torch_cuda, tmp_path, monkeypatch):
    """A FASTQ record ('+' at a line start) far behind the head of a '>' file: the line-parallel reader has already
    counted chunks when it meets it; swga_count_fasta_host starts over with the byte-level reader."""
    from swga_b200 import kmers, synth
    monkeypatch.setenv("SWGA_FASTA_CHUNK", "4096")
    monkeypatch.setenv("SWGA_STREAM_CHUNK", "65536")
    seq = synth.bases(12, 0, 400_000, 0.4)
    fp = str(tmp_path / "mixed.fa")
    synth.write_fasta(fp, seq, np.array([0, 150_000, 400_000], dtype=np.uint64))
    with open(fp, "ab") as f:
        f.write(b"@q1 x\nACGTTGCAACGTTGCAGGA\n+\n@II>IIIIIIIIIIIIIII\n>tail\nGATTACAGATTACAGATTACA\n")
    data = open(fp, "rb").read()
    want_seq, want_rs = oracle.fasta_parse_dsk(data)
    assert len(want_rs) - 1 == 4 and len(want_seq) == 400_000 + 19 + 21
    got = kmers.count_all(fp, 5, 12)
    assert (got.n_bases, got.n_records, got.mode_b) == (len(want_seq), 4, False)
    assert np.array_equal(got.block, oracle.count_dense_allk(want_seq, want_rs, 5, 12, mode_b=0))


def test_file_image_large_multi_chunk(oracle, torch_cuda, tmp_path):
    """Default range / chunk sizes on a file of several chunks (200 Mbp, 5 records, mode B): equal to the
    host-buffer call on the parsed genome (which the partitioned-kernel tests tie to the oracle)."""
    from swga_b200 import fasta, kmers, synth
    n = 200_000_000
    seq = synth.bases(77, 0, n, 0.41)
    seq[50_000_000:50_003_000] = ord("N")
    rs = np.array([0, 1_500_000, 70_000_000, 70_000_005, 140_000_011, n], dtype=np.uint64)
    fp = str(tmp_path / "big.fa")
    synth.write_fasta(fp, seq, rs)
    a = kmers.count_all(fp, 5, 12)
    g = fasta.read_fasta(fp)
    assert np.array_equal(g.seq, seq) and np.array_equal(g.rec_starts, rs)
    b = kmers.count_all(g, 5, 12)
    assert (a.n_bases, a.n_records, a.mode_b) == (n, 5, True)
    assert np.array_equal(a.block, b.block)
    assert int(a.table(12).astype(np.int64).sum()) > 0


def test_file_image_gzip_and_errors(oracle, torch_cuda, tmp_path):
    """A gzip-compressed FASTA gives the same tables (DSK reads both through gzopen); a file that is not FASTA
    and a FASTQ file are refused with the reader's messages."""
    import gzip
    from swga_b200 import _lib, kmers
    data = golden_case_bytes("synth_mode_a")
    (tmp_path / "p.fa").write_bytes(data)
    with gzip.open(str(tmp_path / "z.fa.gz"), "wb") as f:
        f.write(data)
    a, b = kmers.count_all(str(tmp_path / "p.fa"), 5, 9), kmers.count_all(str(tmp_path / "z.fa.gz"), 5, 9)
    assert np.array_equal(a.block, b.block) and (a.n_bases, a.n_records) == (b.n_bases, b.n_records)
    (tmp_path / "bad.fa").write_bytes(b"not fasta\n>x\nACGT\n")
    with pytest.raises(_lib.SwgaError, match="not a FASTA image"):
        kmers.count_all(str(tmp_path / "bad.fa"), 5, 9)
    (tmp_path / "q.fq").write_bytes(b"@r1\nACGTACGTAC\n+\nIIIIIIIIII\n")                 # FASTQ is read (Bank.cpp:192-201)
    q = kmers.count_all(str(tmp_path / "q.fq"), 5, 9)
    assert (q.n_bases, q.n_records) == (10, 1) and int(q.table(5).sum()) == 6
    (tmp_path / "empty.fa").write_bytes(b"")
    e = kmers.count_all(str(tmp_path / "empty.fa"), 5, 9)
    assert e.n_bases == 0 and not e.block.any()


@pytest.mark.parametrize("n_small,big_len,want_b", [(10000, 1_000_001, True), (10001, 1_000_001, False),
                                                    (0, 1_000_001, True), (0, 1_000_000, False)])
def test_file_image_mode_choice_at_the_10001_record_limit(n_small, big_len, want_b, torch_cuda, tmp_path, monkeypatch):
    """DSK looks at the first 10,001 records only (Bank.cpp:470-494): a record longer than 1 Mbp makes mode B when
    it is among them and is not seen otherwise.  The file-image call decides from the head of the file, range batch
    by range batch; the answer and the tables must equal the read-then-count path's."""
    from swga_b200 import fasta, kmers, synth
    monkeypatch.setenv("SWGA_FASTA_CHUNK", "65536")
    big = synth.bases(31, 0, big_len, 0.5)
    big[500:520] = ord("N")
    with open(str(tmp_path / "m.fa"), "wb") as f:
        for i in range(n_small):
            f.write(b">s%d\nACGTTGCA\n" % i)
        f.write(b">big\n")
        body = big[:1_000_000].reshape(-1, 100)                      # 100-column lines (+ a last line of one base)
        f.write(b"\n".join(bytes(r) for r in body) + b"\n" + bytes(big[1_000_000:]) + b"\n")
    g = fasta.read_fasta(str(tmp_path / "m.fa"))
    assert g.n_records == n_small + 1 and g.dsk_mode_b == want_b
    a = kmers.count_all(str(tmp_path / "m.fa"), 5, 8)
    assert a.mode_b == want_b and (a.n_bases, a.n_records) == (g.n_bases, g.n_records)
    assert np.array_equal(a.block, kmers.count_all(g, 5, 8).block)


# ---- the counting seam itself: swga.kmers.count_kmers(k, genome_fp, cwd, threshold) (swga/kmers.py:23-56) ----

@pytest.mark.parametrize("case", CASES)
def test_count_kmers_seam_fresh_path(case, oracle, dsk_golden, torch_cuda, tmp_path, monkeypatch):
    """The fresh path of the seam -- GPU count -> swga_write_dsk_file -> _parse_kmer_binary -- into an EMPTY cwd:
    the returned dict equals the real DSK's counts (golden) and the oracle's, the written file is a DSK result
    file (re-read by the restated reference parser, swga/kmers.py:94-115), the threshold is applied both in the
    file and in the dict, a second call parses the cache file without counting (kmers.py:34-40), and a failure
    removes the partial file and re-raises (kmers.py:45-50)."""
    from swga_b200 import kmers
    data = golden_case_bytes(case)
    fp = tmp_path / (case.replace("ref:", "") + ".fa")
    fp.write_bytes(data)
    cwd = tmp_path / "swga_tmp"
    cwd.mkdir()
    seq, rs = oracle.fasta_parse_dsk(data)
    for ks, gold in sorted(dsk_golden[case].items(), key=lambda kv: int(kv[0])):
        k = int(ks)
        kmers._TABLE_CACHE.clear()
        got = kmers.count_kmers(k, str(fp), str(cwd), 1)
        want_tab = oracle.count_dense(seq, rs, k)
        want = {oracle.code_to_kmer(c, k): int(want_tab[c]) for c in np.nonzero(want_tab)[0]}
        assert got == want, (case, k)
        assert len(got) == gold["distinct"] and sum(got.values()) == gold["sum"]
        if "counts" in gold:
            assert got == gold["counts"]                           # the real dsk's answer, record for record
        out = cwd / ("%s-%dmers.solid_kmers_binary" % (fp.name, k))
        assert out.is_file()
        assert out.stat().st_size == 8 + 12 * len(got)             # [i32 64][i32 k]([u64 code][u32 count])*
        assert dict(oracle.parse_kmer_binary(str(out))) == got
    # threshold (-t): the file holds only counts >= threshold and so does the dict
    k = 5
    for thr in (2, 3):
        cw = tmp_path / ("thr%d" % thr)
        cw.mkdir()
        got = kmers.count_kmers(k, str(fp), str(cw), thr)
        t = oracle.count_dense(seq, rs, k)
        want = {oracle.code_to_kmer(c, k): int(t[c]) for c in np.nonzero(t >= thr)[0]}
        assert got == want
        assert dict(oracle.parse_kmer_binary(str(cw / ("%s-%dmers.solid_kmers_binary" % (fp.name, k))))) == want
    # second call: the cache file is parsed, nothing is counted; the threshold is re-applied on top of it
    def boom(*a, **kw):
        raise AssertionError("count_kmers recounted although the cache file exists")
    with monkeypatch.context() as m:
        m.setattr(kmers, "_tables_for", boom)
        again = kmers.count_kmers(5, str(fp), str(cwd), 1)
        t = oracle.count_dense(seq, rs, 5)
        assert again == {oracle.code_to_kmer(c, 5): int(t[c]) for c in np.nonzero(t)[0]}
        assert kmers.count_kmers(5, str(fp), str(cwd), 4) == {a: b for a, b in again.items() if b >= 4}
    # failure: the partial file is removed and the exception re-raised
    class Boom(Exception):
        pass
    cw = tmp_path / "fail"
    cw.mkdir()
    partial = cw / ("%s-5mers.solid_kmers_binary" % fp.name)

    def fail_after_partial_write(genome_fp, k_):
        partial.write_bytes(b"\x40\x00\x00\x00")
        raise Boom()
    with monkeypatch.context() as m:
        m.setattr(kmers, "_tables_for", fail_after_partial_write)
        with pytest.raises(Boom):
            kmers.count_kmers(5, str(fp), str(cw), 1)
    assert not list(cw.iterdir())


def test_count_kmers_seam_bad_genome(torch_cuda, tmp_path):
    """swga/kmers.py:45-50: when the counter fails (here: the file is not FASTA) nothing is left in cwd."""
    from swga_b200 import _lib, kmers
    (tmp_path / "bad.fa").write_bytes(b"not fasta\n")
    cwd = tmp_path / "t"
    cwd.mkdir()
    with pytest.raises(_lib.SwgaError):
        kmers.count_kmers(5, str(tmp_path / "bad.fa"), str(cwd), 1)
    assert not list(cwd.iterdir())
    with pytest.raises(AssertionError):
        kmers.count_kmers(5, str(tmp_path / "bad.fa"), str(cwd), 1.0)      # kmers.py:32 threshold must be int


def test_full_size_c2_background_vs_oracle(oracle, torch_cuda):
    """BASELINE.md 4.6 / SURVEY 8(d): the 3.1 Gbp background of configs C2 / C3 (24 records, DSK mode B), counted once
    on the device exactly as bench.py's step does, against the CPU oracle on EVERY table k = 5..12, bit for bit.
    The oracle counts the 24 records on the host cores in parallel (same C function, per-record tables summed).
    The bases the oracle reads are the device generator's, copied back; slices of them are compared with the
    host generator (swga_b200.synth) so that both sides provably count the SURVEY 8(d) genome."""
    import time
    torch = torch_cuda
    from swga_b200 import _lib, kmers, synth
    lib, ctx = _lib.load(), _lib.context(0)
    seed, n, gc, n_rec = synth.CONFIGS["C2_bg"]
    assert synth.CONFIGS["C3_bg"] == synth.CONFIGS["C2_bg"]
    ta, tc, tg = synth.thresholds(gc)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    _lib.check(lib.swga_synth_bases(ctx.h, seed, 0, n, ta, tc, tg, _lib.ptr(d), None))
    rs = synth.equal_records(n, n_rec)
    fwd, canon = kmers.count_all_device(d, rs, mode_b=True)
    torch.cuda.synchronize()
    stats = np.zeros(4, dtype=np.uint64)
    _lib.check(lib.swga_count_partition_stats(ctx.h, stats.ctypes.data_as(_lib.c_u64p)))
    got = u32(canon)
    del fwd, canon
    seq = np.empty(n, dtype=np.uint8)
    step = 1 << 28
    for o in range(0, n, step):
        seq[o:o + step] = d[o:o + step].cpu().numpy()
    del d
    torch.cuda.empty_cache()
    for o in (0, 1_234_567_890, 2 ** 31 - 500_000, 2 ** 31 + 7, n - 1_000_000):
        assert np.array_equal(seq[o:o + 1_000_000], synth.bases(seed, o, min(1_000_000, n - o), gc)), o
    t0 = time.time()
    want = oracle.count_dense_allk_parallel(seq, rs, 5, 12, mode_b=1)
    dt = time.time() - t0
    bad = int(np.count_nonzero(got != want))
    print("3.1 Gbp oracle: %.1f s on %d host threads; partition exits %s; mismatching slots %d" % (
        dt, __import__("os").cpu_count(), stats.tolist(), bad))
    assert bad == 0
    assert int(got[-4 ** 12:].astype(np.int64).sum()) == int(np.maximum(0, np.diff(rs.astype(np.int64)) - 11).sum())


@pytest.mark.parametrize("world,kmin,kmax", [(2, 5, 12), (4, 5, 12), (8, 5, 12), (8, 8, 11), (2, 6, 9), (4, 10, 10)])
def test_merge_finalize_kernel_emulated_ranks(world, kmin, kmax, torch_cuda):
    """csrc/merge.cu on ONE device: `world` blocks stand for the ranks' additive tables and the per-rank kernels run
    one after another (rank r sums, finalizes and writes back the slice r / world of every table into all blocks).
    Afterwards every block must equal swga_count_finalize of the element-wise sum -- the kernel is linear, so
    arbitrary 32-bit contents (wrapping sums included) are a complete test."""
    import ctypes
    torch = torch_cuda
    from swga_b200 import _lib
    lib, ctx = _lib.load(), _lib.context(0)
    ok = ctypes.c_int(0)
    _lib.check(lib.swga_merge_supported(kmin, kmax, world, ctypes.byref(ok)))
    assert ok.value == 1
    words = int(lib.swga_tables_words(kmin, kmax))
    g = torch.Generator(device="cuda").manual_seed(1234 + world + kmax)
    blocks = [torch.randint(-2 ** 31, 2 ** 31 - 1, (words,), dtype=torch.int32, device="cuda", generator=g) for _ in range(world)]
    want = blocks[0].clone()
    for b in blocks[1:]:
        want += b
    # the prefix marginalisation T_k[x] += sum_b T_{k+1}[4x + b] (include/swga_b200.h), written out with torch ops
    offs = [int(lib.swga_table_offset(kmin, k)) for k in range(kmin, kmax + 2)]
    for k in range(kmax - 1, kmin - 1, -1):
        lo, up = offs[k - kmin], offs[k + 1 - kmin]
        want[lo:up] += want[up:offs[k + 2 - kmin]].view(-1, 4).sum(1).to(torch.int32)
    single = blocks[0].clone()                                     # world = 1: swga_count_finalize itself
    for b in blocks[1:]:
        single += b
    _lib.check(lib.swga_count_finalize(ctx.h, kmin, kmax, _lib.ptr(single), None))
    assert torch.equal(single, want)
    ptrs = np.array([b.data_ptr() for b in blocks], dtype=np.uint64)
    for rank in range(world):
        _lib.check(lib.swga_merge_finalize(ctx.h, _lib.ptr(ptrs), world, rank, kmin, kmax, None))
    torch.cuda.synchronize()
    for r, b in enumerate(blocks):
        assert torch.equal(b, want), (r, int((b != want).sum()))
    _lib.check(lib.swga_merge_supported(4, 12, world, ctypes.byref(ok)))
    assert ok.value == 0                                           # more than 7 levels below kmax: not covered
