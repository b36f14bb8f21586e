"""GPU parity for kernels (c) locate and (d) batched set scoring, through the C ABI, against the
Python restatement of swga/locate.py, swga/stats.py and swga/score.py, plus the reference's own
known-answer tests (swga/tests/test_locate.py, test_score.py, test_swga.py:132-154)."""
import functools
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TEST_FASTA = b"> record1\nAAGGAAGGCCTTCCTT\n> record2\nAAAATTTT\n> record3\nAAGGCCTT"


class P(object):
    def __init__(self, locations, bg_freq=1):
        self.locations, self.bg_freq = locations, bg_freq


@pytest.fixture(scope="module")
def quirky():
    """3.2 Mbp, 5 records (one empty, one shorter than any primer), AT-rich, with lower-case, N and
    IUPAC stretches that counting aliases to bases but locate must never match."""
    from swga_b200 import fasta, synth
    n = 3_200_000
    seq = synth.bases(21, 0, n, 0.30)
    seq[5000:5400] |= 0x20
    seq[700_000:700_050] = ord("N")
    seq[1_234_567] = ord("R")
    seq[2_000_000:2_000_003] = ord("n")
    rs = np.array([0, 1_000_003, 1_000_003, 1_000_009, 2_500_000, n], dtype=np.uint64)
    names = ["chrA", "empty", "six", "chrB", "chrC"]
    return fasta.Genome(seq, rs, names)


def records_of(g):
    s = g.seq.tobytes().decode("latin-1")
    return [(g.names[r], s[int(g.rec_starts[r]):int(g.rec_starts[r + 1])]) for r in range(g.n_records)]


def pick_primers(g, oracle, ks=(5, 6, 8, 10, 12), per_k=6, seed=3):
    rng = np.random.RandomState(seed)
    out = []
    s = g.seq
    for k in ks:
        got = 0
        while got < per_k:
            p = int(rng.randint(0, g.n_bases - k))
            w = s[p:p + k].tobytes().decode("latin-1")
            if all(c in "ACGT" for c in w):
                out.append(w)
                got += 1
    out += ["AATT", "ACGT", "AAAATTTT", "ATATATATAT", "GGGGGGGGGGGG"]      # palindromes and a rare one
    return out


def test_reference_locate_kats(oracle, tmp_path):
    from swga_b200 import locate
    fp = tmp_path / "test.fasta"
    fp.write_bytes(TEST_FASTA)
    loc = locate.binding_sites("AAGG", str(fp))
    assert len(loc["record1"]) == 4 and loc["record2"] == []                 # test_locate.py:12-15
    assert loc == oracle.binding_sites("AAGG", oracle.fasta_records_pyfaidx(TEST_FASTA))
    ends = locate.chromosome_ends(str(fp))
    assert ends == {"record1": [0, 15], "record2": [16, 23], "record3": [24, 31]}
    lin = locate.linearize_binding_sites([P(loc)], ends)
    assert len(lin) == 10
    assert locate.revcomp("GCAT") == "ATGC"
    with pytest.raises(KeyError):
        locate.binding_sites("AANG", str(fp))
    with pytest.raises(KeyError):
        locate.binding_sites("aagg", str(fp))                                # revcomp() of lower case
    empty = tmp_path / "empty.fa"
    empty.write_bytes(b"")
    with pytest.raises(ValueError):
        locate.binding_sites("AAGG", str(empty))


def test_palindrome_sites_listed_twice(oracle, tmp_path):
    from swga_b200 import locate
    fa = b">simple_genome\n" + b"\n".join([b"AAAATTTT" * 8] * 9) + b"\n"
    fp = tmp_path / "simple.fa"
    fp.write_bytes(fa)
    loc = locate.binding_sites("AAAATTTT", str(fp))
    assert loc == oracle.binding_sites("AAAATTTT", oracle.fasta_records_pyfaidx(fa))
    assert len(loc["simple_genome"]) == 144 and loc["simple_genome"][:2] == [0, 8]


def test_binding_sites_match_oracle_on_quirky_genome(quirky, oracle):
    from swga_b200 import locate
    gi = locate.GenomeIndex(quirky)
    recs = records_of(quirky)
    for primer in pick_primers(quirky, oracle):
        got = gi.binding_sites(primer)
        want = oracle.binding_sites(primer, recs)
        assert list(got.keys()) == [n for n, _ in recs]
        assert got == want, primer


def test_locate_agrees_with_counts_on_clean_genome(oracle):
    """on upper-case ACGT, sum over records of len(binding_sites(p)) == fg_freq(p) for
    non-palindromic primers -- kernels (b) and (c) cross-checked."""
    from swga_b200 import fasta, kmers, locate, synth
    n = 1_500_000
    g = fasta.Genome(synth.bases(31, 0, n, 0.5), np.array([0, 400_000, n], dtype=np.uint64), ["a", "b"])
    tabs = kmers.count_all(g, 5, 12)
    gi = locate.GenomeIndex(g)
    for primer in pick_primers(g, oracle, per_k=4, seed=9)[:-5]:
        k = len(primer)
        f, r = kmers.kmer_to_code(primer), kmers.kmer_to_code(locate.revcomp(primer))
        if f == r:
            continue
        sites = gi.binding_sites(primer)
        assert sum(len(v) for v in sites.values()) == int(tabs.table(k)[min(f, r)])


def _oracle_scores(oracle, locs_of, chr_ends, sets, bgs, max_fg, interactive):
    out = []
    for ids, bg in zip(sets, bgs):
        pl = [locs_of[i] for i in ids if i >= 0]
        out.append(oracle.score_set(pl, max_fg, bg, chr_ends, interactive=interactive))
    return out


def _check(res, i, want, bg):
    score, var, max_dist = want
    assert int(res.fg_max_dist[i]) == max_dist
    if score is False:
        assert not res.passed[i]
        return
    assert res.passed[i]
    assert res.fg_dist_mean[i] == var["fg_dist_mean"]                        # bit-exact (one rounding)
    assert res.fg_dist_gini[i] == var["fg_dist_gini"]                        # bit-exact
    assert res.fg_dist_std[i] == pytest.approx(var["fg_dist_std"], rel=1e-9) # north_star: <= 1e-6 relative
    assert res.score[i] == pytest.approx(score, rel=1e-12)


def test_batched_scoring_matches_oracle(quirky, oracle):
    from swga_b200 import locate, score
    gi = locate.GenomeIndex(quirky)
    recs = records_of(quirky)
    primers = pick_primers(quirky, oracle)
    locs_of = [oracle.binding_sites(p, recs) for p in primers]
    chr_ends = oracle.chromosome_ends(recs)
    assert gi.chromosome_ends() == chr_ends
    sl = score.SiteLists.from_index(gi, primers)
    # per-primer unique site counts agree with the oracle's linearisation
    for i, p in enumerate(primers):
        lin = set()
        for rec, ls in locs_of[i].items():
            lin.update(l + chr_ends[rec][0] for l in ls)
        assert sl.counts()[sl.list_of_primer[i]] == len(lin), p
    rng = np.random.RandomState(5)
    S, max_m = 60, 10
    sets = np.full((S, max_m), -1, dtype=np.int32)
    for s in range(S):
        m = rng.randint(1, max_m + 1)
        sets[s, :m] = rng.choice(len(primers), m, replace=False)
    sets[0, :] = -1
    sets[0, 0] = len(primers) - 1                                            # a set of one rare primer
    bgs = rng.uniform(100, 1e5, S)
    bgs[1] = float("inf")
    from swga_b200 import _lib
    lib, ctx = _lib.load(), _lib.context(0)
    try:
        for cap, sort_impl in ((4096, 0), (4096, 1), (0, 0), (256, 0)):      # sort kernel (routed / merge front end) / dense kernel / mixed
            _lib.check(lib.swga_ctx_set_score_sort_cap(ctx.h, cap))
            _lib.check(lib.swga_ctx_set_score_sort_impl(ctx.h, sort_impl))
            for max_fg, interactive in ((36000, False), (3000, False), (0, True)):
                res = score.score_sets(sl, sets, bgs, max_fg, interactive=interactive)
                want = _oracle_scores(oracle, locs_of, chr_ends, sets, bgs, max_fg, interactive)
                n_pass = 0
                for i in range(S):
                    _check(res, i, want[i], bgs[i])
                    n_pass += bool(res.passed[i])
                assert n_pass > 0
            if cap == 0:
                # interactive sets include gaps far beyond the in-kernel histogram: the gap-sort path ran
                res = score.score_sets(sl, sets, bgs, 0, interactive=True)
                assert ((res.status & 2) != 0).any()
    finally:
        _lib.check(lib.swga_ctx_set_score_sort_cap(ctx.h, 4096))
        _lib.check(lib.swga_ctx_set_score_sort_impl(ctx.h, 0))


def test_reference_score_kats(oracle):
    from swga_b200 import score
    assert score.read_set_finder_line("1,2,3,4 3500.01234") == ([1, 2, 3, 4], 3500.01234)
    s, ns = score.default_score_set("fg_dist_mean/bg_dist_mean", [1, 2, 3, 4], [1, 3, 5, 7], 2, 3.0)
    assert s == 2.0 / 3.0 and "__builtins__" not in ns                       # test_score.py:18-29
    with pytest.raises(NameError):
        score.default_score_set("fg_dist_max", [1], [1, 3, 5, 7], 2, 3.0)
    # SURVEY.md App. B3
    fa = b">simple_genome\n" + b"\n".join([b"AAAATTTT" * 8] * 9) + b"\n"
    recs = oracle.fasta_records_pyfaidx(fa)
    loc = oracle.binding_sites("AAAATTTT", recs)
    ends = oracle.chromosome_ends(recs)
    fun = functools.partial(score.default_score_set, expression=score.DEFAULT_EXPRESSION)
    sc, var, max_dist = score.score_set([P(loc, 0)], 36000, float("inf"), ends, fun)
    assert (sc, max_dist) == (0.0, 8)
    assert var["fg_dist_mean"] == 7.986111111111111 and var["fg_dist_gini"] == 0.001714975845410628
    assert var["fg_dist_std"] == pytest.approx(0.11785113019775792, rel=1e-12)
    assert var["set_size"] == 1 and var["fg_max_dist"] == 8 and set(var) == set(score.VARIABLES)
    assert score.score_set([P(loc, 0)], 7, 1.0, ends, fun) == (False, {}, 8)
    assert score.score_set([P(loc, 0)], 7, 1.0, ends, fun, interactive=True)[2] == 8
    assert score.calculate_bg_dist_mean([P(loc, 2), P(loc, 3)], 10) == 2.0
    assert score.calculate_bg_dist_mean([P(loc, 0)], 10) == float("inf")
    # gini's Python-2 floor: gaps [1,1,1] -> -0.125
    _, ns = score.default_score_set("fg_dist_gini", [1], [10, 11, 12, 13], 1, 1.0)
    assert ns["fg_dist_gini"] == -0.125


def test_set_index_sharding_needs_no_collective(quirky, oracle):
    from swga_b200 import locate, score
    gi = locate.GenomeIndex(quirky)
    primers = pick_primers(quirky, oracle, per_k=3, seed=11)
    sl = score.SiteLists.from_index(gi, primers)
    rng = np.random.RandomState(2)
    S = 37
    sets = np.stack([rng.choice(len(primers), 6, replace=False) for _ in range(S)]).astype(np.int32)
    bgs = rng.uniform(1e3, 1e4, S)
    whole = score.score_sets(sl, sets, bgs, 36000)
    for world in (2, 4, 8):
        parts = []
        for rank in range(world):
            lo, hi = score.shard_sets(S, rank, world)
            if hi > lo:
                parts.append(score.score_sets(sl, sets[lo:hi], bgs[lo:hi], 36000))
        for key in ("fg_max_dist", "n_sites", "fg_dist_mean", "fg_dist_gini", "score"):
            cat = np.concatenate([p[key] for p in parts])
            assert np.array_equal(cat, whole[key], equal_nan=True), (world, key)


def test_scoring_kernels_edge_shapes(oracle):
    """Constructed site lists that drive every branch of the two scoring kernels: merge-sort sizes around the
    16-element segments and the capacity, duplicate positions across lists, sets of up to 31 lists (lists no
    lane of the dense kernel owns), tile segments longer than the 16-byte chunk prefetch covers, and tiles with
    more distinct sites than the compaction list holds (per-lane walk)."""
    from swga_b200 import _lib, score
    rng = np.random.RandomState(11)
    L = 3_000_000
    bounds = [0, 999_999, 1_000_000, L - 1]
    lists = []
    for n in (1, 2, 14, 15, 16, 17, 31, 33, 255, 256, 257, 1000, 2047, 4060):      # sparse, odd sizes
        lists.append(np.sort(rng.choice(L, n, replace=False)))
    lists.append(np.arange(500_000, 540_000, 3))                                   # 2,730 sites per 8,192-position tile
    lists.append(np.arange(500_001, 600_000, 7))
    lists.append(np.arange(2_000_000, 2_001_000))                                  # every position of a stretch
    for n in (9000, 20000, 50000):                                                 # ordinary dense lists
        lists.append(np.sort(rng.choice(L, n, replace=False)))
    lists.append(lists[-1][::2].copy())                                            # shares every position with another list
    # the routed dense kernel picks its form by the fullest 256-position window of a tile: 6-7, 12-13, 15-16, 17-18 sites
    for step in (40, 20, 16, 15):
        lists.append(np.arange(1_200_000 + 100_000 * (step % 7), 1_200_000 + 100_000 * (step % 7) + 60_000, step))
    while len(lists) < 34:
        lists.append(np.sort(rng.choice(L, int(rng.randint(3000, 9000)), replace=False)))
    sl = score.SiteLists.from_host_lists(lists, bounds, L)
    P_ = len(lists)
    max_m = 31
    sets = []
    for i in range(14):
        sets.append([i])                                                           # one sparse list + bounds
    sets += [[0, 1, 2, 3], [4, 5, 6, 7, 8], [9, 10, 11, 12, 13], [13], [11, 13], [14], [15], [14, 15, 16], [16],
             [17, 18, 19, 20], [19, 20], list(range(14, 34)), list(range(3, 34)), list(range(0, 31)),
             [21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 17, 18, 19, 14], [22], [23], [24], [25], [22, 23, 24, 25],
             [21, 22], [0, 24]]
    for _ in range(25):
        m = int(rng.randint(1, max_m + 1))
        sets.append(list(rng.choice(P_, m, replace=False)))
    arr = np.full((len(sets), max_m), -1, dtype=np.int32)
    for i, s in enumerate(sets):
        arr[i, :len(s)] = s
    bgs = rng.uniform(10, 1e4, len(sets))
    lib, ctx = _lib.load(), _lib.context(0)
    try:
        # sort cap x dense-kernel form x sparse-kernel front end: (0, 1, .) = every set through the bitmap form of round 1,
        # (., ., 1) = the sparse kernel always merges the lists; otherwise the routed forms (crowded windows fall back)
        for cap, dense_impl, sort_impl in ((4096, 0, 0), (4096, 0, 1), (0, 0, 0), (0, 1, 0), (64, 0, 0), (16384, 0, 0), (16384, 0, 1)):
            _lib.check(lib.swga_ctx_set_score_sort_cap(ctx.h, cap))
            _lib.check(lib.swga_ctx_set_score_dense_impl(ctx.h, dense_impl))
            _lib.check(lib.swga_ctx_set_score_sort_impl(ctx.h, sort_impl))
            for max_fg in (36000, 2 ** 31 - 1):
                res = score.score_sets(sl, arr, bgs, max_fg)
                for i, s in enumerate(sets):
                    pos = sorted(set(int(x) for l in s for x in lists[l]) | set(bounds))
                    gaps = oracle.seq_diff(list(pos))
                    mx = max(gaps)
                    assert int(res.n_sites[i]) == len(pos) and int(res.fg_max_dist[i]) == mx, (cap, i)
                    assert bool(res.passed[i]) == (mx <= max_fg)
                    assert float(res.fg_dist_mean[i]) == oracle.mean(gaps)
                    assert float(res.fg_dist_std[i]) == pytest.approx(oracle.stdev(gaps), rel=1e-12)
                    if res.passed[i]:
                        assert float(res.fg_dist_gini[i]) == oracle.gini(gaps), (cap, i)
    finally:
        _lib.check(lib.swga_ctx_set_score_sort_cap(ctx.h, 4096))
        _lib.check(lib.swga_ctx_set_score_dense_impl(ctx.h, 0))
        _lib.check(lib.swga_ctx_set_score_sort_impl(ctx.h, 0))


def test_config3_config4_full_size(oracle):
    """BASELINE configs 3 and 4 at their real size: the synthetic 23 Mbp AT-rich foreground in 14 records is counted
    (bit-exact vs the oracle), the C4 primers (its 200 most frequent 8..12-mers, ~25,000 sites each) are located and
    a sample of the C4 candidate sets (6-10 primers, ~196,000 unique sites: the dense kernel) is scored; sites and
    statistics are compared with the oracle's restatement of locate.py / stats.py."""
    import torch
    from swga_b200 import fasta, kmers, locate, score, synth, workloads
    seed, n, gc, n_rec = synth.CONFIGS["C3_fg"]
    seq = synth.bases(seed, 0, n, gc)
    rs = synth.equal_records(n, n_rec)
    g = fasta.Genome(seq, rs, ["chr%d" % (i + 1) for i in range(n_rec)])
    tabs = kmers.count_all(g, 5, 12)
    want = oracle.count_dense_allk(seq, rs, 5, 12, mode_b=int(g.dsk_mode_b))
    assert np.array_equal(np.concatenate([tabs.table(k) for k in range(5, 13)]), want)
    prim = workloads.top_primers(tabs, 5, 12, 200)
    seqs = [kmers.codes_to_kmers([c], k)[0] for k, c, _ in prim]
    gi = locate.GenomeIndex(g)
    sl = score.SiteLists.from_index(gi, seqs)
    S = 12
    sets_, m = workloads.c4_sets(S)
    bgs = np.full(S, 1234.5)
    res = score.score_sets(sl, sets_, bgs, 36000)
    recs = records_of(g)
    ends = oracle.chromosome_ends(recs)
    locs = {}
    for s in range(S):
        ids = [int(x) for x in sets_[s] if x >= 0]
        for i in ids:
            if i not in locs:
                locs[i] = oracle.binding_sites(seqs[i], recs)
                assert gi.binding_sites(seqs[i]) == locs[i]                  # kernel (c) at full size
        lin = oracle.linearize_binding_sites([locs[i] for i in ids], ends)
        gaps = oracle.seq_diff(lin)
        assert int(res.n_sites[s]) == len(lin) and int(res.fg_max_dist[s]) == max(gaps)
        assert float(res.fg_dist_mean[s]) == oracle.mean(gaps)
        assert float(res.fg_dist_std[s]) == pytest.approx(oracle.stdev(gaps), rel=1e-12)
        assert bool(res.passed[s]) == (max(gaps) <= 36000)
        if res.passed[s]:
            assert float(res.fg_dist_gini[s]) == oracle.gini(gaps)
            assert float(res.score[s]) == pytest.approx(oracle.mean(gaps) * oracle.gini(gaps) / 1234.5, rel=1e-12)
