"""Seeded random small cases through all three stages against the oracle: byte palettes with N / n / lower case /
IUPAC / control bytes, ragged and empty records, every k range -- the shapes a fixed fixture never has."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PALETTES = [b"ACGT", b"ACGTN", b"ACGTNnacgt", b"ACGTRYKMSWN", b"AT", b"ACGTacgtN\x00\xff*-"]


def random_genome(rng, max_len=6000):
    from swga_b200 import fasta
    pal = np.frombuffer(PALETTES[rng.randint(len(PALETTES))], dtype=np.uint8)
    n = int(rng.choice([0, 1, 3, 11, 12, 13, 31, 32, 33, 64, 500, rng.randint(0, max_len)]))
    seq = pal[rng.randint(0, len(pal), n)].astype(np.uint8)
    if n and rng.rand() < 0.5:                                   # a long single-base run and a repeat
        a = rng.randint(0, n)
        seq[a:a + rng.randint(1, 200)] = pal[0]
    n_cut = int(rng.randint(0, 6))
    cuts = sorted(set([0, n] + [int(rng.randint(0, n + 1)) for _ in range(n_cut)]))
    if rng.rand() < 0.3 and n:
        cuts = sorted(cuts + [cuts[len(cuts) // 2]])               # an empty record
    rs = np.array(cuts, dtype=np.uint64)
    return fasta.Genome(seq, rs, ["r%d" % i for i in range(len(rs) - 1)])


def test_counting_fuzz(oracle):
    from swga_b200 import kmers
    rng = np.random.RandomState(1234)
    for it in range(60):
        g = random_genome(rng)
        if g.n_records == 0:
            continue
        kmin = int(rng.randint(2, 13))
        kmax = int(rng.randint(kmin, 14))
        tabs = kmers.count_all(g, kmin, kmax)
        for k in range(kmin, kmax + 1):
            want = oracle.count_dense(g.seq, g.rec_starts, k)
            assert np.array_equal(tabs.table(k), want), (it, k, g.n_bases, g.rec_starts.tolist())


def test_locate_and_score_fuzz(oracle):
    from swga_b200 import fasta, locate, score
    rng = np.random.RandomState(4321)
    for it in range(12):
        pal = np.frombuffer(PALETTES[rng.randint(len(PALETTES))], dtype=np.uint8)
        n = int(rng.randint(200, 40000))
        seq = pal[rng.randint(0, len(pal), n)].astype(np.uint8)
        cuts = sorted(set([0, n] + [int(rng.randint(1, n)) for _ in range(int(rng.randint(0, 4)))]))
        g = fasta.Genome(seq, np.array(cuts, dtype=np.uint64), ["c%d" % i for i in range(len(cuts) - 1)])
        s = seq.tobytes().decode("latin-1")
        recs = [(g.names[r], s[cuts[r]:cuts[r + 1]]) for r in range(len(cuts) - 1)]
        gi = locate.GenomeIndex(g)
        primers = []
        for _ in range(10):
            k = int(rng.randint(4, 11))
            p = "".join("ACGT"[x] for x in rng.randint(0, 4, k))
            primers.append(p)
        for p in primers:
            assert gi.binding_sites(p) == oracle.binding_sites(p, recs), (it, p)
        ends = oracle.chromosome_ends(recs)
        assert gi.chromosome_ends() == ends
        locs = [oracle.binding_sites(p, recs) for p in primers]
        sl = score.SiteLists.from_index(gi, primers)
        S = 20
        sets_ = np.full((S, 6), -1, dtype=np.int32)
        for i in range(S):
            m = int(rng.randint(1, 7))
            sets_[i, :m] = rng.choice(len(primers), m, replace=False)
        bgs = rng.uniform(1, 1e4, S)
        for max_fg in (50, 5000, 2 ** 31 - 1):
            res = score.score_sets(sl, sets_, bgs, max_fg)
            for i in range(S):
                ids = [int(x) for x in sets_[i] if x >= 0]
                try:
                    sc, var, md = oracle.score_set([locs[j] for j in ids], max_fg, float(bgs[i]), ends)
                except ZeroDivisionError:                        # one gap only: stats.stdev divides by n - 1 = 0
                    assert int(res.n_sites[i]) == 2              # (swga/stats.py:17-19; the kernel reports NaN)
                    continue
                assert int(res.fg_max_dist[i]) == md, (it, i)
                assert bool(res.passed[i]) == (sc is not False)
                if sc is not False:
                    assert float(res.fg_dist_mean[i]) == var["fg_dist_mean"]
                    g_, w_ = float(res.fg_dist_gini[i]), var["fg_dist_gini"]
                    assert g_ == w_ or (g_ != g_ and w_ != w_), (it, i, g_, w_)
                    if len(oracle.seq_diff(oracle.linearize_binding_sites([locs[j] for j in ids], ends))) > 1:
                        assert float(res.fg_dist_std[i]) == pytest.approx(var["fg_dist_std"], rel=1e-12, nan_ok=True)
