"""Locate + score legs of __graft_entry__.smoke(): checked against the oracle passed in by the caller
(this module never imports oracle/ itself)."""
import numpy as np


def run(genome, tables, O):
    from . import locate, score
    gi = locate.GenomeIndex(genome)
    s = genome.seq.tobytes().decode("latin-1")
    recs = [(genome.names[r], s[int(genome.rec_starts[r]):int(genome.rec_starts[r + 1])])
            for r in range(genome.n_records)]
    primers = ["AAAAT", "ATATAT", "TTTAAAT", "AATTAATT", s[5000:5010], s[200000:200012]]
    locs = []
    for p in primers:
        got = gi.binding_sites(p)
        want = O.binding_sites(p, recs)
        assert got == want, "locate mismatch for %s" % p
        locs.append(want)
    ends = O.chromosome_ends(recs)
    sl = score.SiteLists.from_index(gi, primers)
    sets = np.array([[0, 1, 2, -1], [3, 4, 5, 0], [5, -1, -1, -1]], dtype=np.int32)
    bgs = np.array([1000.0, 2500.0, float("inf")])
    res = score.score_sets(sl, sets, bgs, 36000)
    for i in range(len(sets)):
        want = O.score_set([locs[j] for j in sets[i] if j >= 0], 36000, bgs[i], ends)
        assert int(res.fg_max_dist[i]) == want[2]
        if want[0] is False:
            assert not res.passed[i]
            continue
        assert res.fg_dist_mean[i] == want[1]["fg_dist_mean"] and res.fg_dist_gini[i] == want[1]["fg_dist_gini"]
        assert abs(res.fg_dist_std[i] - want[1]["fg_dist_std"]) <= 1e-9 * want[1]["fg_dist_std"]
    print("smoke ok: locate (%d primers) and set scoring (%d sets) match the oracle" % (len(primers), len(sets)))
