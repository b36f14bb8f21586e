"""swga_build_edges (the heterodimer graph of `swga find_sets`, all pairs in one launch) against the reference's
known answers (swga/tests/test_graph.py:28-51, test_find_sets_cmd.py:18-30) and the oracle's restatement of
graph.build_edges + kmers.max_sequential_nt on random primer lists."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def P(_id, seq, **kw):
    from swga_b200 import pipeline
    p = pipeline.Primer(seq, kw.get("fg_freq", 0), kw.get("bg_freq", 0), kw.get("ratio", 0.0))
    p._id = _id
    return p


def oracle_edges(primers, max_binding, oracle):
    return [[a._id, b._id] for a, b in itertools.combinations(primers, 2)
            if a.seq not in b.seq and b.seq not in a.seq and oracle.max_sequential_nt(a.seq, b.seq) <= max_binding]


def test_reference_graph_kats_on_device(tmp_path):
    from swga_b200 import pipeline
    ref = P(0, "ATGCTC")
    for seq in ("CAGCAT", "GAGGTA", "ATCGAG"):
        assert pipeline.build_edges([ref, P(1, seq)], 2) == []
    assert pipeline.build_edges([ref, P(4, "TTCCAC")], 2) == [[0, 4]]
    assert pipeline.build_edges([ref, P(5, "ATGC")], 2) == []
    assert pipeline.build_edges([ref], 2) == [] and pipeline.build_edges([], 2) == []
    primers = [P(None, "ATGC", fg_freq=1, bg_freq=2, ratio=1.0), P(None, "GGCC", fg_freq=1, bg_freq=3, ratio=0.5),
               P(None, "CCTA", fg_freq=2, bg_freq=0, ratio=float("inf"))]
    out = tmp_path / "graph"
    pipeline.build_graph(primers, 3, str(out))
    assert out.read_text() == "p sp 3 3\nn 1 1\nn 2 2\nn 3 3\ne 1 2\ne 1 3\ne 2 3\n"
    with pytest.raises(KeyError):
        pipeline.build_edges([ref, P(1, "ACGN")], 2)


@pytest.mark.parametrize("n,lo,hi,seed", [(2, 5, 12, 0), (33, 5, 12, 1), (97, 4, 9, 2), (200, 8, 12, 3), (260, 1, 40, 4)])
def test_random_primer_lists_equal_oracle(n, lo, hi, seed, oracle):
    from swga_b200 import pipeline
    rng = np.random.RandomState(seed)
    seqs = set()
    while len(seqs) < n:
        ln = int(rng.randint(lo, hi + 1))
        s = "".join("ACGT"[i] for i in rng.randint(0, 4, ln))
        if rng.rand() < 0.2 and seqs:                              # a substring / extension of an existing primer
            t = sorted(seqs)[int(rng.randint(0, len(seqs)))]
            s = t[: max(1, len(t) - 2)] if rng.rand() < 0.5 else t + s[:3]
        seqs.add(s)
    primers = [P(i + 1, s) for i, s in enumerate(sorted(seqs, key=lambda s: (rng.rand(), s)))]
    for max_binding in (0, 2, 3, 5):
        assert pipeline.build_edges(primers, max_binding) == oracle_edges(primers, max_binding, oracle), (n, max_binding)
    assert pipeline.build_edges(primers, 3) == pipeline.build_edges_host(primers, 3)
