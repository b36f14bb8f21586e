"""The reference's own known-answer tests for the host remainder of the path, run against swga_b200's host
code (no GPU): heterodimer graph (swga/tests/test_graph.py:28-51), DIMACS output (test_find_sets_cmd.py:18-30),
the set_finder protocol (test_find_sets_cmd.py:33-61, test_score.py:4-7) and the golden set_finder stream."""
import json
import os

import pytest

from swga_b200 import pipeline, score, sets

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def P(_id, seq, **kw):
    p = pipeline.Primer(seq, kw.get("fg_freq", 0), kw.get("bg_freq", 0), kw.get("ratio", 0.0))
    p._id = _id
    return p


def test_reference_graph_kats():
    """The host statement of build_edges; tests/test_graph_gpu.py runs the same answers through the kernel."""
    ref = P(0, "ATGCTC")
    for seq in ("CAGCAT", "GAGGTA", "ATCGAG"):                           # heterodimers: no edge
        assert pipeline.build_edges_host([ref, P(1, seq)], 2) == []
    assert pipeline.build_edges_host([ref, P(4, "TTCCAC")], 2) == [[0, 4]]    # valid pair
    assert pipeline.build_edges_host([ref, P(5, "ATGC")], 2) == []            # subsequence


def test_reference_build_graph_kat(tmp_path):
    primers = [P(None, "ATGC", fg_freq=1, bg_freq=2, ratio=1.0), P(None, "GGCC", fg_freq=1, bg_freq=3, ratio=0.5),
               P(None, "CCTA", fg_freq=2, bg_freq=0, ratio=float("inf"))]
    out = tmp_path / "graph"
    pipeline.build_graph(primers, 3, str(out), edges_fn=pipeline.build_edges_host)
    # ids by ratio descending (primers.py:255-268); vertex weight = bg_freq or 1 (graph.py:45-52)
    assert out.read_text() == "p sp 3 3\nn 1 1\nn 2 2\nn 3 3\ne 1 2\ne 1 3\ne 2 3\n"


def test_reference_set_finder_kat(tmp_path):
    try:
        sets.set_finder()
    except IOError:
        pytest.skip("set_finder binary not built (needs the reference sources at build time)")
    graph = ("p sp 6 7\nn 1 1\nn 2 2\nn 3 1\nn 4 4\nn 5 5\nn 6 6\n"
             "e 1 2\ne 1 3\ne 2 3\ne 2 4\ne 4 5\ne 4 6\ne 5 6\n")
    fp = tmp_path / "testgraph"
    fp.write_text(graph)
    gen = sets.find(min_bg_bind_dist=2, min_size=3, max_size=3, bg_length=10, graph_fp=str(fp))
    output = list(gen)
    assert len(output) == 1
    ids, bg_dist_mean = score.read_set_finder_line(output[0])
    assert ids == [1, 2, 3] and bg_dist_mean == 2.5
    with open(os.path.join(GOLDEN, "set_finder_golden.json")) as f:
        gold = json.load(f)
    assert gold["graph"] == graph


def test_read_set_finder_line_kat():
    assert score.read_set_finder_line("1,2,3 1.5") == ([1, 2, 3], 1.5)   # swga/tests/test_score.py:4-7
    with pytest.raises(ValueError):
        score.read_set_finder_line("not a line")


def test_native_set_line_parser_equals_read_set_finder_line():
    """swga_parse_set_lines against score.read_set_finder_line (swga/score.py:16-24) on well-formed lines (the
    %f format of set_finder.c:442) and on every malformation Python rejects with ValueError."""
    import numpy as np
    rng = np.random.RandomState(5)
    lines = []
    for _ in range(2000):
        m = int(rng.randint(1, 8))
        ids = rng.randint(1, 5000, m)
        lines.append(",".join(str(int(i)) for i in ids) + " %f" % rng.uniform(0, 1e7))
    bad = ["", "1,2", "1,2 ", " 1.0", "1,,2 3.0", "1,2 3.0 4", "a,b 1.0", "1,2 x", "1;2 1.0", "1,2  1.0", "3, 4 1.0"]
    lines[100:100] = bad
    block = ("\n".join(lines) + "\n7,8 1.25").encode()                    # the last line is incomplete
    ids, w, ok, consumed = score.parse_set_lines(block, 7)
    assert consumed == len(block) - len("7,8 1.25") and len(ids) == len(lines)
    for i, line in enumerate(lines):
        try:
            want_ids, want_w = score.read_set_finder_line(line)
        except ValueError:
            assert not ok[i], line
            continue
        assert ok[i], line
        assert [int(x) for x in ids[i] if x >= 0] == want_ids and w[i] == want_w
    ids2, w2, ok2, c2 = score.parse_set_lines(b"1,2,3,4 1.0\n", 3)          # more ids than max_m: flagged
    assert not ok2[0] and c2 == 12


def test_default_score_set_with_duplicated_locations(oracle):
    """The reference's default_score_set takes ANY list of locations: stats.seq_diff keeps the zero gaps of
    repeated positions (stats.py:22-33).  linearize_binding_sites never produces such a list, so the case is
    evaluated on the host (swga_b200/score.py) -- against the restated stats.py."""
    for locs in ([1, 3, 3, 5, 7, 7, 7, 20], [5, 5, 9, 100, 100, 101], [0, 0, 1, 2], [10, 4, 4, 30, 10]):
        for expr in ("fg_dist_mean * fg_dist_gini / bg_dist_mean", "fg_dist_std + set_size + fg_max_dist"):
            want_score, want_ns = oracle.default_score_set(expr, [1, 2, 3], list(locs), 13, 2.5)
            got_score, got_ns = score.default_score_set(expr, [1, 2, 3], list(locs), 13, 2.5)
            assert got_ns["fg_dist_mean"] == want_ns["fg_dist_mean"]
            assert got_ns["fg_dist_gini"] == want_ns["fg_dist_gini"]
            assert got_ns["fg_dist_std"] == pytest.approx(want_ns["fg_dist_std"], rel=1e-12)
            assert got_score == pytest.approx(want_score, rel=1e-12)
    with pytest.raises(ZeroDivisionError):
        score.default_score_set("fg_dist_mean", [1], [4, 4, 4], 0, 1.0)          # all gaps 0: fair_area == 0


def test_melting_temperature_restatement():
    """swga_b200/melting.py (parity unpinned: the reference's `melt` package is not vendored).  Sanity of the
    nearest-neighbour model: GC raises Tm, length raises Tm, the value is strand-symmetric for a palindrome,
    swga's default window (15..45 C, filter.yaml) selects 8..12-mers of ordinary composition."""
    from swga_b200 import melting
    assert melting.temp("GGGCCCGG") > melting.temp("ACGTACGT") > melting.temp("AAAATTTT")
    assert melting.temp("ACGTACGTACGT") > melting.temp("ACGTACGT")
    assert melting.temp("aaaatttt") == melting.temp("AAAATTTT")
    assert 15 < melting.temp("AAAATTTT") < 45 and 15 < melting.temp("GATTACAGATTA") < 45
    assert abs(melting.temp("ACGTACGT") - melting.temp("ACGTACGT", uncorrected=True)) < 5
    with pytest.raises(ValueError):
        melting.temp("ACGN")


def test_filter_skip_filtering_activates_every_listed_primer(tmp_path, capsys):
    """Primers.activate(self.max_primers) (swga/primers.py:275-290, commands/filter.py:42): max_primers is only the
    number below which the 'Fewer than' note is printed -- every primer of the list is marked active."""
    from swga_b200 import commands, workspace
    db = str(tmp_path / "swga.db")
    w = workspace.Workspace(db)
    w.create_tables()
    w.metadata = dict(version=workspace.VERSION, fg_file="fg.fa", bg_file="bg.fa", ex_file=None, fg_length=100000,
                      bg_length=2000000)
    seqs = ["AAGT", "AACC", "AGGA", "CCTG", "GATT", "TTGC"]
    w.add_primers([dict(seq=s, fg_freq=5, bg_freq=1, ratio=5.0) for s in seqs], add_revcomp=False)
    w.close()
    lst = tmp_path / "primers.txt"
    lst.write_text("".join(s + "\n" for s in seqs[:5]))
    act = commands.filter(db_name=db, input=str(lst), skip_filtering=True, max_primers=3)
    assert sorted(act) == sorted(seqs[:5])
    assert "Fewer than" not in capsys.readouterr().err
    act = commands.filter(db_name=db, input=str(lst), skip_filtering=True, max_primers=10)
    assert sorted(act) == sorted(seqs[:5])
    assert "Fewer than 10 primers were selected (5 passed all the filters)" in capsys.readouterr().err
    with workspace.connection(db) as w:
        assert sorted(p.seq for p in w.primers('"active" = 1')) == sorted(seqs[:5])
