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
