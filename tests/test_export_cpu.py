"""swga_b200.export against the oracle's loop-for-loop restatement of swga/export.py and stats.lorenz
(host code only: no GPU)."""
import io
import os

import numpy as np
import pytest


def _random_case(rng, n_rec, n_primers):
    names = ["r%d" % i for i in range(n_rec)]
    lens = [int(rng.randint(1, 400)) for _ in names]
    ends, so_far = {}, 0
    for nm, ln in zip(names, lens):
        ends[nm] = [so_far, so_far + ln - 1]
        so_far += ln
    locs = []
    for _ in range(n_primers):
        d = {}
        for nm, ln in zip(names, lens):
            fwd = sorted(set(rng.randint(0, ln, int(rng.randint(0, 12))).tolist()))
            rev = fwd if rng.rand() < 0.2 else sorted(set(rng.randint(0, ln, int(rng.randint(0, 12))).tolist()))
            d[nm] = list(fwd) + list(rev)                       # forward list + reverse-complement list (palindrome: equal)
        locs.append(d)
    return ends, locs


def test_bedgraph_windows_equal_reference_loops(oracle):
    from swga_b200 import export
    rng = np.random.RandomState(3)
    for it in range(60):
        ends, locs = _random_case(rng, int(rng.randint(1, 4)), int(rng.randint(1, 5)))
        window = int(rng.choice([5, 7, 10, 50, 300, 1000]))
        step = None if rng.rand() < 0.5 else int(rng.randint(1, 80))
        got = export.bedgraph_rows(locs, ends, window, step, warn=lambda m: None)
        assert got == oracle.bedgraph_hits_per_record(locs, ends, window, step), (it, window, step)
    with pytest.raises(ValueError):
        export.bedgraph_rows(locs, ends, 4, None, warn=lambda m: None)       # step = 4 / 5 = 0: xrange refuses


def test_lorenz_equals_reference_loop(oracle):
    from swga_b200 import export
    rng = np.random.RandomState(4)
    for it in range(40):
        d = rng.randint(0, 10 ** int(rng.randint(1, 9)), int(rng.randint(1, 200))).tolist()
        if sum(d) == 0:
            continue
        assert export.lorenz(list(d)) == oracle.lorenz(list(d))
    ends, locs = _random_case(rng, 3, 4)
    sites = oracle.linearize_binding_sites(locs, ends)
    assert export.set_lorenz(locs, ends) == oracle.lorenz(oracle.seq_diff(sites))
    for x in (1.0, 0.5, 1 / 3.0, 2 / 3.0, 1e-5, 123456789012345.0, 1e16, 0.1 + 0.2):
        assert export.py2_str(x) == oracle.py2_float_str(x)
    assert [export.py2_str(v) for v in (1.0, 1 / 3.0, 1e16)] == ["1.0", "0.333333333333", "1e+16"]


def test_table_export_and_bed_files(tmp_path, oracle):
    from swga_b200 import export, workspace
    db = str(tmp_path / "p.db")
    ws = workspace.Workspace(db)
    ws.create_tables()
    ws.metadata = dict(version=workspace.VERSION, fg_file="/x/fg.genome.fa", bg_file="bg.fa", ex_file=None,
                       fg_length=1000, bg_length=5000)
    ws.add_primers([dict(seq="ACGTA", fg_freq=9, bg_freq=3, ratio=3.0), dict(seq="GGATC", fg_freq=4, bg_freq=0, ratio=float("inf")),
                    dict(seq="TTTAC", fg_freq=7, bg_freq=7, ratio=1.0)], add_revcomp=False)
    ps = ws.primers()
    for i, p in enumerate(ps):
        p._id, p.gini, p.locations = i + 1, 0.25 * (i + 1), {"a": [1 + i, 30], "b": [5]}
    ws.update_primers(ps, fields=("_id", "gini", "_locations"))
    ws.add_set(1, ["ACGTA", "GGATC"], 0.5, "f", set_size=2, bg_dist_mean=10.0, fg_dist_std=1.5, fg_dist_gini=0.1,
               fg_max_dist=40, fg_dist_mean=12.5)
    ws.add_set(2, ["TTTAC"], 0.25, "f", set_size=1, bg_dist_mean=4.0, fg_dist_std=2.5, fg_dist_gini=0.2, fg_max_dist=50,
               fg_dist_mean=3.5)
    out = io.StringIO()
    export.export_rows(export.PRIMER_FIELDS, export.select_rows(ws, "primers", order_by="bg_freq", limit=2), out)
    assert out.getvalue() == "seq\tfg_freq\tbg_freq\tratio\tgini\ttm\nGGATC\t4\t0\tinf\t0.5\t\nACGTA\t9\t3\t3.0\t0.25\t\n"
    out = io.StringIO()
    export.export_rows(export.SET_FIELDS, export.select_rows(ws, "sets", order_by="score", descending=True), out, header=False)
    assert out.getvalue() == ("1\t0.5\t2\t10.0\t40\t12.5\t1.5\t0.1\tf\tACGTA,GGATC\n"
                              "2\t0.25\t1\t4.0\t50\t3.5\t2.5\t0.2\tf\tTTTAC\n")
    assert [r["_id"] for r in export.select_rows(ws, "sets", ids=[2], order_by="score", limit=1)] == [2]
    with pytest.raises(ValueError):
        export.select_rows(ws, "sets", order_by="nope")
    # BED files: <out>/<fg name without its last extension>_export/set_<id>/{whole_set,<seq>}.bed
    folder = export.write_bedfiles(1, [(p.seq, p.locations) for p in ws.primers_by_seqs(["ACGTA", "GGATC"])],
                                   "/x/fg.genome.fa", str(tmp_path))
    assert folder == os.path.join(str(tmp_path), "fg.genome_export", "set_1")
    assert open(os.path.join(folder, "ACGTA.bed")).read() == "track name=ACGTA\na 1 6\na 30 35\nb 5 10\n"
    whole = open(os.path.join(folder, "whole_set.bed")).read().splitlines()
    assert whole[0] == "track name=Set_1" and sorted(whole[1:]) == sorted(["a 1 6", "a 30 35", "b 5 10", "a 2 7", "a 30 35", "b 5 10"])
    ends = {"a": [0, 99], "b": [100, 149]}
    fp = export.write_bedgraph(1, [p.locations for p in ws.primers_by_seqs(["ACGTA", "GGATC"])], "/x/fg.genome.fa", ends,
                               str(tmp_path), opts_str="color=1", window_size=20, step_size=10)
    lines = open(fp).read().splitlines()
    want = oracle.bedgraph_hits_per_record([{"a": [1, 30], "b": [5]}, {"a": [2, 30], "b": [5]}], ends, 20, 10)
    assert lines[0] == "track type=bedGraph color=1" and lines[1:] == ["%s %d %d %d" % (r, m, m, h) for r, m, h in want]
    out = io.StringIO()
    export.export_lorenz([(1, [{"a": [1, 30], "b": [5]}, {"a": [2, 30], "b": [5]}])], out, ends)
    curve = oracle.lorenz(oracle.seq_diff(oracle.linearize_binding_sites([{"a": [1, 30], "b": [5]}, {"a": [2, 30], "b": [5]}], ends)))
    assert out.getvalue() == "SetID\tCDF\n1\t" + ",".join(oracle.py2_float_str(v) for v in curve) + "\n"
    ws.close()


def test_export_and_summary_commands_through_the_cli(tmp_path, capsys):
    """`python -m swga_b200.commands --db X export ... | summary`: option wiring of the argparse front end on a
    workspace filled through the workspace API (neither command touches the GPU)."""
    from swga_b200 import commands, workspace
    fg = tmp_path / "fg.fa"
    fg.write_bytes(b">a\n" + b"ACGTTGCA" * 20 + b"\n>b\n" + b"GGATCCTA" * 10 + b"\n")
    db = str(tmp_path / "p.db")
    ws = workspace.Workspace(db)
    ws.create_tables()
    ws.metadata = dict(version=workspace.VERSION, fg_file=str(fg), bg_file="bg.fa", ex_file=None, fg_length=240, bg_length=5000)
    ws.add_primers([dict(seq="ACGTT", fg_freq=20, bg_freq=3, ratio=20 / 3.0), dict(seq="GGATC", fg_freq=10, bg_freq=0, ratio=float("inf"))],
                   add_revcomp=False)
    ps = ws.primers()
    ps[0].locations, ps[1].locations = {"a": [0, 8, 16], "b": []}, {"a": [], "b": [0, 8]}
    for i, p in enumerate(ps):
        p._id, p.active, p.tm = i + 1, True, 30.0 + i
    ws.update_primers(ps)
    ws.add_set(1, ["ACGTT", "GGATC"], 0.5, "f", set_size=2, bg_dist_mean=10.0, fg_dist_std=1.5, fg_dist_gini=0.1, fg_max_dist=40,
               fg_dist_mean=12.5)
    ws.close()
    out = tmp_path / "primers.tsv"
    assert commands.main(["--db", db, "export", "primers", "--output", str(out), "--order_by", "fg_freq", "--descending",
                          "--limit", "1", "--unknown_option", "7"]) == 0
    assert out.read_text() == "seq\tfg_freq\tbg_freq\tratio\tgini\ttm\nACGTT\t20\t3\t%r\t\t30.0\n" % (20 / 3.0)
    assert commands.main(["--db", db, "export", "sets", "--no_header"]) == 0
    assert capsys.readouterr().out == "1\t0.5\t2\t10.0\t40\t12.5\t1.5\t0.1\tf\tACGTT,GGATC\n"
    assert commands.main(["--db", db, "export", "lorenz", "--ids", "1"]) == 0
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == "SetID\tCDF" and lines[1].startswith("1\t") and lines[1].endswith(",1.0")
    assert commands.main(["--db", db, "export", "bedgraph", "--ids", "1", "--output_folder", str(tmp_path / "x"),
                          "--window_size", "50", "--step_size", "25"]) == 0
    bg = (tmp_path / "x" / "fg_export" / "set_1" / "set_1.bedgraph").read_text().splitlines()
    assert bg[0] == "track type=bedGraph None" and bg[1] == "a 25 25 3"
    assert commands.main(["--db", db, "summary"]) == 0
    text = capsys.readouterr().out
    assert "There are 2 primers in the database." in text and "The best scoring set is #1, with 2 primers" in text
    assert "ranges between 30.00C and 31.00C" in text
