"""`init -> count -> filter -> find_sets -> score -> export / summary / activate` through swga_b200.commands on a workspace database,
compared with the oracle restatement of the reference's commands (swga/commands/*.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ws_dir(tmp_path_factory):
    from swga_b200 import synth
    d = tmp_path_factory.mktemp("wsdb")
    fg_seq = synth.bases(71, 0, 300_000, 0.32)
    bg_seq = synth.bases(72, 0, 3_000_000, 0.44)
    synth.write_fasta(str(d / "fg.fa"), fg_seq, np.array([0, 120_000, 300_000], dtype=np.uint64), ["chrA", "chrB"])
    synth.write_fasta(str(d / "bg.fa"), bg_seq, np.array([0, 3_000_000], dtype=np.uint64))
    return d


def test_commands_end_to_end(ws_dir, oracle):
    from swga_b200 import commands, workspace
    db = str(ws_dir / "primers.db")
    fg, bg = str(ws_dir / "fg.fa"), str(ws_dir / "bg.fa")
    params = commands.init(fg, bg, db_name=db)
    # lengths as the reference takes them (init.py:196-209): grep -v '>' | wc -c = bases + one newline per line
    import subprocess
    fg_len, bg_len = (int(subprocess.check_output("grep -v '>' %s | wc -c" % f, shell=True)) for f in (fg, bg))
    assert fg_len > 300_000 and bg_len > 3_000_000
    with workspace.connection(db) as ws:
        assert ws.metadata["fg_length"] == fg_len and ws.metadata["bg_length"] == bg_len
    assert params["count"]["min_fg_bind"] == int(1e-5 * fg_len) and params["count"]["max_bg_bind"] == int(6.7e-6 * bg_len)
    with pytest.raises(commands.CommandError):
        commands.init(fg, bg, db_name=db)                                # existing workspace, no force

    # ---- count: every row the reference's merge would insert (count.py:84-135), both strands
    n = commands.count(db_name=db, min_size=6, max_size=8, min_fg_bind=15, max_bg_bind=200, max_dimer_bp=3)
    fgd, bgd = open(fg, "rb").read(), open(bg, "rb").read()
    want = {}
    for k in range(6, 9):
        f, b = oracle.count_kmers_dict(fgd, k), oracle.count_kmers_dict(bgd, k)
        for seq in f:
            p = oracle.primer_dict(seq, f, b, 15, 200, 3)
            if p:
                want[seq] = (p["fg_freq"], p["bg_freq"], p["ratio"])
                want[oracle.revcomp(seq)] = want[seq]
    with workspace.connection(db) as ws:
        got = {p.seq: (p.fg_freq, p.bg_freq, p.ratio) for p in ws.primers()}
    assert n == len(got) and got == want
    with pytest.raises(commands.CommandError):
        commands.count(db_name=db, min_size=6, max_size=6)               # would delete rows: needs force

    # ---- filter: the active set, locations and Gini of the primers that reached that step
    from swga_b200 import melting
    active = commands.filter(db_name=db, max_primers=30, min_fg_bind=15, max_bg_bind=200, max_gini=0.85,
                             min_tm=-20, max_tm=60)
    assert 4 < len(active) <= 30
    recs = oracle.fasta_records_pyfaidx(fgd)
    ends = oracle.chromosome_ends(recs)
    with workspace.connection(db) as ws:
        act = ws.primers('"active" = 1')
        assert sorted(p.seq for p in act) == sorted(active)
        for p in act:
            assert p.tm == melting.temp(p.seq) and -20 <= p.tm <= 60     # filter_tm_range always runs (primers.py:151-165)
            loc = oracle.binding_sites(p.seq, recs)
            assert p.locations == loc
            assert p.gini == oracle.gini(oracle.seq_diff(oracle.linearize_binding_sites([loc], ends)))
    # the reference's chain on the same rows: fg >= , bg <=, the 30 least bg-binding, gini <= 0.85
    rows = sorted(((s,) + v for s, v in want.items() if v[0] >= 15 and v[1] <= 200), key=lambda r: r[2])
    assert len(active) <= len(rows)

    # ---- find_sets: rows of `set` equal the oracle's replay of the same set_finder stream
    graph = str(ws_dir / "g.dimacs")
    found = commands.find_sets(db_name=db, min_size=2, max_size=3, max_dimer_bp=3, min_bg_bind_dist=2000,
                               max_fg_bind_dist=20000, max_sets=12, graph_fp=graph)
    assert 0 < len(found) <= 12
    with workspace.connection(db) as ws:
        by_id = {p._id: p for p in ws.primers('"active" = 1')}
        sets_db = ws.sets()
    assert sorted(by_id) == list(range(1, len(by_id) + 1))               # Primers.assign_ids
    assert [r["_id"] for r in sets_db] == list(range(1, len(found) + 1))
    from swga_b200 import sets
    gen = sets.find(min_bg_bind_dist=2000, min_size=2, max_size=3, bg_length=bg_len, graph_fp=graph)
    passed = []
    for line in gen:
        ids, bgm = oracle.read_set_finder_line(line)
        sc, var, md = oracle.score_set([by_id[i].locations for i in ids], 20000, bgm, ends)
        if sc is not False:
            passed.append((ids, sc, var, md))
            if len(passed) >= 12:
                break
    gen.close()
    assert len(passed) == len(sets_db)
    for row, (ids, sc, var, md) in zip(sets_db, passed):
        assert sorted(row["primers"]) == sorted(by_id[i].seq for i in ids)
        assert row["fg_max_dist"] == md and row["set_size"] == len(ids)
        assert row["fg_dist_mean"] == var["fg_dist_mean"] and row["fg_dist_gini"] == var["fg_dist_gini"]
        assert row["score"] == pytest.approx(sc, rel=1e-12)
    with pytest.raises(commands.CommandError):
        commands.find_sets(db_name=db, graph_fp=graph)                   # sets exist: needs force

    # ---- score: a user-defined set, never rejected, stored under a negative id
    seqs = [by_id[1].seq, by_id[2].seq, by_id[3].seq]
    out = commands.score_set(db_name=db, primers=seqs, force=True)
    bgm = oracle.calculate_bg_dist_mean([by_id[i].bg_freq for i in (1, 2, 3)], bg_len)
    sc, var, md = oracle.score_set([by_id[i].locations for i in (1, 2, 3)], 0, bgm, ends, interactive=True)
    assert out["_id"] == -1 and out["fg_max_dist"] == md and out["score"] == pytest.approx(sc, rel=1e-12)
    out2 = commands.score_set(db_name=db, primers=[by_id[4].seq, by_id[5].seq], force=True)
    assert out2["_id"] == -2

    # ---- export (sets table, Lorenz curve, bedgraph, BED files), summary, activate
    import io
    from swga_b200 import export
    buf = io.StringIO()
    commands.export("sets", db_name=db, output=buf, order_by="score", limit=3)
    lines = buf.getvalue().splitlines()
    assert lines[0].split("\t") == list(export.SET_FIELDS) and len(lines) == 1 + min(3, len(sets_db) + 2)
    first = sets_db[0]
    locs1 = [next(p for p in by_id.values() if p.seq == s).locations for s in first["primers"]]
    buf = io.StringIO()
    commands.export("lorenz", db_name=db, output=buf, ids=[first["_id"]])
    curve = oracle.lorenz(oracle.seq_diff(oracle.linearize_binding_sites(locs1, ends)))
    assert buf.getvalue() == "SetID\tCDF\n%d\t%s\n" % (first["_id"], ",".join(oracle.py2_float_str(v) for v in curve))
    outdir = str(ws_dir / "exports")
    commands.export("bedgraph", db_name=db, ids=[first["_id"]], output_folder=outdir, window_size=5000.0, opts_str="x=1")
    bg_lines = open(os.path.join(outdir, "fg_export", "set_%d" % first["_id"], "set_%d.bedgraph" % first["_id"])).read().splitlines()
    want_rows = oracle.bedgraph_hits_per_record(locs1, ends, 5000)
    assert bg_lines[0] == "track type=bedGraph x=1" and bg_lines[1:] == ["%s %d %d %d" % (r[0], r[1], r[1], r[2]) for r in want_rows]
    assert sum(r[2] for r in want_rows) > 0
    commands.export("bedfile", db_name=db, ids=[first["_id"]], output_folder=outdir)
    whole = open(os.path.join(outdir, "fg_export", "set_%d" % first["_id"], "whole_set.bed")).read().splitlines()
    assert len(whole) == 1 + sum(len(v) for l in locs1 for v in l.values())
    info = commands.summary(db_name=db, out=io.StringIO())
    assert info["nsets"] == len(sets_db) + 2 and info["nactive"] == len(by_id) and info["best_set"]["_id"] is not None
    with workspace.connection(db) as ws:
        idle = [p.seq for p in ws.primers('"active" = 0 AND "_locations" IS NULL')[:2]]
    lst = ws_dir / "activate.txt"
    lst.write_text("\n".join(idle + ["ACGTACGTAC"]) + "\n")                 # the last one is not in the database
    assert sorted(commands.activate(str(lst), db_name=db)) == sorted(idle)
    with workspace.connection(db) as ws:
        for p in ws.primers_by_seqs(idle):
            assert p.active and p.locations == oracle.binding_sites(p.seq, recs)


def test_count_input_adds_listed_primers_like_the_reference(ws_dir, oracle, tmp_path):
    """`swga count --input FILE` (count.py:29-68): listed k-mers are looked up in the DSK-canonical dicts only
    (a primer whose other strand is canonical is skipped), existing ones are skipped, no reverse complements."""
    from swga_b200 import commands, workspace
    db = str(tmp_path / "p2.db")
    fg, bg = str(ws_dir / "fg.fa"), str(ws_dir / "bg.fa")
    commands.init(fg, bg, db_name=db)
    fgd, bgd = open(fg, "rb").read(), open(bg, "rb").read()
    f7, b7 = oracle.count_kmers_dict(fgd, 7), oracle.count_kmers_dict(bgd, 7)
    canon = sorted(f7)[:3]
    other = next(oracle.revcomp(s) for s in sorted(f7)[3:] if oracle.revcomp(s) not in f7)   # non-canonical strand
    lst = tmp_path / "primers.txt"
    lst.write_text("\n".join(canon + [other]) + "\n")
    cwd = os.getcwd()
    os.chdir(str(tmp_path))                                              # the DSK cache goes to ./.swga_tmp
    try:
        assert commands.count(db_name=db, input=str(lst)) == 3
        assert commands.count(db_name=db, input=str(lst)) == 0           # all known (or skipped) now
    finally:
        os.chdir(cwd)
    with workspace.connection(db) as ws:
        got = {p.seq: (p.fg_freq, p.bg_freq) for p in ws.primers()}
    assert got == {s: (f7[s], b7.get(s, 0)) for s in canon}
