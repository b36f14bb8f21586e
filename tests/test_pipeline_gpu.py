"""count -> filter -> locate -> gini -> find_sets (host set_finder) on a small synthetic workspace,
every stage compared with the oracle restatement of the reference (BASELINE config 5 in miniature)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def workspace(tmp_path_factory):
    from swga_b200 import synth
    d = tmp_path_factory.mktemp("ws")
    fg_seq = synth.bases(61, 0, 260_000, 0.30)
    bg_seq = synth.bases(62, 0, 2_400_000, 0.42)
    ex_seq = synth.bases(63, 0, 20_000, 0.30)
    fg_rs = np.array([0, 100_000, 260_000], dtype=np.uint64)
    fg, bg, ex = str(d / "fg.fa"), str(d / "bg.fa"), str(d / "ex.fa")
    synth.write_fasta(fg, fg_seq, fg_rs, ["chr1", "chr2"])
    synth.write_fasta(bg, bg_seq, np.array([0, 1_200_000, 2_400_000], dtype=np.uint64))
    synth.write_fasta(ex, ex_seq, np.array([0, 20_000], dtype=np.uint64))
    return d, fg, bg, ex


def test_count_primers_matches_reference_merge(workspace, oracle):
    from swga_b200 import pipeline
    d, fg, bg, ex = workspace
    kw = dict(min_size=5, max_size=9, min_fg_bind=3, max_bg_bind=40, max_dimer_bp=3, exclude_threshold=2)
    rows = pipeline.count_primers(fg, bg, ex, **kw)
    got = {r["seq"]: (r["fg_freq"], r["bg_freq"], r["ratio"]) for r in rows}
    assert len(got) == len(rows)                                         # seq is the primary key
    want = {}
    fgd, bgd, exd = open(fg, "rb").read(), open(bg, "rb").read(), open(ex, "rb").read()
    for k in range(5, 10):
        f, b = oracle.count_kmers_dict(fgd, k), oracle.count_kmers_dict(bgd, k)
        e = oracle.count_kmers_dict(exd, k, 2)
        for seq in f:
            if seq in e:
                continue
            p = oracle.primer_dict(seq, f, b, 3, 40, 3)
            if p:
                want[seq] = (p["fg_freq"], p["bg_freq"], p["ratio"])
                want[oracle.revcomp(seq)] = want[seq]                    # Primers.add(add_revcomp=True)
    assert got == want
    # -1 disables the background check; no exclusion file
    rows2 = pipeline.count_primers(fg, bg, None, min_size=6, max_size=6, min_fg_bind=50, max_bg_bind=-1, max_dimer_bp=2,
                                   add_revcomp=False)
    f, b = oracle.count_kmers_dict(fgd, 6), oracle.count_kmers_dict(bgd, 6)
    want2 = {s for s in f if oracle.primer_dict(s, f, b, 50, -1, 2)}
    assert {r["seq"] for r in rows2} == want2


def test_filter_and_find_sets_match_oracle(workspace, oracle):
    from swga_b200 import fasta, locate, pipeline, score
    d, fg, bg, ex = workspace
    rows = pipeline.count_primers(fg, bg, None, min_size=7, max_size=8, min_fg_bind=20, max_bg_bind=300, max_dimer_bp=3)
    primers = [pipeline.Primer(r["seq"], r["fg_freq"], r["bg_freq"], r["ratio"]) for r in rows]
    gi = locate.GenomeIndex(fg)
    active = pipeline.filter_primers(primers, gi, min_fg_bind=20, max_bg_bind=300, max_primers=40, max_gini=0.8)
    assert 5 < len(active) <= 40
    recs = oracle.fasta_records_pyfaidx(open(fg, "rb").read())
    ends = oracle.chromosome_ends(recs)
    for p in active:                                                      # workspace.py:150-158
        want_loc = oracle.binding_sites(p.seq, recs)
        assert p.locations == want_loc
        lin = oracle.linearize_binding_sites([want_loc], ends)
        assert p.gini == oracle.gini(oracle.seq_diff(lin))
    bg_len = fasta.read_fasta(bg).n_bases
    graph = str(d / "compatibility_graph.dimacs")
    found = pipeline.find_sets(active, gi, bg_len, min_bg_bind_dist=3000, max_fg_bind_dist=15000, min_size=2,
                               max_size=4, max_dimer_bp=3, max_sets=25, graph_fp=graph)
    assert 0 < len(found) <= 25 and [r["_id"] for r in found] == list(range(1, len(found) + 1))
    # replay the same set_finder stream through the oracle's process_lines loop
    from swga_b200 import sets
    by_id = {p._id: p for p in active}
    passed = []
    gen = sets.find(min_bg_bind_dist=3000, min_size=2, max_size=4, bg_length=bg_len, graph_fp=graph)
    for line in gen:
        ids, bgm = oracle.read_set_finder_line(line)
        sc, var, md = oracle.score_set([by_id[i].locations for i in ids], 15000, bgm, ends)
        if sc is not False:
            passed.append((ids, sc, var, md))
            if len(passed) >= 25:
                break
    gen.close()
    assert len(passed) == len(found)
    for row, (ids, sc, var, md) in zip(found, passed):
        assert row["primer_ids"] == ids and row["fg_max_dist"] == md and row["set_size"] == len(ids)
        assert row["fg_dist_mean"] == var["fg_dist_mean"] and row["fg_dist_gini"] == var["fg_dist_gini"]
        assert row["fg_dist_std"] == pytest.approx(var["fg_dist_std"], rel=1e-9)
        assert row["score"] == pytest.approx(sc, rel=1e-12)
    # a custom expression is evaluated with the reference's namespace
    found2 = pipeline.find_sets(active, gi, bg_len, 3000, 15000, 2, 4, max_sets=3, graph_fp=graph,
                                score_expression="fg_max_dist / set_size")
    for row in found2:
        assert row["score"] == row["fg_max_dist"] // row["set_size"]       # Python-2 integer division


def test_config5_full_size(oracle, tmp_path):
    """BASELINE config 5 at its real size: synthetic 1.3 Mbp foreground (Wolbachia-sized) vs 140 Mbp background
    (Drosophila-sized, 6 records -> DSK mode B), full count -> filter -> locate -> find_sets with the host
    set_finder, the reference's default parameters.  Counting is compared bit for bit with the oracle's dense
    tables (the 140 Mbp background takes the partitioned kernel), the rest with the oracle's replay."""
    import torch
    from swga_b200 import fasta, kmers, locate, pipeline, sets, synth
    genomes = {}
    for name in ("C5_fg", "C5_bg"):
        seed, n, gc, n_rec = synth.CONFIGS[name]
        seq = synth.bases(seed, 0, n, gc)
        rs = synth.equal_records(n, n_rec)
        g = fasta.Genome(seq, rs, ["%s_r%d" % (name, i + 1) for i in range(n_rec)])
        genomes[name] = g
        d = torch.from_numpy(seq).cuda()
        _, canon = kmers.count_all_device(d, rs, mode_b=g.dsk_mode_b)
        torch.cuda.synchronize()
        want = oracle.count_dense_allk(seq, rs, 5, 12, mode_b=int(g.dsk_mode_b))
        assert np.array_equal(canon.cpu().numpy().view(np.uint32), want), name
        del d, canon, want
    fg, bg = genomes["C5_fg"], genomes["C5_bg"]
    assert bg.dsk_mode_b and fg.dsk_mode_b                       # one record of 1.3 Mbp > 1,000,000 as well
    min_fg_bind, max_bg_bind = int(1e-5 * fg.n_bases), int(6.7e-6 * bg.n_bases)
    rows = pipeline.count_primers(fg, bg, None, min_size=5, max_size=12, min_fg_bind=min_fg_bind,
                                  max_bg_bind=max_bg_bind, max_dimer_bp=3)
    assert len(rows) > 1000
    primers = [pipeline.Primer(r["seq"], r["fg_freq"], r["bg_freq"], r["ratio"]) for r in rows]
    gi = locate.GenomeIndex(fg)
    active = pipeline.filter_primers(primers, gi, min_fg_bind=min_fg_bind, max_bg_bind=max_bg_bind, max_primers=200,
                                     max_gini=0.6)
    assert 20 < len(active) <= 200
    recs = [(fg.names[r], fg.seq[int(fg.rec_starts[r]):int(fg.rec_starts[r + 1])].tobytes().decode("latin-1"))
            for r in range(fg.n_records)]
    ends = oracle.chromosome_ends(recs)
    for p in active[:40]:
        loc = oracle.binding_sites(p.seq, recs)
        assert p.locations == loc and sum(len(v) for v in loc.values()) == p.fg_freq
        assert p.gini == oracle.gini(oracle.seq_diff(oracle.linearize_binding_sites([loc], ends)))
    graph = str(tmp_path / "c5.dimacs")
    found = pipeline.find_sets(active, gi, int(bg.n_bases), min_bg_bind_dist=30000, max_fg_bind_dist=36000, min_size=2,
                               max_size=7, max_dimer_bp=3, max_sets=10, graph_fp=graph)
    by_id = {p._id: p for p in active}
    assert 0 < len(found) <= 10
    # every stored set is what the oracle computes for the same primers, and it passes the filter
    for row in found:
        sc, var, md = oracle.score_set([by_id[i].locations for i in row["primer_ids"]], 36000, row["bg_dist_mean"], ends)
        assert sc is not False and row["fg_max_dist"] == md
        assert row["fg_dist_mean"] == var["fg_dist_mean"] and row["fg_dist_gini"] == var["fg_dist_gini"]
        assert row["score"] == pytest.approx(sc, rel=1e-12)
    # ... and they are the FIRST passing sets of the set_finder stream: replay its first lines through the oracle
    gen = sets.find(min_bg_bind_dist=30000, min_size=2, max_size=7, bg_length=int(bg.n_bases), graph_fp=graph)
    head, oracle_pass = [], []
    for line in gen:
        ids, bgm = oracle.read_set_finder_line(line)
        head.append(tuple(ids))
        sc, var, md = oracle.score_set([by_id[i].locations for i in ids], 36000, bgm, ends)
        if sc is not False:
            oracle_pass.append(tuple(ids))
        if len(head) >= 1500 or len(oracle_pass) >= 10:
            break
    gen.close()
    in_head = set(head)
    assert [tuple(r["primer_ids"]) for r in found if tuple(r["primer_ids"]) in in_head][:len(oracle_pass)] == oracle_pass
