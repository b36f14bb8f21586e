"""The `swga init | count | filter | find_sets | score` command surface over the workspace database
(swga/commands/*.py), with the data-parallel work done by the B200 kernels:

  init       commands/init.py:86-190      genome lengths -> _metadata, default binding-rate parameters
  count      commands/count.py:20-135     kernels (a)+(b) for both genomes, the merge on the device
                                          (k_filter_primers), rows into `primer`
  filter     commands/filter.py:17-42     SQL-free restatement of the Primers filter chain; locate (c)
                                          and per-primer Gini (d) for the primers that reach that step
  find_sets  commands/find_sets.py:22-103 graph + host set_finder + batched scoring (d), rows into `set`
  score      commands/score.py:16-77      one user-supplied set, interactive (never rejected)

Option names and defaults are the reference's (`commands/specfiles/*.yaml`).  Prompts (`click.confirm`)
become the `force` argument: without it a command that would delete rows raises instead of asking.
Melting temperatures come from the third-party `melt` package in the reference (parity unpinned,
SURVEY.md 8c): `filter` applies the Tm range only to primers whose `tm` column is already filled
(e.g. by the reference) or when a `tm_fn` is supplied.

    python -m swga_b200.commands count --min_size 5 --max_size 12 ...      (inside a workspace directory)
"""
import argparse
import os
import sys

from . import fasta, kmers, locate, pipeline, score, workspace

MIN_FG_RATE = 0.00001            # commands/init.py:34-38
MAX_BG_RATE = 0.0000067
GRAPH_FP = "compatibility_graph.dimacs"
OUTPUT_DIR = ".swga_tmp"         # commands/count.py:17


class CommandError(RuntimeError):
    pass


def message(msg):
    sys.stderr.write(str(msg) + "\n")


def _open(db_name):
    ws = workspace.connection(db_name)
    return ws, ws.metadata


def default_parameters(fg_length, bg_length):
    """The values `swga init` writes into parameters.cfg (commands/init.py:150-175, specfiles/*.yaml)."""
    return {
        "count": dict(min_size=5, max_size=12, min_fg_bind=int(MIN_FG_RATE * fg_length),
                      max_bg_bind=int(MAX_BG_RATE * bg_length), max_dimer_bp=3, exclude_threshold=1),
        "filter": dict(max_primers=200, min_fg_bind=int(MIN_FG_RATE * fg_length), max_bg_bind=int(MAX_BG_RATE * bg_length),
                       min_tm=15, max_tm=45, max_gini=0.6),
        "find_sets": dict(min_size=2, max_size=7, max_dimer_bp=3, min_bg_bind_dist=30000, max_fg_bind_dist=36000,
                          max_sets=-1, workers=1, score_expression=score.DEFAULT_EXPRESSION),
        "score": dict(score_expression=score.DEFAULT_EXPRESSION),
    }


# ---------------------------------------------------------------------------------------------- init

def fasta_len_quick(fasta_fp):
    """commands/init.py:196-209: the reference's genome "length" is `grep -v '>' file | wc -c`, i.e. the bytes
    of every line that contains no '>' PLUS one newline per such line.  These numbers, not
    the base counts, feed min_fg_bind / max_bg_bind, set_finder's --bg_len and calculate_bg_dist_mean."""
    import mmap
    import re
    size = os.path.getsize(fasta_fp)
    if size == 0:
        return 0
    with open(fasta_fp, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            total = size + (0 if mm[size - 1:size] == b"\n" else 1)      # grep terminates the last line
            for m in re.finditer(rb"[^\n]*>[^\n]*(?:\n|$)", mm):
                total -= (m.end() - m.start()) + (0 if mm[m.end() - 1:m.end()] == b"\n" else 1)
        finally:
            mm.close()
    return total


def init(fg_genome_fp, bg_genome_fp, exclude_fp=None, db_name=workspace.DEFAULT_DB_FNAME, force=False):
    """commands/init.py: record the genomes and their lengths; returns the default parameters."""
    if os.path.isfile(db_name):
        if not force:
            raise CommandError("Existing workspace %s: pass force=True to overwrite" % db_name)
        os.remove(db_name)
    fg = fasta.read_fasta(fg_genome_fp)                           # also validates the file (fasta_stats)
    fg_length, bg_length = fasta_len_quick(fg_genome_fp), fasta_len_quick(bg_genome_fp)
    with workspace.Workspace(db_name) as ws:
        ws.create_tables()
        ws.metadata = dict(version=workspace.VERSION, fg_file=os.path.abspath(fg_genome_fp),
                           bg_file=os.path.abspath(bg_genome_fp),
                           ex_file=os.path.abspath(exclude_fp) if exclude_fp else None,
                           fg_length=fg_length, bg_length=bg_length)
    message("Foreground filepath: {}\n  Length:  {} bp\n  Records: {}".format(fg_genome_fp, fg_length, fg.n_records))
    message("Background filepath: {}\n  Length:  {} bp".format(bg_genome_fp, bg_length))
    return default_parameters(fg_length, bg_length)


# ---------------------------------------------------------------------------------------------- count

def count(db_name=workspace.DEFAULT_DB_FNAME, min_size=5, max_size=12, min_fg_bind=None, max_bg_bind=None,
          max_dimer_bp=3, exclude_threshold=1, input=None, force=False):
    """Count.run (commands/count.py:22-114).  With `input` (a file of primer sequences, one per line) only
    those k-mers are counted and added without their reverse complements (count.py:29-68)."""
    ws, meta = _open(db_name)
    with ws:
        d = default_parameters(meta["fg_length"], meta["bg_length"])["count"]
        min_fg_bind = d["min_fg_bind"] if min_fg_bind is None else min_fg_bind
        max_bg_bind = d["max_bg_bind"] if max_bg_bind is None else max_bg_bind
        if input is not None:
            return _count_specific(ws, meta, input)
        if ws.count_primers() > 0:
            if not force:
                raise CommandError("Remove all previously-found primers and re-count? pass force=True")
            ws.reset_primers()
        rows = pipeline.count_primers(meta["fg_file"], meta["bg_file"], meta["ex_file"], min_size=min_size,
                                      max_size=max_size, min_fg_bind=min_fg_bind, max_bg_bind=max_bg_bind,
                                      max_dimer_bp=max_dimer_bp, exclude_threshold=exclude_threshold, add_revcomp=False)
        # Primers.add(add_revcomp=True) inserts seq and revcomp(seq); a palindrome would violate the primary
        # key there (count.py:112 + primers.py:77-82) -- here it is inserted once
        seen, out = set(), []
        for r in rows:
            for s in (r["seq"], locate.revcomp(r["seq"])):
                if s not in seen:
                    seen.add(s)
                    out.append(dict(r, seq=s))
        n = ws.add_primers(out, add_revcomp=False)
        message("{} primers written to {}".format(n, db_name))
        return n


def _count_specific(ws, meta, input_fp):
    with open(input_fp) as f:
        wanted = kmers.parse_kmer_file(f)
    have = {p.seq for p in ws.primers_by_seqs(wanted)}
    for s in wanted:
        if s in have:
            message("Primer {} already exists in db, skipping...".format(s))
    wanted = [s for s in wanted if s not in have]
    rows = []
    os.makedirs(OUTPUT_DIR, exist_ok=True)
    by_len = {}
    for s in wanted:
        by_len.setdefault(len(s), []).append(s)
    for k, mers in by_len.items():
        fg = kmers.count_kmers(k, meta["fg_file"], OUTPUT_DIR, 1)
        bg = kmers.count_kmers(k, meta["bg_file"], OUTPUT_DIR, 1)
        for mer in mers:
            # primer_dict(mer, fg, bg, 0, INF, INF) (count.py:51-56,117-135).  The dicts hold DSK-canonical k-mers
            # only, so -- as in the reference -- a listed primer whose OTHER strand is the canonical one is
            # reported missing and skipped, and only the canonical string is looked up in bg.
            if mer not in fg:
                message("{} does not exist in foreground genome, skipping...".format(mer))
                continue
            f_, b_ = fg[mer], bg.get(mer, 0)
            rows.append(dict(seq=mer, fg_freq=f_, bg_freq=b_, ratio=(f_ / float(b_)) if b_ > 0 else float("inf")))
    return ws.add_primers(rows, add_revcomp=False)


# ---------------------------------------------------------------------------------------------- filter

def filter(db_name=workspace.DEFAULT_DB_FNAME, max_primers=200, min_fg_bind=None, max_bg_bind=None, min_tm=15,
           max_tm=45, max_gini=0.6, input=None, skip_filtering=False, tm_fn=None, gindex=None):
    """Filter.run (commands/filter.py:17-42) followed by Primers.activate (primers.py:300-311)."""
    ws, meta = _open(db_name)
    with ws:
        d = default_parameters(meta["fg_length"], meta["bg_length"])["filter"]
        min_fg_bind = d["min_fg_bind"] if min_fg_bind is None else min_fg_bind
        max_bg_bind = d["max_bg_bind"] if max_bg_bind is None else max_bg_bind
        if input is not None:
            with open(input) as f:
                primers = ws.primers_by_seqs(kmers.parse_kmer_file(f))
        else:
            skip_filtering = False
            primers = ws.primers()
        n0 = len(primers)
        ws.deactivate_all()
        if not skip_filtering:
            ps = [p for p in primers if p.fg_freq >= min_fg_bind]
            message("{}/{} primers bind the foreground genome >= {} times".format(len(ps), n0, min_fg_bind))
            n1 = len(ps)
            ps = [p for p in ps if p.bg_freq <= max_bg_bind]
            message("{}/{} primers bind the background genome <= {} times".format(len(ps), n1, max_bg_bind))
            if tm_fn is not None:
                for p in ps:
                    if p.tm is None:
                        p.tm = tm_fn(p.seq)
                ws.update_primers(ps, fields=("tm",))
            if tm_fn is not None or all(p.tm is not None for p in ps):
                n2 = len(ps)
                ps = [p for p in ps if min_tm <= p.tm <= max_tm]
                message("{}/{} primers have a melting temp between {} and {} C".format(len(ps), n2, min_tm, max_tm))
            else:
                message("melting temperatures are not available (third-party `melt`): Tm range not applied")
            ps = pipeline.limit_to(ps, max_primers) if ps else []
            gindex = gindex or locate.genome_index(meta["fg_file"])
            pipeline.update_locations(ps, gindex)
            pipeline.update_gini(ps, gindex)
            ws.update_primers(ps, fields=("_locations", "gini"))
            n3 = len(ps)
            ps = [p for p in ps if p.gini <= max_gini]
            message("{}/{} primers have a Gini coefficient <= {}".format(len(ps), n3, max_gini))
            primers = ps
        # Primers.activate(max_primers): mark, at most max_primers
        primers = primers[:max_primers]
        for p in primers:
            p.active = True
        ws.update_primers(primers, fields=("active",))
        message("Marked {} primers as active.".format(len(primers)))
        return [p.seq for p in primers]


# ---------------------------------------------------------------------------------------------- find_sets

def find_sets(db_name=workspace.DEFAULT_DB_FNAME, min_size=2, max_size=7, max_dimer_bp=3, min_bg_bind_dist=30000,
              max_fg_bind_dist=36000, max_sets=-1, workers=1, score_expression=score.DEFAULT_EXPRESSION,
              force=False, gindex=None, graph_fp=GRAPH_FP):
    """FindSets.run + process_lines (commands/find_sets.py:22-103)."""
    ws, meta = _open(db_name)
    with ws:
        if ws.count_sets() > 0:
            if not force:
                raise CommandError("Remove all previously-found sets? pass force=True")
            ws.reset_sets()
        active = ws.primers('"active" = 1')
        if not active:
            raise CommandError("No active primers found. Run `swga filter` or `swga activate` first.")
        ws.reset_ids()
        ordered = pipeline.assign_ids(active)                       # graph.build_graph -> Primers.assign_ids
        ws.update_primers(ordered, fields=("_id",))
        gindex = gindex or locate.genome_index(meta["fg_file"])
        pipeline.update_locations(ordered, gindex)
        found = pipeline.find_sets(ordered, gindex, meta["bg_length"], min_bg_bind_dist, max_fg_bind_dist,
                                   min_size=min_size, max_size=max_size, max_dimer_bp=max_dimer_bp, max_sets=max_sets,
                                   workers=workers, score_expression=score_expression, graph_fp=graph_fp)
        for row in found:
            ws.add_set(row["_id"], row["primers"], row["score"], row["scoring_fn"],
                       **{v: row[v] for v in workspace.SET_VARIABLES})
        return found


# ---------------------------------------------------------------------------------------------- score

def score_set(db_name=workspace.DEFAULT_DB_FNAME, input=None, primers=None, score_expression=score.DEFAULT_EXPRESSION,
              force=False, gindex=None):
    """Score.run (commands/score.py:16-77): statistics of one user-defined set (never rejected); with
    `force` it is stored under the next negative id."""
    ws, meta = _open(db_name)
    with ws:
        if primers is None:
            with open(input) as f:
                primers = kmers.parse_kmer_file(f)
        ps = ws.primers_by_seqs(primers)
        if not ps:
            raise CommandError("No primers specified exist in database, aborting.")
        gindex = gindex or locate.genome_index(meta["fg_file"])
        pipeline.update_locations(ps, gindex)
        bg_dist_mean = score.calculate_bg_dist_mean(ps, meta["bg_length"])
        sl = score.SiteLists.from_index(gindex, [p.seq for p in ps])
        import numpy as np
        res = score.score_sets(sl, np.arange(len(ps), dtype=np.int32)[None, :], np.asarray([bg_dist_mean]), 0,
                               interactive=True)
        variables = score._variables(res, 0, len(ps), bg_dist_mean)
        default_expr = score_expression.strip() == score.DEFAULT_EXPRESSION
        sc = float(res.score[0]) if default_expr else score.evaluate_expression(score_expression, variables)
        message("Set statistics:\n - " + "\n - ".join("{}: {}".format(k, v) for k, v in
                                                      dict(variables, score=sc, scoring_fn=score_expression).items()))
        set_id = None
        if force:
            set_id = (ws.min_set_id() or 0) - 1
            if set_id >= 0:
                set_id = -1
            _, created = ws.add_set(set_id, [p.seq for p in ps], sc, score_expression, **variables)
            message("Set {} added successfully.".format(set_id) if created else "That primer set already exists.")
        return dict(variables, score=sc, scoring_fn=score_expression, _id=set_id)


# ---------------------------------------------------------------------------------------------- CLI

def main(argv=None):
    ap = argparse.ArgumentParser(prog="swga_b200", description=__doc__.split("\n")[0])
    ap.add_argument("--db", default=workspace.DEFAULT_DB_FNAME)
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("init")
    p.add_argument("-f", "--fg_genome_fp", required=True)
    p.add_argument("-b", "--bg_genome_fp", required=True)
    p.add_argument("-e", "--exclude_fp")
    p.add_argument("--force", action="store_true")
    p = sub.add_parser("count")
    for name, typ in (("min_size", int), ("max_size", int), ("min_fg_bind", int), ("max_bg_bind", int),
                      ("max_dimer_bp", int), ("exclude_threshold", int)):
        p.add_argument("--" + name, type=typ)
    p.add_argument("--input")
    p.add_argument("--force", action="store_true")
    p = sub.add_parser("filter")
    for name, typ in (("max_primers", int), ("min_fg_bind", int), ("max_bg_bind", int), ("min_tm", float),
                      ("max_tm", float), ("max_gini", float)):
        p.add_argument("--" + name, type=typ)
    p.add_argument("--input")
    p.add_argument("--skip_filtering", action="store_true")
    p = sub.add_parser("find_sets")
    for name, typ in (("min_size", int), ("max_size", int), ("max_dimer_bp", int), ("min_bg_bind_dist", int),
                      ("max_fg_bind_dist", int), ("max_sets", int), ("workers", int), ("score_expression", str)):
        p.add_argument("--" + name, type=typ)
    p.add_argument("--force", action="store_true")
    p = sub.add_parser("score")
    p.add_argument("--input", required=True)
    p.add_argument("--score_expression")
    p.add_argument("--force", action="store_true")
    args, _unknown = ap.parse_known_args(argv)                  # unknown options are tolerated (_command.py:48)
    kw = {k: v for k, v in vars(args).items() if k not in ("cmd", "db") and v is not None and v is not False}
    if args.cmd == "init":
        init(args.fg_genome_fp, args.bg_genome_fp, args.exclude_fp, db_name=args.db, force=args.force)
    else:
        fn = {"count": count, "filter": filter, "find_sets": find_sets, "score": score_set}[args.cmd]
        fn(db_name=args.db, **kw)
    return 0


if __name__ == "__main__":
    sys.exit(main())
