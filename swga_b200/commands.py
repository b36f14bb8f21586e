"""The `swga init | count | filter | find_sets | score | export | activate | summary` command surface over the workspace database
(swga/commands/*.py), with the data-parallel work done by the B200 kernels:

  init       commands/init.py:86-190      genome lengths -> _metadata, default binding-rate parameters
  count      commands/count.py:20-135     kernels (a)+(b) for both genomes, the merge on the device
                                          (k_filter_primers), rows into `primer`
  filter     commands/filter.py:17-42     SQL-free restatement of the Primers filter chain; locate (c)
                                          and per-primer Gini (d) for the primers that reach that step
  find_sets  commands/find_sets.py:22-103 graph + host set_finder + batched scoring (d), rows into `set`
  score      commands/score.py:16-77      one user-supplied set, interactive (never rejected)
  export     commands/export.py:36-165    primers / sets as TSV, Lorenz curve, BED files, bedgraph (swga_b200/export.py)
  activate   commands/activate.py:6-18    locations for a list of primers (locate kernel), mark them active
  summary    commands/summary.py:55-124   counts and the best set (plain text, no terminal colours)

Option names and defaults are the reference's (`commands/specfiles/*.yaml`).  Prompts (`click.confirm`)
become the `force` argument: without it a command that would delete rows raises instead of asking.
Melting temperatures come from the third-party `melt` package in the reference (source not vendored: parity
unpinned); `swga_b200/melting.py` restates the published nearest-neighbour method it documents and
`filter` / `activate` fill in missing `tm` values with it (or with a caller's `tm_fn`) before the Tm range is
applied, as Primers.filter_tm_range does (primers.py:151-165).

    python -m swga_b200.commands count --min_size 5 --max_size 12 ...      (inside a workspace directory)
"""
import argparse
import os
import sys

from . import export as _export
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
            # filter_tm_range (primers.py:151-165): missing melting temperatures are computed first
            # (update_melt_temps, primers.py:241-253), then the range is ALWAYS applied
            if tm_fn is None:
                from . import melting
                tm_fn = melting.temp
            todo = [p for p in ps if p.tm is None]
            if todo:
                message("Finding melting temps for {} primers...".format(len(todo)))
                for p in todo:
                    p.tm = tm_fn(p.seq)
                ws.update_primers(todo, fields=("tm",))
            n2 = len(ps)
            ps = [p for p in ps if min_tm <= p.tm <= max_tm]
            message("{}/{} primers have a melting temp between {} and {} C".format(len(ps), n2, min_tm, max_tm))
            ps = pipeline.limit_to(ps, max_primers) if ps else []
            gindex = gindex or locate.genome_index(meta["fg_file"])
            pipeline.update_locations(ps, gindex)
            pipeline.update_gini(ps, gindex)
            ws.update_primers(ps, fields=("_locations", "gini"))
            n3 = len(ps)
            ps = [p for p in ps if p.gini <= max_gini]
            message("{}/{} primers have a Gini coefficient <= {}".format(len(ps), n3, max_gini))
            primers = ps
        # Primers.activate(min_active=max_primers) (primers.py:275-290): EVERY primer left is marked; the argument
        # is only the number below which the note is printed
        for p in primers:
            p.active = True
        ws.update_primers(primers, fields=("active",))
        message("Marked {} primers as active.".format(len(primers)))
        if len(primers) < max_primers:
            message("Note: Fewer than {} primers were selected ({} passed all the filters). You may want to try "
                    "less restrictive filtering parameters.".format(max_primers, len(primers)))
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


# ---------------------------------------------------------------------------------------------- export / activate / summary

def export(what="sets", db_name=workspace.DEFAULT_DB_FNAME, output=None, order_by=None, descending=False, limit=-1,
           ids=None, no_header=False, opts_str=None, window_size=10000, step_size=None, output_folder=None):
    """Export.run (commands/export.py:36-76).  `output`: a writable text file (default stdout)."""
    ws, meta = _open(db_name)
    out = output if output is not None else sys.stdout
    limit = None if limit is None or limit < 0 else limit
    with ws:
        sel = dict(ids=ids, order_by=order_by, descending=descending, limit=limit)
        if what in ("set", "sets"):
            _export.export_rows(_export.SET_FIELDS, _export.select_rows(ws, "sets", **sel), out, not no_header)
        if what in ("primer", "primers"):
            _export.export_rows(_export.PRIMER_FIELDS, _export.select_rows(ws, "primers", **sel), out, not no_header)
        if what == "lorenz" or "bed" in what:
            sets_ = _export.select_rows(ws, "sets", **sel)
            locs = {}
            for st in sets_:
                ps = {p.seq: p for p in ws.primers_by_seqs(st["primers"])}
                locs[st["_id"]] = [(seq, ps[seq].locations) for seq in st["primers"]]
            if what == "lorenz":
                chr_ends = locate.chromosome_ends(meta["fg_file"])
                _export.export_lorenz([(i, [l for _, l in pl]) for i, pl in locs.items()], out, chr_ends, not no_header)
            outpath = output_folder if output_folder else os.getcwd()
            if what == "bedfile":
                for i, pl in locs.items():
                    message("Exporting set {} and its primers as bedfiles...".format(i))
                    _export.write_bedfiles(i, pl, meta["fg_file"], outpath)
            elif what == "bedgraph":
                chr_ends = locate.chromosome_ends(meta["fg_file"])
                for i, pl in locs.items():
                    message("Exporting set {} as bedgraph...".format(i))
                    _export.write_bedgraph(i, [l for _, l in pl], meta["fg_file"], chr_ends, outpath, opts_str,
                                           window_size, step_size)


def activate(input, db_name=workspace.DEFAULT_DB_FNAME, reset=False, tm_fn=None, gindex=None):
    """Activate.run (commands/activate.py:6-18): fill in melting temperatures (swga_b200.melting unless a `tm_fn`
    is supplied) and binding locations of the listed primers, mark them active.  `reset` is accepted and, as in
    the reference, not used."""
    ws, meta = _open(db_name)
    with ws:
        with open(input) as f:
            seqs = kmers.parse_kmer_file(f)
        ps = ws.primers_by_seqs(seqs)
        have = set(p.seq for p in ps)
        for s_ in seqs:
            if s_ not in have:
                message(s_ + " not in the database; skipping. Add it manually with `swga count --input <file>` ")
        if tm_fn is None:
            from . import melting
            tm_fn = melting.temp
        todo = [p for p in ps if p.tm is None]
        for p in todo:
            p.tm = tm_fn(p.seq)
        ws.update_primers(todo, fields=("tm",))
        todo = [p for p in ps if p.locations is None]
        if todo:
            message("Finding binding locations for {} primers...".format(len(todo)))
            gindex = gindex or locate.genome_index(meta["fg_file"])
            pipeline.update_locations(todo, gindex)
            ws.update_primers(todo, fields=("_locations",))
        for p in ps:
            p.active = True
        ws.update_primers(ps, fields=("active",))
        message("Marked {} primers as active.".format(len(ps)))
        if len(ps) < 1:
            message("Note: Fewer than 1 primers were selected ({} passed all the filters). You may want to try less "
                    "restrictive filtering parameters.".format(len(ps)))
        return [p.seq for p in ps]


def summary(db_name=workspace.DEFAULT_DB_FNAME, out=None):
    """Summary.run (commands/summary.py:55-124) as a dict; the text is written to `out` (default stdout)."""
    ws, meta = _open(db_name)
    with ws:
        avg_fg, avg_bg, nprimers = ws.db.execute('SELECT AVG("fg_freq"), AVG("bg_freq"), COUNT("seq") FROM "primer"').fetchone()
        avg_fg, avg_bg = (0, 0) if avg_fg is None or avg_bg is None else (avg_fg, avg_bg)
        nactive = ws.db.execute('SELECT COUNT(*) FROM "primer" WHERE "active" = 1').fetchone()[0]
        min_tm, max_tm, avg_tm = ws.db.execute('SELECT MIN("tm"), MAX("tm"), AVG("tm") FROM "primer" WHERE "active" = 1').fetchone()
        nsets = ws.count_sets()
        best = ws.sets('"score"')[0] if nsets > 0 else None
    info = dict(fg_genome_fp=meta["fg_file"], bg_genome_fp=meta["bg_file"], exclude_fp=meta["ex_file"], nprimers=nprimers,
                nactive=nactive, avg_fg_bind=avg_fg, avg_bg_bind=avg_bg, fg_bind_ratio=avg_fg / float(meta["fg_length"]),
                bg_bind_ratio=avg_bg / float(meta["bg_length"]), min_tm=min_tm, max_tm=max_tm, avg_tm=avg_tm, nsets=nsets,
                best_set=best)
    lines = ["swga_b200 (workspace v{})".format(meta["version"]), "---------------",
             "  Foreground genome:\n    {fg_genome_fp}\n  Background genome:\n    {bg_genome_fp}\n"
             "  Excluded sequences:\n    {exclude_fp}".format(**info), "", "PRIMER SUMMARY", "---------------",
             "There are {} primers in the database. {}".format(nprimers, "Run `swga count` to find possible primers."
                                                               if nprimers == 0 else ""),
             "{} are marked as active. {}".format(nactive, "Run `swga filter` to identify primers to use."
                                                  if nactive == 0 else ""),
             "The average number of foreground genome binding sites is {avg_fg_bind:.0f}.\n"
             "    (avg binding / genome_length = {fg_bind_ratio:05f})".format(**info),
             "The average number of background genome binding sites is {avg_bg_bind:.0f}.\n"
             "    (avg binding / genome_length = {bg_bind_ratio:05f})".format(**info),
             ("The melting temp of the primers ranges between {min_tm:.2f}C and {max_tm:.2f}C with an average of "
              "{avg_tm:.2f}C.".format(**info) if nactive > 0 and min_tm and max_tm else
              "No melting temps have been calculated yet."), "", "SETS SUMMARY", "---------------",
             "There are {} sets in the database.".format(nsets)]
    if best is not None:
        lines += ["The best scoring set is #{}, with {} primers and a score of {:03f}.".format(
                      best["_id"], best["set_size"], best["score"]),
                  "Various statistics:"] + \
                 [" - {key:.<16}: {val: >10s}".format(key=k, val=best[k] if isinstance(best[k], str) else "{:,G}".format(best[k]))
                  for k in _export.SET_FIELDS if k not in ("_id", "pids", "score", "primers")] + \
                 ["The primers in Set {} are:".format(best["_id"]), ", ".join(best["primers"])]
    else:
        lines.append("Run `swga find_sets` after identifying valid primers to begin collecting sets.")
    (out if out is not None else sys.stdout).write("\n".join(lines) + "\n")
    return info


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
    p = sub.add_parser("export")
    p.add_argument("what", nargs="?", default="sets",
                   choices=["set", "sets", "primer", "primers", "bedfile", "bedgraph", "lorenz"])
    p.add_argument("--output", type=argparse.FileType("w"))
    p.add_argument("--order_by")
    p.add_argument("--descending", action="store_true")
    p.add_argument("--limit", type=int)
    p.add_argument("--ids", type=int, nargs="*")
    p.add_argument("--no_header", action="store_true")
    p.add_argument("--opts_str")
    p.add_argument("--window_size", type=float)
    p.add_argument("--step_size", type=int)
    p.add_argument("--output_folder")
    p = sub.add_parser("activate")
    p.add_argument("input")
    p.add_argument("--reset", action="store_true")
    sub.add_parser("summary")
    args, _unknown = ap.parse_known_args(argv)                  # unknown options are tolerated (_command.py:48)
    kw = {k: v for k, v in vars(args).items() if k not in ("cmd", "db") and v is not None and v is not False}
    if args.cmd == "init":
        init(args.fg_genome_fp, args.bg_genome_fp, args.exclude_fp, db_name=args.db, force=args.force)
    else:
        fn = {"count": count, "filter": filter, "find_sets": find_sets, "score": score_set, "export": export,
              "activate": activate, "summary": summary}[args.cmd]
        fn(db_name=args.db, **kw)
    return 0


if __name__ == "__main__":
    sys.exit(main())
