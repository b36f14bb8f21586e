"""`swga export` / `activate` / `summary` helpers (swga/export.py, swga/commands/export.py,
swga/stats.py:57-81).  Cold host code over the workspace database: the site lists it reads are the `_locations`
the device locate wrote during `filter`; the windows and the curve are numpy restatements of the reference's
per-base Python loops, with its quirks kept:

  * BedGraph takes a record's length from `chromosome_ends(...)[record][1] + 1` (export.py:99-100), which is the
    record's END ON THE CONCATENATED GENOME plus one, so records after the first get windows past their real end
    (they hold no sites);
  * the window range is `xrange(0, int(record_length - window), step)` (export.py:123-124): the last full window
    is never written, and a record no longer than the window yields no line;
  * a site is counted once per primer list entry (`Counter.update(Counter(locations))`, export.py:119-121): a
    palindromic primer, whose forward and reverse-complement lists hold the same positions, counts twice;
  * the Lorenz values are written with Python 2's `str(float)` (12 significant digits, commands/export.py:157).

Records are visited in file order (the reference iterates Python-2 dicts, whose order is arbitrary).
"""
import csv
import os

import numpy as np

from . import locate

PRIMER_FIELDS = ("seq", "fg_freq", "bg_freq", "ratio", "gini", "tm")                       # workspace.py:125-134
SET_FIELDS = ("_id", "score", "set_size", "bg_dist_mean", "fg_max_dist", "fg_dist_mean", "fg_dist_std",
              "fg_dist_gini", "scoring_fn", "primers")                                       # workspace.py:189-203


def _message(msg):
    import sys
    sys.stderr.write(msg + "\n")


def py2_str(x):
    """str() of a Python-2 float: '%.12g', with '.0' appended when that looks like an integer."""
    if not isinstance(x, float):
        return str(x)
    s = "%.12g" % x
    return s if any(c in s for c in ".enai") else s + ".0"


# ------------------------------------------------------------------------------------------ tables

def export_rows(fields, rows, outfile, header=True):
    """commands/export.py:112-139: tab-delimited rows of the exported fields; a list field (a set's primers)
    is joined with ','; None is written as the empty string (csv's rule)."""
    w = csv.DictWriter(outfile, list(fields), delimiter="\t", extrasaction="ignore", lineterminator="\n")
    if header:
        w.writeheader()
    for row in rows:
        d = {}
        for f in fields:
            v = row[f] if isinstance(row, dict) else getattr(row, f)
            if isinstance(v, (list, tuple)):
                v = ",".join(str(x) for x in v)
            d[f] = v
        w.writerow(d)


def select_rows(ws, what, ids=None, order_by=None, descending=False, limit=None):
    """Export.get_items (commands/export.py:78-109): rows by id, or ordered / limited."""
    table_fields = {"primers": ("_id", "seq", "fg_freq", "bg_freq", "ratio", "tm", "_locations", "gini", "active"),
                    "sets": ("_id", "_hash", "score", "set_size", "bg_dist_mean", "fg_dist_std", "fg_dist_gini",
                             "fg_max_dist", "fg_dist_mean", "scoring_fn")}[what]
    if ids and (order_by or limit):
        _message("ID(s) specified: ignoring --order_by/--limit options.")
        order_by = limit = None
    if order_by and order_by not in table_fields:
        raise ValueError("Cannot order by '{}'. Valid choices are {}".format(order_by, ", ".join(table_fields)))
    order = ('"%s"%s' % (order_by, " DESC" if descending else "")) if order_by else ""
    if what == "primers":
        rows = ws.primers('"_id" IN (%s)' % ",".join("?" * len(ids)) if ids else "", list(ids or ()), order)
    else:
        rows = ws.sets(order or '"_id"')
        if ids:
            rows = [r for r in rows if r["_id"] in set(ids)]
    if limit is not None and limit >= 0 and not ids:
        rows = rows[:limit] if limit else rows                 # `if limit:` -- 0 means no limit (export.py:103)
    return rows


# ------------------------------------------------------------------------------------------ Lorenz

def lorenz(distances):
    """stats.lorenz (stats.py:57-81): running total of the sorted values over the grand total."""
    d = np.sort(np.asarray(distances, dtype=np.int64))
    if len(d) == 0:
        raise ValueError("max() arg is an empty sequence")
    lz = np.cumsum(d)
    return (lz / float(lz.max())).tolist()


def set_lorenz(primer_locations, chr_ends):
    """The curve of one set (commands/export.py:150-156): gaps of the linearised, de-duplicated sites."""
    sites = np.sort(np.asarray(locate.linearize_binding_sites(primer_locations, chr_ends), dtype=np.int64))
    return lorenz(np.diff(sites))


def export_lorenz(sets, outfile, chr_ends, header=True):
    """`sets`: iterable of (set id, [locations dict per primer])."""
    w = csv.DictWriter(outfile, fieldnames=["SetID", "CDF"], delimiter="\t", lineterminator="\n")
    if header:
        w.writeheader()
    for set_id, locs in sets:
        w.writerow({"SetID": set_id, "CDF": ",".join(py2_str(float(v)) for v in set_lorenz(locs, chr_ends))})


# ------------------------------------------------------------------------------------------ BED

def _mk_folder(folder, fg_genome_fp, setstr):                                                # export.py:9-15
    fg_name = ".".join(os.path.basename(fg_genome_fp).split(".")[:-1]) + "_export"
    out = os.path.join(folder, fg_name, setstr)
    if not os.path.isdir(out):
        os.makedirs(out)
    return out


def write_bedfiles(set_id, primers, fg_genome_fp, output_fp):
    """BedFile.write (export.py:18-58): one BED file per primer and `whole_set.bed` with all of them.
    `primers`: [(seq, locations dict)]."""
    folder = _mk_folder(output_fp, fg_genome_fp, "set_%s" % set_id)
    with open(os.path.join(folder, "whole_set.bed"), "w") as whole:
        whole.write("track name=Set_{}\n".format(set_id))
        for seq, locations in primers:
            with open(os.path.join(folder, seq + ".bed"), "w") as pf:
                pf.write("track name={}\n".format(seq))
                lines = ["{} {} {}\n".format(rec, l, l + len(seq)) for rec, locs in locations.items() for l in locs]
                pf.writelines(lines)
                whole.writelines(lines)
    _message("Bedfiles written to " + folder)
    return folder


def bedgraph_rows(primer_locations, chr_ends, window_size, step_size=None, warn=_message):
    """BedGraph._hits_per_record (export.py:91-133) -> [(record, midpoint, hits)].  The per-base Counter sum over
    every window becomes two binary searches in the record's sorted site list."""
    window_size = int(window_size)
    step_size = int(step_size) if step_size else window_size // 5
    rows = []
    for rec, ends in chr_ends.items():
        record_length = ends[1] + 1
        w = window_size
        if w > record_length:
            w = record_length
            warn("In [{}]: window size larger than record; set to {}".format(rec, w))
        s = step_size
        if s > w:
            s = w
            warn("In [{}]: step size larger than window size ({}), set to {}".format(rec, w, s))
        sites = np.sort(np.concatenate([np.asarray(loc[rec], dtype=np.int64) for loc in primer_locations] +
                                       [np.zeros(0, dtype=np.int64)]))
        if s == 0:
            raise ValueError("xrange() arg 3 must not be zero")            # window_size < 5 without a step_size
        starts = np.arange(0, int(record_length - w), s, dtype=np.int64)
        hits = np.searchsorted(sites, starts + w, side="left") - np.searchsorted(sites, starts, side="left")
        mids = (2 * starts + w) // 2
        rows += [(rec, int(m), int(h)) for m, h in zip(mids.tolist(), hits.tolist())]
    return rows


def write_bedgraph(set_id, primer_locations, fg_genome_fp, chr_ends, output_fp, opts_str=None, window_size=10000,
                   step_size=None):
    """BedGraph.write (export.py:135-148)."""
    folder = _mk_folder(output_fp, fg_genome_fp, "set_%s" % set_id)
    fp = os.path.join(folder, "set_{}.bedgraph".format(set_id))
    with open(fp, "w") as f:
        f.write("track type=bedGraph {}\n".format(opts_str))
        f.writelines("{} {} {} {}\n".format(rec, m, m, h)
                     for rec, m, h in bedgraph_rows(primer_locations, chr_ends, window_size, step_size))
    _message("Bedgraph written to {}".format(fp))
    return fp
