"""The callers either side of the hot path (8f ranks 1 and 3), re-hosted
on the device kernels: `swga count`'s merge, `swga filter`'s locate + per-primer Gini, and the
`swga find_sets` scoring loop fed by the host `set_finder`.

  count_primers   swga/commands/count.py:70-135   (kernels a, b + k_filter_primers)
  limit_to        swga/primers.py:168-185
  update_locations / update_gini   swga/primers.py:213-239, swga/workspace.py:150-158   (kernels c, d)
  assign_ids / build_graph         swga/primers.py:255-268, swga/graph.py:7-60 (host, P <= a few hundred)
  find_sets       swga/commands/find_sets.py:22-103 (batched: kernel d scores a block of set_finder lines
                  per launch; pass / max_sets bookkeeping is applied in stream order so set ids match)
"""
import itertools
import sys

import numpy as np

from . import _lib, fasta, kmers, locate, score, sets


class Primer(object):
    """Plain record with the columns of the reference's `primer` table (swga/workspace.py:115-123)."""

    __slots__ = ("_id", "seq", "fg_freq", "bg_freq", "ratio", "tm", "locations", "gini", "active")

    def __init__(self, seq, fg_freq=0, bg_freq=0, ratio=0.0):
        self._id, self.seq, self.fg_freq, self.bg_freq, self.ratio = None, seq, fg_freq, bg_freq, ratio
        self.tm, self.locations, self.gini, self.active = None, None, None, False

    def __repr__(self):
        return "Primer {0}:{1}".format(self._id, self.seq)


def _message(msg):
    sys.stderr.write(msg + "\n")


# ------------------------------------------------------------------------------ swga count

def count_tables_device(genome, kmin, kmax, device=None):
    """Canonical tables of one genome left on the device (int32 CUDA tensor).  The bases are streamed from the
    host buffer in chunks (H2D overlapped with pack + count; a pageable buffer goes through the library's pinned
    staging blocks, which is cheaper than pinning a genome-sized buffer for one pass)."""
    import torch
    image = None
    if not isinstance(genome, fasta.Genome):
        image = fasta.read_image(genome)                    # a path: parsed chunk by chunk inside the library
    lib = _lib.load()
    ctx = _lib.context(device)
    dev = torch.device("cuda", ctx.device)
    words = int(lib.swga_tables_words(kmin, kmax))
    with torch.cuda.device(dev):
        cs = torch.cuda.ExternalStream(lib.swga_ctx_compute_stream(ctx.h), device=dev)
        with torch.cuda.stream(cs):
            fwd = torch.zeros(words, dtype=torch.int32, device=dev)
            canon = torch.zeros(words, dtype=torch.int32, device=dev)
            if image is not None:
                _lib.check(lib.swga_count_fasta_stream(ctx.h, _lib.ptr(image) if len(image) else None, len(image), -1,
                                                       kmin, kmax, _lib.ptr(fwd), None, None, None))
                _lib.check(lib.swga_count_finalize(ctx.h, kmin, kmax, _lib.ptr(fwd), cs.cuda_stream))
                _lib.check(lib.swga_fold_canonical(ctx.h, kmin, kmax, _lib.ptr(fwd), _lib.ptr(canon), cs.cuda_stream))
            elif genome.n_bases:
                rs = np.ascontiguousarray(genome.rec_starts, dtype=np.uint64)
                _lib.check(lib.swga_count_stream_host(ctx.h, _lib.ptr(genome.seq), genome.n_bases, _lib.ptr(rs),
                                                      genome.n_records, int(genome.dsk_mode_b), kmin, kmax,
                                                      genome.n_bases, _lib.ptr(fwd)))
                _lib.check(lib.swga_count_finalize(ctx.h, kmin, kmax, _lib.ptr(fwd), cs.cuda_stream))
                _lib.check(lib.swga_fold_canonical(ctx.h, kmin, kmax, _lib.ptr(fwd), _lib.ptr(canon), cs.cuda_stream))
        cs.synchronize()
    return canon


def count_primers(fg_genome, bg_genome, ex_genome=None, min_size=5, max_size=12, min_fg_bind=0, max_bg_bind=-1,
                  max_dimer_bp=3, exclude_threshold=1, add_revcomp=True, device=None):
    """What `Count.count_kmers` inserts into the primer table (count.py:70-114): for every k in
    [min_size, max_size], every canonical foreground k-mer that passes primer_dict (count.py:117-135)
    and is not in the exclusion set, as {'seq','fg_freq','bg_freq','ratio'}; with add_revcomp the reverse
    complement of each is added with identical numbers (primers.py:71-82).  Row order is unspecified
    (as in the reference)."""
    import ctypes
    import torch
    lib = _lib.load()
    ctx = _lib.context(device)
    P = _lib.ptr
    fg = count_tables_device(fg_genome, min_size, max_size, device)
    bg = count_tables_device(bg_genome, min_size, max_size, device)
    ex = count_tables_device(ex_genome, min_size, max_size, device) if ex_genome is not None else None
    dev = fg.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        cap = 1 << 16
        while True:
            rows = torch.empty(cap * 4, dtype=torch.int32, device=dev)
            n_rows = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.check(lib.swga_filter_primers(ctx.h, P(fg), P(bg), P(ex), min_size, max_size, max(0, int(min_fg_bind)),
                                               int(max_bg_bind), int(max_dimer_bp), max(1, int(exclude_threshold)),
                                               P(rows), cap, P(n_rows), stream))
            n = int(n_rows.item())
            if n <= cap:
                break
            cap = n
        r = rows[: 4 * n].cpu().numpy().view(np.uint32).reshape(n, 4)
    out = []
    for k in range(min_size, max_size + 1):
        sel = r[r[:, 0] == k]
        seqs = kmers.codes_to_kmers(sel[:, 1], k)
        for s, f, b in zip(seqs, sel[:, 2].tolist(), sel[:, 3].tolist()):
            ratio = f / float(b) if b > 0 else float("inf")
            out.append({"seq": s, "fg_freq": f, "bg_freq": b, "ratio": ratio})
    if add_revcomp:
        out += [dict(p, seq=locate.revcomp(p["seq"])) for p in out]
    return out


# ------------------------------------------------------------------------------ swga filter

def limit_to(primers, n):
    """primers.py:168-185: the n least background-binding primers, then ordered by ratio descending
    (ties keep input order: the reference leaves them to SQLite)."""
    if n < 1:
        raise ValueError("n must be greater than 1")
    first = sorted(primers, key=lambda p: p.bg_freq)[:n]
    return sorted(first, key=lambda p: -p.ratio)


def update_locations(primers, gindex):
    """Primer._update_locations (workspace.py:150-152) for every primer that has none."""
    for p in primers:
        if p.locations is None:
            p.locations = gindex.binding_sites(p.seq)
    return primers


def update_gini(primers, gindex):
    """Primer._update_gini (workspace.py:154-158) for every primer that has none: the Gini coefficient of
    the gaps between the primer's own (both-strand) sites and the record bounds -- one single-primer
    set per primer through kernel (d)."""
    todo = [p for p in primers if p.gini is None]
    if not todo:
        return primers
    sl = score.SiteLists.from_index(gindex, [p.seq for p in todo])
    sets_ = np.arange(len(todo), dtype=np.int32)[:, None]
    res = score.score_sets(sl, sets_, np.ones(len(todo)), 0xFFFFFFFF, interactive=True)
    for p, g in zip(todo, res.fg_dist_gini.tolist()):
        p.gini = g
    return primers


def filter_primers(primers, gindex, min_fg_bind, max_bg_bind, max_primers=200, max_gini=0.6, tm_range=None):
    """`swga filter` (commands/filter.py:32-40) without the melting-temperature step (the `melt` package
    is not part of this path; pass primers with `.tm` set and tm_range=(lo, hi) to apply it)."""
    ps = [p for p in primers if p.fg_freq >= min_fg_bind]
    ps = [p for p in ps if p.bg_freq <= max_bg_bind]
    if tm_range is not None:
        ps = [p for p in ps if p.tm is not None and tm_range[0] <= p.tm <= tm_range[1]]
    ps = limit_to(ps, max_primers)
    update_locations(ps, gindex)
    update_gini(ps, gindex)
    ps = [p for p in ps if p.gini <= max_gini]
    for p in ps:
        p.active = True
    return ps


# ------------------------------------------------------------------------------ swga find_sets

def assign_ids(primers):
    """primers.py:255-268: ids 1..n by ratio descending."""
    ordered = sorted(primers, key=lambda p: -p.ratio)
    for i, p in enumerate(ordered):
        p._id = i + 1
    return ordered


def build_edges_host(primers, max_binding):
    """graph.py:7-17 as the reference writes it (a Python double loop per pair): the statement of what
    `build_edges` computes, kept for the known-answer tests; not used by the commands."""
    edges = []
    for p1, p2 in itertools.combinations(primers, 2):
        if (p1.seq not in p2.seq) and (p2.seq not in p1.seq):
            if kmers.max_sequential_nt(p1.seq, p2.seq) <= max_binding:
                edges.append([p1._id, p2._id])
    return edges


def build_edges(primers, max_binding, device=None):
    """graph.py:7-17: heterodimer-compatible pairs (neither contains the other, longest complementary run
    <= max_binding), in itertools.combinations order -- all pairs in one launch (swga_build_edges)."""
    import torch
    primers = list(primers)
    n = len(primers)
    if n < 2:
        return []
    seqs = [p.seq for p in primers]
    for s in seqs:
        if set(s) - set("ACGT"):
            raise KeyError("primer %r: max_sequential_nt knows A, C, G, T only (swga/kmers.py:64-70)" % s)
    lens = np.asarray([len(s) for s in seqs], dtype=np.uint32)
    stride = (int(lens.max()) + 15) // 16 * 16
    rows = np.zeros((n, stride), dtype=np.uint8)
    for i, s in enumerate(seqs):
        rows[i, :len(s)] = np.frombuffer(s.encode("ascii"), dtype=np.uint8)
    lib = _lib.load()
    ctx = _lib.context(device)
    dev = torch.device("cuda", ctx.device)
    wpr = (n + 31) // 32
    with torch.cuda.device(dev):
        d_rows, d_len = torch.from_numpy(rows).to(dev), torch.from_numpy(lens.view(np.int32)).to(dev)
        d_adj = torch.zeros(n * wpr, dtype=torch.int32, device=dev)
        _lib.check(lib.swga_build_edges(ctx.h, _lib.ptr(d_rows), n, stride, _lib.ptr(d_len), int(lens.max()),
                                        int(max_binding), _lib.ptr(d_adj), wpr, torch.cuda.current_stream().cuda_stream))
        adj = d_adj.cpu().numpy()
    bits = np.unpackbits(adj.view(np.uint8).reshape(n, wpr * 4), axis=1, bitorder="little")[:, :n]
    ii, jj = np.nonzero(bits)
    ids = [p._id for p in primers]
    return [[ids[i], ids[j]] for i, j in zip(ii.tolist(), jj.tolist())]


def build_graph(primers, max_hetdimer_bind, outfile, edges_fn=None):
    """graph.py:20-60: DIMACS graph, vertex weight = bg_freq or 1."""
    primers = assign_ids(primers)
    edges = (edges_fn or build_edges)(primers, max_hetdimer_bind)
    if not edges:
        raise ValueError("No compatible primers. Try relaxing your parameters.")
    with open(outfile, "w") as out:
        out.write("p sp {} {}\n".format(len(primers), len(edges)))
        for p in primers:
            out.write("n {} {}\n".format(p._id, p.bg_freq if p.bg_freq > 0 else 1))
        for e in edges:
            out.write("e {} {}\n".format(e[0], e[1]))
    return primers


def score_lines(lines, site_lists, index_of_id, max_fg_bind_dist, max_m):
    """Parse a block of set_finder lines and score it in one launch.  Returns (parsed, result) where
    parsed[i] = (ids, bg_dist_mean) or None for an unparsable line (find_sets.py:59-63)."""
    parsed = []
    rows, bgs = [], []
    for line in lines:
        try:
            ids, bg = score.read_set_finder_line(line)
            row = [index_of_id[i] for i in ids]
        except (ValueError, KeyError):
            score.warn("Could not parse line:\n\t" + str(line))
            parsed.append(None)
            continue
        parsed.append((ids, bg))
        rows.append(row + [-1] * (max_m - len(row)))
        bgs.append(bg)
    if not rows:
        return parsed, None
    res = score.score_sets(site_lists, np.asarray(rows, dtype=np.int32), np.asarray(bgs), max_fg_bind_dist)
    return parsed, res


def find_sets(primers, gindex, bg_length, min_bg_bind_dist, max_fg_bind_dist, min_size=2, max_size=7,
              max_dimer_bp=3, max_sets=-1, workers=1, score_expression=score.DEFAULT_EXPRESSION,
              graph_fp="compatibility_graph.dimacs", batch=1 << 15, lines=None, on_status=None):
    """`FindSets.run` + `process_lines` (commands/find_sets.py:22-103).  Returns the rows `Set.add` would
    receive, in order: dicts with `_id` (1, 2, ... in pass order), `primers` (sequences), `primer_ids`,
    `score`, `scoring_fn` and the six variables.  `lines` may supply the candidate-set lines directly
    (then no graph is built and set_finder is not started)."""
    if max_sets < 1:
        max_sets = float("inf")
    ordered = assign_ids(primers) if any(p._id is None for p in primers) else sorted(primers, key=lambda p: p._id)
    own_gen = lines is None
    sl = score.SiteLists.from_index(gindex, [p.seq for p in ordered])
    index_of_id = {p._id: i for i, p in enumerate(ordered)}
    default_expr = score_expression.strip() == score.DEFAULT_EXPRESSION
    out = []
    passed = processed = 0
    smallest_max_dist = float("inf")

    def emit(res, j, ids, bg):
        variables = score._variables(res, j, len(ids), bg)
        sc = float(res.score[j]) if default_expr else score.evaluate_expression(score_expression, variables)
        out.append(dict(variables, _id=passed, primer_ids=ids, primers=[ordered[index_of_id[i]].seq for i in ids],
                        score=sc, scoring_fn=score_expression))

    if own_gen and workers <= 1:
        # ---- block path: raw stdout blocks -> native parser -> one scoring launch per block ----
        build_graph(ordered, max_dimer_bp, graph_fp)
        _message("Finding sets. If nothing appears, try relaxing your parameters.")
        max_m = max(max_size, 1)
        lut = np.full(max(index_of_id) + 2, -1, dtype=np.int32)               # set_finder vertex id -> primer index
        for i, k in index_of_id.items():
            lut[i] = k
        blocks = sets.find_blocks(min_bg_bind_dist=min_bg_bind_dist, min_size=min_size, max_size=max_size,
                                  bg_length=bg_length, graph_fp=graph_fp)
        carry = b""
        try:
            done = False
            for chunk in blocks:
                buf = carry + chunk
                ids, weight, ok, consumed = score.parse_set_lines(buf, max_m)
                carry = buf[consumed:]
                if len(ids) == 0:
                    continue
                in_range = (ids >= -1) & (ids < len(lut))
                idx = np.where(ids >= 0, lut[np.clip(ids, 0, len(lut) - 1)], -1)
                ok &= in_range.all(axis=1) & ((idx >= 0) == (ids >= 0)).all(axis=1)
                if not ok.all():
                    text = buf[:consumed].split(b"\n")
                    for li in np.nonzero(~ok)[0]:
                        score.warn("Could not parse line:\n\t" + text[int(li)].decode("ascii", "replace"))
                sel = np.nonzero(ok)[0]
                if len(sel) == 0:
                    continue
                res = score.score_sets(sl, idx[sel], weight[sel], max_fg_bind_dist)
                md = res.fg_max_dist
                for j in np.nonzero(res.passed)[0]:
                    j = int(j)
                    passed += 1
                    row_ids = [int(x) for x in ids[sel[j]] if x >= 0]
                    emit(res, j, row_ids, float(weight[sel[j]]))
                    if passed >= max_sets:
                        processed += j + 1
                        smallest_max_dist = min(smallest_max_dist, int(md[:j + 1].min()))
                        _message("\nDone (scored %i sets)" % passed)
                        done = True
                        break
                if done:
                    break
                processed += len(sel)
                smallest_max_dist = min(smallest_max_dist, int(md.min()))
                if on_status:
                    on_status(processed, passed, smallest_max_dist)
        finally:
            blocks.close()
        return out

    if own_gen:
        build_graph(ordered, max_dimer_bp, graph_fp)
        _message("Finding sets. If nothing appears, try relaxing your parameters.")
        lines = sets.find(min_bg_bind_dist=min_bg_bind_dist, min_size=min_size, max_size=max_size,
                          bg_length=bg_length, graph_fp=graph_fp, workers=workers)
    it = iter(lines)
    n_take = 64
    try:
        done = False
        while not done:
            block = list(itertools.islice(it, n_take))
            if not block:
                break
            n_take = min(batch, n_take * 4)
            parsed, res = score_lines(block, sl, index_of_id, max_fg_bind_dist, max(max_size, 1))
            j = 0
            for p in parsed:
                if p is None:
                    continue
                ids, bg = p
                processed += 1
                max_dist = int(res.fg_max_dist[j])
                smallest_max_dist = min(smallest_max_dist, max_dist)
                if on_status:
                    on_status(processed, passed, smallest_max_dist)
                if res.passed[j]:
                    passed += 1
                    emit(res, j, ids, bg)
                    if passed >= max_sets:
                        _message("\nDone (scored %i sets)" % passed)
                        done = True
                        break
                j += 1
    finally:
        if own_gen:
            lines.close()
    return out
