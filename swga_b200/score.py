"""Batched scoring of candidate primer sets on the B200 -- the drop-in for swga/score.py (and the
per-set loop of swga/commands/find_sets.py:53-103).

Reference seams (same names and argument meaning):
  * `read_set_finder_line(line)`                               swga/score.py:16-24
  * `default_score_set(expression, primer_set, primer_locs, max_dist, bg_dist_mean)`   :27-53
  * `calculate_bg_dist_mean(primers, bg_length)`               :56-70
  * `score_set(primers, max_fg_bind_dist, bg_dist_mean, chr_ends, score_fun, interactive)`  :73-98

Native form: `SiteLists` (every primer's sorted, de-duplicated, linearised site list on the device,
plus the record-bounds list and the tile offsets kernel (d) walks) and `score_sets`, which scores S
sets per launch (csrc/score.cu).  Sets shard by set index across GPUs with no collective.
"""
import ast
import ctypes
import functools
import math
import sys

import numpy as np

from . import _lib, kmers

DEFAULT_EXPRESSION = "(fg_dist_mean * fg_dist_gini) / (bg_dist_mean)"      # find_sets.yaml:39-40
VARIABLES = ("set_size", "fg_dist_mean", "fg_dist_std", "fg_dist_gini", "bg_dist_mean", "fg_max_dist")


def warn(msg):                                                              # swga/__init__.py:33-40
    sys.stderr.write("Warning: " + msg + "\n")


def read_set_finder_line(line):
    """'id,id,... weight' -> ([ids], weight)  (the weight was printed with %f, set_finder.c:442)."""
    if isinstance(line, bytes):
        line = line.decode()
    primer_set, weight = line.strip("\n").split(" ")
    return [int(_) for _ in primer_set.split(",")], float(weight)


def parse_set_lines(block, max_m, cap=None):
    """Every complete line of a block of set_finder output at once (native; what read_set_finder_line does per
    line).  Returns (ids int32 [n][max_m] -1 padded, weight float64 [n], ok bool [n], consumed bytes)."""
    lib = _lib.load()
    if isinstance(block, str):
        block = block.encode("ascii", "replace")
    cap = block.count(b"\n") if cap is None else cap
    ids = np.empty((max(cap, 1), max_m), dtype=np.int32)
    w = np.empty(max(cap, 1), dtype=np.float64)
    ok = np.zeros(max(cap, 1), dtype=np.uint8)
    n_lines, consumed = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _lib.check(lib.swga_parse_set_lines(block, len(block), max_m, cap, _lib.ptr(ids), _lib.ptr(w), _lib.ptr(ok),
                                        ctypes.byref(n_lines), ctypes.byref(consumed)))
    n = n_lines.value
    return ids[:n], w[:n], ok[:n].astype(bool), consumed.value


def calculate_bg_dist_mean(primers, bg_length):
    """float(bg_length) / sum(bg_freq), or inf with a warning when nothing binds the background."""
    total_bg_freq = sum(p.bg_freq for p in primers)
    if total_bg_freq == 0:
        warn("No primers appear in the background genome: bg_dist_mean set as infinite")
        return float("Inf")
    return float(bg_length) / total_bg_freq


# ------------------------------------------------------------------------------ expression evaluation

class _Py2Div(ast.NodeTransformer):
    """The reference evaluates the expression under Python 2: int / int floors."""

    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="__py2div", ctx=ast.Load()),
                                              args=[node.left, node.right], keywords=[]), node)
        return node


def _py2div(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) and not isinstance(a, bool):
        return a // b
    return a / b


def evaluate_expression(expression, namespace):
    """score.py:42-51: eval(expression) over the six metric names only; NameError lists them."""
    tree = ast.fix_missing_locations(_Py2Div().visit(ast.parse(expression.strip(), mode="eval")))
    env = {"__builtins__": {}, "__py2div": _py2div}
    try:
        return eval(compile(tree, "<score_expression>", "eval"), env, dict(namespace))
    except NameError as e:
        raise NameError(str(e) + ". Permitted variables are %s. Refer to README or docs for help."
                        % ", ".join(VARIABLES))


# ------------------------------------------------------------------------------ device site lists

class SiteLists(object):
    """Concatenated sorted-unique linearised site lists on the device: one per primer, then the list
    of every record's first and last position (locate.py:30-32), with tile offsets."""

    def __init__(self, sites, site_off, n_primers, genome_len, list_of_primer=None, device=None):
        import torch
        self.lib = _lib.load()
        is_t = isinstance(sites, torch.Tensor)
        self.ctx = _lib.context(device if device is not None else (sites.device.index if is_t else None))
        self.dev = torch.device("cuda", self.ctx.device)
        self.sites = sites if is_t else torch.from_numpy(
            np.ascontiguousarray(sites, dtype=np.uint32).view(np.int32)).to(self.dev)
        # swga_score_sets reads the lists with 16-byte vector loads: 16 readable words past the last site
        self.n_sites_total = int(self.sites.numel())
        self.sites = torch.cat([self.sites, torch.zeros(16, dtype=torch.int32, device=self.dev)])
        self.site_off = np.ascontiguousarray(site_off, dtype=np.uint64)
        self.n_lists = len(self.site_off) - 1
        self.n_primers = n_primers
        self.bounds_list = n_primers                      # the last list
        assert self.n_lists == n_primers + 1
        self.list_of_primer = (np.arange(n_primers, dtype=np.int32) if list_of_primer is None
                               else np.asarray(list_of_primer, dtype=np.int32))
        self.tile_shift = int(self.lib.swga_score_tile_shift())
        self.n_tiles = max(1, (int(genome_len) + (1 << self.tile_shift) - 1) >> self.tile_shift)
        P = _lib.ptr
        with torch.cuda.device(self.dev):
            self.d_site_off = torch.from_numpy(self.site_off.view(np.int64)).to(self.dev)
            self.toff = torch.empty(self.n_lists * (self.n_tiles + 1), dtype=torch.int32, device=self.dev)
            if self.sites.numel() == 0:
                self.sites = torch.zeros(1, dtype=torch.int32, device=self.dev)
            _lib.check(self.lib.swga_tile_offsets(self.ctx.h, P(self.sites), P(self.d_site_off), self.n_lists,
                                                  self.tile_shift, self.n_tiles, P(self.toff),
                                                  torch.cuda.current_stream().cuda_stream))

    def counts(self):
        return np.diff(self.site_off.astype(np.int64))

    @staticmethod
    def bounds_of(rec_starts):
        """Sorted unique {start_r, end_r} over records (chromosome_ends, locate.py:59-73)."""
        rs = np.asarray(rec_starts, dtype=np.int64)
        b = np.concatenate([rs[:-1], rs[1:] - 1])
        return np.unique(b[b >= 0]).astype(np.uint32)

    @classmethod
    def from_index(cls, gindex, primer_seqs):
        """Locate every primer (both strands) through the CSR index of `gindex` (locate.GenomeIndex) and
        union the strands on the device."""
        import torch
        from .locate import revcomp
        lib, ctx, dev = gindex.lib, gindex.ctx, gindex.dev
        P = _lib.ptr
        seqs = [str(s).upper() for s in primer_seqs]
        by_k = {}
        for i, s in enumerate(seqs):
            by_k.setdefault(len(s), []).append(i)
        parts, part_counts, list_of_primer = [], [], np.zeros(len(seqs), dtype=np.int32)
        next_list = 0
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            for k in sorted(by_k):
                idxs = by_k[k]
                codes = np.empty(2 * len(idxs), dtype=np.uint32)
                pal = np.zeros(len(idxs), dtype=np.uint8)
                for j, i in enumerate(idxs):
                    f, r = kmers.kmer_to_code(seqs[i]), kmers.kmer_to_code(revcomp(seqs[i]))
                    codes[2 * j], codes[2 * j + 1] = f, r
                    pal[j] = f == r
                hits, hit_off = gindex.locate_codes(k, codes)
                cnt = np.diff(hit_off.astype(np.int64))
                ucnt = cnt[0::2] + np.where(pal == 1, 0, cnt[1::2])
                out_off = np.concatenate([[0], np.cumsum(ucnt)]).astype(np.uint64)
                out = torch.empty(max(int(out_off[-1]), 1), dtype=torch.int32, device=dev)
                d_hit_off = torch.from_numpy(hit_off.view(np.int64)).to(dev)
                d_out_off = torch.from_numpy(out_off.view(np.int64)).to(dev)
                d_pal = torch.from_numpy(pal).to(dev)
                _lib.check(lib.swga_union_strands(ctx.h, P(hits), P(d_hit_off), P(d_pal), len(idxs), P(d_out_off),
                                                  P(out), stream))
                parts.append(out[: int(out_off[-1])])
                part_counts.append(ucnt)
                for j, i in enumerate(idxs):
                    list_of_primer[i] = next_list + j
                next_list += len(idxs)
            bounds = cls.bounds_of(gindex.rec_starts)
            parts.append(torch.from_numpy(bounds.view(np.int32)).to(dev))
            part_counts.append(np.array([len(bounds)], dtype=np.int64))
            sites = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int32, device=dev)
        counts = np.concatenate(part_counts).astype(np.uint64)
        site_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        return cls(sites, site_off, len(seqs), gindex.n, list_of_primer, device=ctx.device)

    @classmethod
    def from_host_lists(cls, lists, bounds, genome_len, device=None):
        """lists: per primer, linearised positions (any order, duplicates allowed); bounds: positions of
        record starts/ends.  Sorting / de-duplication here is host glue for stored `Primer.locations`."""
        ls = [np.unique(np.asarray(l, dtype=np.int64)).astype(np.uint32) for l in lists]
        ls.append(np.unique(np.asarray(bounds, dtype=np.int64)).astype(np.uint32))
        site_off = np.concatenate([[0], np.cumsum([len(l) for l in ls])]).astype(np.uint64)
        sites = np.concatenate(ls) if ls else np.zeros(0, dtype=np.uint32)
        return cls(sites, site_off, len(lists), genome_len, device=device)


def shard_sets(n_sets, rank, world):
    """Contiguous slice [lo, hi) of the set indices scored by `rank` (no collective is needed)."""
    per = -(-n_sets // world)
    lo = min(n_sets, rank * per)
    return lo, min(n_sets, lo + per)


MAX_SET_SIZE = 31


class ScoreResult(dict):
    __getattr__ = dict.__getitem__


def score_sets(site_lists, sets, bg_dist_mean, max_fg_bind_dist, interactive=False, d_sets=None, return_device=False):
    """Score S sets in one launch.

    sets          int array [S][max_m] of PRIMER indices (into the SiteLists' primers), -1 padded
    bg_dist_mean  float64 [S]
    Returns arrays: n_sites, fg_max_dist, fg_dist_mean, fg_dist_std, fg_dist_gini, score (default
    expression), passed (bool).  Sets that fail the max_dist filter (non-interactive) keep their
    max_dist; their other variables are not meaningful, as in score.py:89-90 (bulk calls do not even
    copy them back: they read 0).
    """
    import torch
    sl = site_lists
    lib, ctx, dev = sl.lib, sl.ctx, sl.dev
    P = _lib.ptr
    if d_sets is None:
        sets = np.asarray(sets, dtype=np.int32)
        if sets.ndim != 2:
            raise ValueError("sets must be [S][max_m]")
        S, max_m = sets.shape
    else:
        S, max_m = d_sets.shape
    if max_m > MAX_SET_SIZE:
        # kernel (d) merges at most 32 lists per set: 31 primers + the record bounds (csrc/score.cu SC_MAX_LISTS);
        # the reference has no limit, its default --max_size is 7 (find_sets.yaml:10-13)
        raise ValueError("a primer set may hold at most %d primers on the device path, got %d" % (MAX_SET_SIZE, max_m))
    mfd = max_fg_bind_dist
    mfd = 0xFFFFFFFF if (mfd == float("inf") or mfd > 0xFFFFFFFF) else max(0, int(mfd))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        if d_sets is None:
            # primer index -> list id on the device (identical primers share one list); -1 stays -1
            d_raw = torch.from_numpy(np.ascontiguousarray(sets)).to(dev)
            lop = getattr(sl, "_d_list_of_primer", None)
            if lop is None:
                lop = sl._d_list_of_primer = torch.from_numpy(
                    np.ascontiguousarray(sl.list_of_primer if sl.n_primers else np.zeros(1), dtype=np.int32)).to(dev)
            d_sets = torch.where(d_raw >= 0, lop[d_raw.clamp(0, max(sl.n_primers - 1, 0)).long()], d_raw.new_full((), -1))
        d_bg = bg_dist_mean if isinstance(bg_dist_mean, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(bg_dist_mean, dtype=np.float64)).to(dev)
        o_n = torch.empty(S, dtype=torch.int32, device=dev)
        o_max = torch.empty(S, dtype=torch.int32, device=dev)
        o_mean, o_std, o_gini, o_score = (torch.empty(S, dtype=torch.float64, device=dev) for _ in range(4))
        o_status = torch.empty(S, dtype=torch.uint8, device=dev)
        _lib.check(lib.swga_score_sets(ctx.h, P(sl.sites), P(sl.toff), sl.n_tiles, sl.bounds_list, P(d_sets), S, max_m,
                                       P(d_bg), mfd, int(bool(interactive)), P(o_n), P(o_max), P(o_mean), P(o_std),
                                       P(o_gini), P(o_score), P(o_status), stream))
        if return_device:
            return ScoreResult(n_sites=o_n, fg_max_dist=o_max, fg_dist_mean=o_mean, fg_dist_std=o_std,
                               fg_dist_gini=o_gini, score=o_score, status=o_status)
        outs = (o_n, o_max, o_mean, o_std, o_gini, o_score, o_status)
        if S <= (1 << 17):
            # block-sized calls (the find_sets loop): results come back through pinned host tensors (torch caches
            # pinned blocks, so no allocation after the first calls), seven copies in flight and ONE synchronisation;
            # the arrays returned are views of those tensors
            host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t, non_blocking=True) for t in outs]
            torch.cuda.current_stream().synchronize()
            h_n, h_max, h_mean, h_std, h_gini, h_score, status = (t.numpy() for t in host)
        else:
            # bulk call: status and max_dist of every set; the other 36 bytes per set only for the sets that passed
            # (with the default max_fg_bind_dist most candidate sets are rejected), scattered into zero-filled
            # arrays whose untouched pages are never materialised
            status, h_max = o_status.cpu().numpy(), o_max.cpu().numpy()
            n_pass = int(np.count_nonzero(status & 1))
            if 2 * n_pass > S:
                h_n, h_mean, h_std, h_gini, h_score = (t.cpu().numpy() for t in (o_n, o_mean, o_std, o_gini, o_score))
            else:
                sel = torch.nonzero(o_status & 1).squeeze(1)
                idx = sel.cpu().numpy()
                h_n = np.zeros(S, dtype=np.int32)
                h_mean, h_std, h_gini, h_score = (np.zeros(S, dtype=np.float64) for _ in range(4))
                for dst, src in ((h_n, o_n), (h_mean, o_mean), (h_std, o_std), (h_gini, o_gini), (h_score, o_score)):
                    dst[idx] = src[sel].cpu().numpy()
        res = ScoreResult(
            n_sites=h_n.view(np.uint32), fg_max_dist=h_max.view(np.uint32).astype(np.int64),
            fg_dist_mean=h_mean, fg_dist_std=h_std, fg_dist_gini=h_gini,
            score=h_score, passed=(status & 1).astype(bool), status=status)
        # sets whose largest gap exceeds the in-kernel histogram: exact Gini through the sort path
        todo = np.nonzero(((status & 2) != 0) & ((status & 1) != 0))[0]
        bg = d_bg.cpu().numpy() if len(todo) else None
        for s in todo:
            g = ctypes.c_double()
            _lib.check(lib.swga_gini_unbounded(ctx.h, P(sl.sites), P(sl.toff), sl.n_tiles, sl.bounds_list,
                                               P(d_sets[int(s)]), max_m, int(res.n_sites[s]) + 1, ctypes.byref(g), stream))
            res.fg_dist_gini[s] = g.value
            with np.errstate(all="ignore"):
                res.score[s] = (res.fg_dist_mean[s] * g.value) / bg[s]
    return res


# ------------------------------------------------------------------------------ reference-shaped calls

def _variables(res, i, set_size, bg_dist_mean):
    return {
        "set_size": set_size,
        "fg_dist_mean": float(res.fg_dist_mean[i]),
        "fg_dist_std": float(res.fg_dist_std[i]),
        "fg_dist_gini": float(res.fg_dist_gini[i]),
        "bg_dist_mean": bg_dist_mean,
        "fg_max_dist": int(res.fg_max_dist[i]),
    }


def default_score_set(expression, primer_set, primer_locs, max_dist, bg_dist_mean):
    """Evaluate `expression` over the metrics of one list of binding locations (score.py:27-53).
    Returns (score, namespace).  The gap statistics come from the device kernel."""
    locs = np.asarray(list(primer_locs), dtype=np.int64)
    if len(np.unique(locs)) != len(locs):
        # Not what linearize_binding_sites hands over (it de-duplicates, locate.py:33), but the reference takes any
        # list: stats.seq_diff keeps the zero gaps of repeated locations (stats.py:22-33).  The device kernel
        # merges sorted-unique site lists, so this rare, tiny case is evaluated on the host with the same exact
        # integer forms as the kernel: mean = H/n, gini = (2F - (2P - H)) / 2F with F = floor(Hn/2).
        return _score_duplicated_locations(expression, primer_set, locs, max_dist, bg_dist_mean)
    if len(locs) < 2:
        # the reference divides by zero in stats.mean/stdev here
        raise ZeroDivisionError("float division by zero")
    sl = SiteLists.from_host_lists([locs], [], int(locs.max()) + 1)
    res = score_sets(sl, np.array([[0]], dtype=np.int32), np.array([float(bg_dist_mean)]), 0xFFFFFFFF, interactive=True)
    if res.n_sites[0] == 2:
        raise ZeroDivisionError("float division by zero")        # stats.stdev with one gap: n - 1 == 0
    namespace = _variables(res, 0, len(primer_set), bg_dist_mean)
    namespace["fg_max_dist"] = max_dist
    if math.isnan(namespace["fg_dist_gini"]):
        raise ZeroDivisionError("float division by zero")        # stats.gini: fair_area == 0
    score = evaluate_expression(expression, namespace)
    return score, namespace


def _score_duplicated_locations(expression, primer_set, locs, max_dist, bg_dist_mean):
    gaps = np.diff(np.sort(locs)).astype(object)              # Python ints: no overflow, exact sums
    n = len(gaps)
    if n < 2:
        raise ZeroDivisionError("float division by zero")
    H = int(sum(gaps))
    mean = float(H) / n                                        # stats.py:10-12
    mu = H / float(n)
    std = math.sqrt(sum((float(x) - mu) ** 2 for x in gaps) / (float(n) - 1.0))      # stats.py:15-19, in gap order
    asc = sorted(gaps)
    prefix = 0
    two_area = 0
    for v in asc:                                              # stats.py:46-52: area += height - v / 2.
        prefix += v
        two_area += 2 * prefix - v
    fair = (H * n) // 2                                        # Python-2 floor division, stats.py:53
    if fair == 0:
        raise ZeroDivisionError("float division by zero")
    gini = (fair - two_area / 2.0) / fair
    namespace = {"set_size": len(primer_set), "fg_dist_mean": mean, "fg_dist_std": std, "fg_dist_gini": gini,
                 "bg_dist_mean": bg_dist_mean, "fg_max_dist": max_dist}
    return evaluate_expression(expression, namespace), namespace


def score_set(primers, max_fg_bind_dist, bg_dist_mean, chr_ends, score_fun, interactive=False):
    """Score one set (swga/score.py:73-98).  `primers` carry `.locations` ({record: [pos]}); returns
    (score | False, variables, max_dist)."""
    primers = list(primers)
    lists, bounds = [], []
    for p in primers:
        locations = p.locations if hasattr(p, "locations") else p
        lin = []
        for rec, locs in locations.items():
            chr_start, chr_end = chr_ends[rec]
            lin.append(np.asarray(locs, dtype=np.int64) + chr_start)
            bounds += [chr_start, chr_end]
        lists.append(np.concatenate(lin) if lin else np.zeros(0, dtype=np.int64))
    if not bounds:
        raise ValueError("Binding sites for primers not found!")
    genome_len = max(int(max(bounds)), max((int(l.max()) for l in lists if len(l)), default=0)) + 1
    sl = SiteLists.from_host_lists(lists, bounds, genome_len)
    sets = np.arange(len(primers), dtype=np.int32)[None, :] if primers else np.full((1, 1), -1, dtype=np.int32)
    bg = float(bg_dist_mean)
    res = score_sets(sl, sets, np.array([bg]), max_fg_bind_dist, interactive=interactive)
    if res.n_sites[0] < 2:
        raise ValueError("max() arg is an empty sequence")       # score.py:85 on a single location
    max_dist = int(res.fg_max_dist[0])
    if not interactive and max_dist > max_fg_bind_dist:
        return False, {}, max_dist
    variables = _variables(res, 0, len(primers), bg_dist_mean)
    expression = None
    if isinstance(score_fun, functools.partial) and score_fun.func is default_score_set:
        expression = score_fun.keywords.get("expression")
    elif isinstance(score_fun, str):
        expression = score_fun
    if expression is not None:
        if res.n_sites[0] == 2 or math.isnan(variables["fg_dist_gini"]):
            raise ZeroDivisionError("float division by zero")
        return evaluate_expression(expression, variables), variables, max_dist
    # a foreign scoring callable gets the reference's arguments (host glue: the union itself)
    locs = np.unique(np.concatenate(lists + [np.asarray(bounds, dtype=np.int64)])).tolist()
    set_score, variables = score_fun(primer_set=primers, primer_locs=locs, max_dist=max_dist, bg_dist_mean=bg_dist_mean)
    return set_score, variables, max_dist
