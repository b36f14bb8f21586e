"""Scratch: k_score_sort front ends side by side (routed = 0, merge = 1) on the 'typical' sets (200 12-mers of ~250 sites,
6-10 primers per set): sets/s and equality of every output array.  python tools/ab_sort.py [n_sets] [fg_count]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swga_b200 import _lib, fasta, kmers, locate, score, synth, workloads

S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
fgc = int(sys.argv[2]) if len(sys.argv) > 2 else 400
lib, ctx = _lib.load(), _lib.context(0)
seed, n, gc, n_rec = synth.CONFIGS["C3_fg"]
seq = synth.bases(seed, 0, n, gc)
rs = synth.equal_records(n, n_rec)
g = fasta.Genome(seq, rs, ["chr%d" % (i + 1) for i in range(n_rec)])
tabs = kmers.count_all(g, 5, 12)
t12 = tabs.table(12)
order = np.argsort(-t12.astype(np.int64), kind="stable")
lo = int(np.searchsorted(-t12[order].astype(np.int64), -fgc))
prim = [(12, int(c), int(t12[c])) for c in order[lo:lo + 200]]
seqs = [kmers.codes_to_kmers([c], k)[0] for k, c, _ in prim]
gi = locate.GenomeIndex(g)
sl = score.SiteLists.from_index(gi, seqs)
sets_, m = workloads.c4_sets(S)
mapped = np.where(sets_ >= 0, sl.list_of_primer[np.clip(sets_, 0, 199)], -1).astype(np.int32)
d_sets = torch.from_numpy(mapped).cuda()
bg = torch.full((S,), 1000.0, dtype=torch.float64, device="cuda")
out = {}
for max_fg in (36000, 2 ** 31 - 1):
    for impl in (0, 1):
        _lib.check(lib.swga_ctx_set_score_sort_impl(ctx.h, impl))
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = score.score_sets(sl, None, bg, max_fg, d_sets=d_sets, return_device=True)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        arrs = {k: getattr(res, k).cpu().numpy() for k in ("n_sites", "fg_max_dist", "fg_dist_mean", "fg_dist_std", "fg_dist_gini", "score", "status")}
        out[impl] = arrs
        print("max_fg=%d impl=%d: %d sets in %.2f ms -> %.4g sets/s; sites/set %.0f passed %.1f%%" % (
            max_fg, impl, S, ms, S / ms * 1e3, arrs["n_sites"].view(np.uint32).mean(), 100 * (arrs["status"] & 1).mean()), flush=True)
    for k in out[0]:
        a, b = out[0][k], out[1][k]
        same = np.array_equal(a.view(np.uint8), b.view(np.uint8))
        if not same and a.dtype.kind == "f":
            ok = ~(np.isnan(a) & np.isnan(b))
            print("  %s: bytes differ; max rel %.3g; nan mismatch %d" % (k, np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)) if ok.any() else 0, int((np.isnan(a) != np.isnan(b)).sum())))
        else:
            print("  %s: %s" % (k, "identical" if same else "DIFFERENT (%d)" % int((a != b).sum())))
_lib.check(lib.swga_ctx_set_score_sort_impl(ctx.h, 0))
