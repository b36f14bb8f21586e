"""Scratch: k_score_dense variants side by side (routed = 0, bitmap = 1) on the C4-dense sets: sets/s and equality of
every output array.  python tools/ab_dense.py [n_sets]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swga_b200 import _lib, fasta, kmers, locate, score, synth, workloads

S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
lib, ctx = _lib.load(), _lib.context(0)
seed, n, gc, n_rec = synth.CONFIGS["C3_fg"]
seq = synth.bases(seed, 0, n, gc)
rs = synth.equal_records(n, n_rec)
g = fasta.Genome(seq, rs, ["chr%d" % (i + 1) for i in range(n_rec)])
tabs = kmers.count_all(g, 5, 12)
prim = workloads.top_primers(tabs, 5, 12, 200)
seqs = [kmers.codes_to_kmers([c], k)[0] for k, c, _ in prim]
gi = locate.GenomeIndex(g)
sl = score.SiteLists.from_index(gi, seqs)
sets_, m = workloads.c4_sets(S)
mapped = np.where(sets_ >= 0, sl.list_of_primer[np.clip(sets_, 0, 199)], -1).astype(np.int32)
d_sets = torch.from_numpy(mapped).cuda()
bg = torch.full((S,), 1000.0, dtype=torch.float64, device="cuda")
out = {}
for max_fg in (36000, 200, 2 ** 31 - 1):
    for impl in (0, 1):
        _lib.check(lib.swga_ctx_set_score_dense_impl(ctx.h, impl))
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = score.score_sets(sl, None, bg, max_fg, d_sets=d_sets, return_device=True)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        arrs = {k: getattr(res, k).cpu().numpy() for k in ("n_sites", "fg_max_dist", "fg_dist_mean", "fg_dist_std", "fg_dist_gini", "score", "status")}
        out[impl] = arrs
        print("max_fg=%d impl=%d: %d sets in %.2f ms -> %.4g sets/s; sites/set %.0f passed %.1f%% flag2 %.1f%%" % (
            max_fg, impl, S, ms, S / ms * 1e3, arrs["n_sites"].view(np.uint32).mean(), 100 * (arrs["status"] & 1).mean(),
            100 * ((arrs["status"] & 2) != 0).mean()), flush=True)
    for k in out[0]:
        a, b = out[0][k], out[1][k]
        same = np.array_equal(a.view(np.uint8), b.view(np.uint8))
        if not same and a.dtype.kind == "f":
            close = np.allclose(a, b, rtol=1e-12, equal_nan=True)
            print("  %s: bytes differ, allclose(1e-12)=%s, max rel %.3g" % (k, close, np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))
        else:
            print("  %s: %s" % (k, "identical" if same else "DIFFERENT"))
_lib.check(lib.swga_ctx_set_score_dense_impl(ctx.h, 0))
