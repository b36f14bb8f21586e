"""Scratch: C4 scoring throughput (sets/s) on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swga_b200 import _lib, fasta, kmers, locate, score, synth, workloads

S = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
mode = sys.argv[2] if len(sys.argv) > 2 else "dense"
lib, ctx = _lib.load(), _lib.context(0)
seed, n, gc, n_rec = synth.CONFIGS["C3_fg"]
t0 = time.time()
seq = synth.bases(seed, 0, n, gc)
rs = synth.equal_records(n, n_rec)
g = fasta.Genome(seq, rs, ["chr%d" % (i + 1) for i in range(n_rec)])
tabs = kmers.count_all(g, 5, 12)
if mode == "dense":
    prim = workloads.top_primers(tabs, 5, 12, 200)
else:   # "typical": 200 12-mers with a fg count around the reference's default min_fg_bind scale
    t12 = tabs.table(12)
    order = np.argsort(-t12.astype(np.int64), kind="stable")
    lo = int(np.searchsorted(-t12[order].astype(np.int64), -int(sys.argv[3]) if len(sys.argv) > 3 else -400))
    prim = [(12, int(c), int(t12[c])) for c in order[lo:lo + 200]]
seqs = [kmers.codes_to_kmers([c], k)[0] for k, c, _ in prim]
print("primers: k in", sorted({k for k, _, _ in prim}), "fg counts", prim[0][2], "...", prim[-1][2], "setup %.1fs" % (time.time() - t0), flush=True)
gi = locate.GenomeIndex(g)
t0 = time.time()
sl = score.SiteLists.from_index(gi, seqs)
torch.cuda.synchronize()
print("index+sitelists %.3fs; sites per primer mean %.0f; total %d" % (time.time() - t0, sl.counts()[:-1].mean(), sl.counts().sum()), flush=True)
sets_, m = workloads.c4_sets(S)
mapped = np.where(sets_ >= 0, sl.list_of_primer[np.clip(sets_, 0, 199)], -1).astype(np.int32)
d_sets = torch.from_numpy(mapped).cuda()
bg = torch.full((S,), 1000.0, dtype=torch.float64, device="cuda")
for max_fg in (36000, 2 ** 31 - 1):
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = score.score_sets(sl, None, bg, max_fg, d_sets=d_sets, return_device=True)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = res.status.cpu().numpy()
    L = res.n_sites.cpu().numpy().view(np.uint32)
    print("max_fg=%d: %d sets in %.2f ms -> %.3g sets/s; mean unique sites/set %.0f; passed %.1f%%; unbounded-flag %.1f%%"
          % (max_fg, S, ms, S / ms * 1e3, L.mean(), 100 * (st & 1).mean(), 100 * ((st & 2) != 0).mean()), flush=True)
