"""Scratch: heterodimer graph build, host double loop (the reference's form) vs swga_build_edges."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swga_b200 import pipeline
rng = np.random.RandomState(0)
for n in (200, 1000):
    seqs = sorted({"".join("ACGT"[i] for i in rng.randint(0, 4, int(rng.randint(8, 13)))) for _ in range(n * 2)})[:n]
    ps = []
    for i, s in enumerate(seqs):
        p = pipeline.Primer(s); p._id = i + 1; ps.append(p)
    pipeline.build_edges(ps[:4], 3); torch.cuda.synchronize()
    t0 = time.perf_counter(); e_dev = pipeline.build_edges(ps, 3); t1 = time.perf_counter()
    if n <= 200:
        e_host = pipeline.build_edges_host(ps, 3); t2 = time.perf_counter()
        assert e_dev == e_host
        print("P=%d: %d edges, device %.4f s, host loop %.3f s" % (n, len(e_dev), t1 - t0, t2 - t1), flush=True)
    else:
        print("P=%d: %d edges, device %.4f s" % (n, len(e_dev), t1 - t0), flush=True)
