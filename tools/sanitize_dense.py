"""Scratch: a small dense-kernel scoring run for compute-sanitizer (memcheck / racecheck): constructed site lists that
take every form of the routed k_score_dense (sparse and bursty windows, > 16 sites per window, shared positions,
> 16 lists), all sets through the dense kernel (sort cap 0), compared with the bitmap form.
compute-sanitizer --tool racecheck python tools/sanitize_dense.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swga_b200 import _lib, score

rng = np.random.RandomState(3)
L = 400_000
bounds = [0, 199_999, 200_000, L - 1]
lists = [np.sort(rng.choice(L, n, replace=False)) for n in (50, 700, 3000, 9000, 20000)]
for step in (3, 15, 16, 20, 40):
    lists.append(np.arange(50_000 + 1000 * step, 50_000 + 1000 * step + 30_000, step))
lists.append(lists[4][::2].copy())
while len(lists) < 20:
    lists.append(np.sort(rng.choice(L, int(rng.randint(500, 4000)), replace=False)))
sl = score.SiteLists.from_host_lists(lists, bounds, L)
sets = [[0], [1, 2], [3, 4, 10], [5], [6, 7, 8, 9], list(range(20)), list(range(3, 14)), [4, 10], [2, 5, 9]]
arr = np.full((len(sets), 31), -1, dtype=np.int32)
for i, s in enumerate(sets):
    arr[i, :len(s)] = s
bgs = np.full(len(sets), 1000.0)
lib, ctx = _lib.load(), _lib.context(0)
_lib.check(lib.swga_ctx_set_score_sort_cap(ctx.h, 0))
out = {}
for impl in (0, 1):
    _lib.check(lib.swga_ctx_set_score_dense_impl(ctx.h, impl))
    r = score.score_sets(sl, arr, bgs, 2 ** 31 - 1)
    out[impl] = [np.asarray(getattr(r, k)) for k in ("n_sites", "fg_max_dist", "fg_dist_mean", "fg_dist_std", "fg_dist_gini")]
for a, b in zip(out[0], out[1]):
    assert np.array_equal(a, b, equal_nan=True), (a, b)
print("dense forms agree on %d sets; n_sites" % len(sets), out[0][0])
