"""Synthetic workload C4 of SURVEY.md 8(d): 200 primers over the 23 Mbp AT-rich foreground and 10^7
candidate sets of 6-10 primers, all position-indexed integer recipes (bench + tests agree bit for bit)."""
import numpy as np

from . import synth


def c4_sets(n_sets, n_primers=200, seed=8, start=0):
    """sets[s] = the first m(s) distinct values of h((seed<<40) + 64 s + t) mod P, t = 1, 2, ...,
    m(s) = 6 + h((seed<<40) + 64 s) mod 5; -1 padded to 10 columns."""
    s = np.arange(start, start + n_sets, dtype=np.uint64)
    base = (np.uint64(seed) << np.uint64(40)) + np.uint64(64) * s
    m = (6 + (synth.splitmix64(base) % np.uint64(5))).astype(np.int32)
    out = np.full((n_sets, 10), -1, dtype=np.int32)
    filled = np.zeros(n_sets, dtype=np.int32)
    for t in range(1, 64):
        need = filled < m
        if not need.any():
            break
        c = (synth.splitmix64(base + np.uint64(t)) % np.uint64(n_primers)).astype(np.int32)
        dup = (out == c[:, None]).any(axis=1)
        take = need & ~dup
        idx = np.nonzero(take)[0]
        out[idx, filled[idx]] = c[idx]
        filled[idx] += 1
    return out, m


def top_primers(canon_tables, kmin, kmax, n=200, k_lo=8, k_hi=12):
    """The n canonical k-mers, k in [k_lo, k_hi], with the largest foreground count (ties by k then code).
    canon_tables: kmers.CountTables on the host.  Returns [(k, code, count)]."""
    cand = []
    for k in range(max(kmin, k_lo), min(kmax, k_hi) + 1):
        t = canon_tables.table(k)
        top = np.argpartition(t, -min(n, len(t)))[-min(n, len(t)):]
        cand += [(int(t[c]), k, int(c)) for c in top]
    cand.sort(key=lambda x: (-x[0], x[1], x[2]))
    return [(k, c, cnt) for cnt, k, c in cand[:n]]
