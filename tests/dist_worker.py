"""Worker for tests/test_dist_cpu.py: one rank of a world_size-N gloo job on CPU.

Exercises the host-side multi-GPU logic of swga_b200.dist (chunk + right-halo sharding, local record
starts, the allreduce of the additive table form, set-index sharding) with a numpy stand-in for the
CUDA accumulate kernel, and checks the merged result against the oracle on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import swga_oracle as O            # noqa: E402
from swga_b200 import dist, score, synth       # noqa: E402

KMIN, KMAX = 3, 7


def additive_tables(seq, rec_starts, n_starts, split_at_n):
    """numpy statement of what swga_count_accumulate adds for one slice (include/swga_b200.h)."""
    n = len(seq)
    code = (seq >> 1) & 3
    brk = np.zeros(n, dtype=bool)                  # brk[i]: no window may hold base i and i+1
    if split_at_n:
        bad = seq == ord("N")
        brk |= bad
        brk[:-1] |= bad[1:]
    for s in rec_starts[1:-1]:
        if 0 < s <= n:
            brk[int(s) - 1] = True
    brk[n - 1] = True
    offs = np.cumsum([0] + [4 ** k for k in range(KMIN, KMAX + 1)])
    out = np.zeros(offs[-1], dtype=np.int64)
    nxt = np.full(n + 1, n, dtype=np.int64)        # next break index at or after i
    for i in range(n - 1, -1, -1):
        nxt[i] = i if brk[i] else nxt[i + 1]
    for p in range(n_starts):
        v = min(KMAX, int(nxt[p]) - p + 1)
        if v >= KMIN:
            c = 0
            for b in code[p:p + v]:
                c = c * 4 + int(b)
            out[offs[v - KMIN] + c] += 1
    return out.astype(np.int32)


def finalize(add):
    offs = np.cumsum([0] + [4 ** k for k in range(KMIN, KMAX + 1)])
    t = {k: add[offs[k - KMIN]:offs[k - KMIN + 1]].astype(np.int64) for k in range(KMIN, KMAX + 1)}
    for k in range(KMAX - 1, KMIN - 1, -1):
        t[k] = t[k] + t[k + 1].reshape(-1, 4).sum(1)
    canon = {}
    for k, tk in t.items():
        codes = np.arange(4 ** k)
        rc = np.zeros_like(codes)
        c = codes.copy()
        for _ in range(k):
            rc = (rc << 2) | ((c & 3) ^ 2)
            c >>= 2
        canon[k] = np.where(codes < rc, tk + tk[rc], np.where(codes == rc, tk, 0))
    return canon


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    td.init_process_group("gloo", rank=rank, world_size=world)
    n = 20011
    seq = synth.bases(55, 0, n, 0.35)
    seq[9990:10010] = ord("N")                     # an N run across the 2-rank cut
    rs = np.array([0, 17, 5000, 5001, 10005, 15000, n], dtype=np.uint64)
    for split_at_n in (True, False):
        sh = dist.shard_genome(seq, rs, rank, world, kmax=KMAX)
        add = torch.from_numpy(additive_tables(sh.seq, sh.rec_starts, sh.n_starts, split_at_n))
        dist.allreduce_tables(add)
        canon = finalize(add.numpy())
        if rank == 0:
            for k in range(KMIN, KMAX + 1):
                want = O.count_dense(seq, rs, k, mode_b=0 if split_at_n else 1).astype(np.int64)
                assert np.array_equal(canon[k], want), (split_at_n, k)
    # set-index sharding covers every set exactly once, in order, with no exchange
    S = 1003
    lo, hi = score.shard_sets(S, rank, world)
    mine = torch.zeros(S, dtype=torch.int32)
    mine[lo:hi] = 1
    td.all_reduce(mine)
    assert int(mine.min()) == 1 and int(mine.max()) == 1
    a, b = dist.shard_bounds(n, rank, world)
    cover = torch.zeros(n, dtype=torch.int32)
    cover[a:b] = 1
    td.all_reduce(cover)
    assert int(cover.min()) == 1 and int(cover.max()) == 1 and a % 32 == 0
    td.barrier()
    td.destroy_process_group()
    if rank == 0:
        print("dist worker ok")


if __name__ == "__main__":
    main()
