"""Multi-GPU decomposition of the hot path (one process per GPU, torch.distributed).

Counting shards the linearised genome into contiguous chunks of START positions; each rank also
receives the kmax-1 bases after its last start (right halo), so a k-mer is owned by the rank that
owns its first base.  Record boundaries inside the chunk (and its halo) travel as chunk-local record
starts.  Per-rank tables are in the additive "top-k + tails" form (include/swga_b200.h), merged by
ONE allreduce(sum) over NCCL/NVLink, then finalised identically on every rank.

Set scoring shards by set index with no collective (see swga_b200/score.py).
"""
import numpy as np

from . import _lib


class Shard(object):
    def __init__(self, seq, rec_starts, n_starts, start):
        self.seq, self.rec_starts, self.n_starts, self.start = seq, rec_starts, n_starts, start


def shard_bounds(n, rank, world, align=32):
    """Owned start positions [a, b) of `rank`; cuts are multiples of `align` (16-byte aligned loads)."""
    per = -(-n // world)
    per = -(-per // align) * align
    a = min(n, rank * per)
    b = min(n, a + per)
    return a, b


def shard_genome(seq, rec_starts, rank, world, kmax):
    """Slice (views, no copy) of the genome this rank packs and counts."""
    n = len(seq)
    a, b = shard_bounds(n, rank, world)
    e = min(n, b + kmax - 1)
    rs = np.asarray(rec_starts, dtype=np.uint64)
    inner = rs[(rs > a) & (rs < e)] - np.uint64(a)             # record starts strictly inside the slice
    local = np.concatenate([[0], inner, [e - a]]).astype(np.uint64)
    return Shard(seq[a:e], local, b - a, a)


def finalize_tables(d_fwd, kmin, kmax, d_canon=None):
    """swga_count_finalize + swga_fold_canonical on an (allreduced) additive table block."""
    import torch
    lib = _lib.load()
    ctx = _lib.context(d_fwd.device.index)
    if d_canon is None:
        d_canon = torch.empty_like(d_fwd)
    stream = torch.cuda.current_stream(d_fwd.device).cuda_stream
    _lib.check(lib.swga_count_finalize(ctx.h, kmin, kmax, _lib.ptr(d_fwd), stream))
    _lib.check(lib.swga_fold_canonical(ctx.h, kmin, kmax, _lib.ptr(d_fwd), _lib.ptr(d_canon), stream))
    return d_canon


def allreduce_tables(d_fwd):
    """Sum the additive table blocks of all ranks in place (uint32 counts carried in int32 lanes:
    two's-complement addition is the same operation).  No-op without an initialised process group."""
    import torch.distributed as td
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        td.all_reduce(d_fwd, op=td.ReduceOp.SUM)
    return d_fwd
