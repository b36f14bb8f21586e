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


class TableMerge(object):
    """The cross-GPU merge of the additive table blocks, as two halves so that other work (the replicated
    foreground count) can be enqueued in between:

        block()                     the device block this rank accumulates its background shard into
        start(d_fwd)                begin merging (asynchronous with respect to the caller's stream)
        finish(d_fwd, d_canon)      wait, then leave the FINALIZED forward tables in d_fwd and the canonical
                                    tables in d_canon on every rank

    This implementation: one NCCL all_reduce(SUM) of the block over NVLink, then finalize + fold on every rank.
    `events` (optional pair of CUDA events) brackets the time the caller's stream spends waiting for the merge."""

    name = "nccl all_reduce(sum, 89.5 MB) + finalize + fold on every rank"

    def __init__(self, words, kmin, kmax, device):
        self.words, self.kmin, self.kmax, self.device = words, kmin, kmax, device
        self.work = None

    def block(self):
        import torch
        return torch.empty(self.words, dtype=torch.int32, device=self.device)

    def start(self, d_fwd, events=None):
        import torch.distributed as td
        self.work = td.all_reduce(d_fwd, op=td.ReduceOp.SUM, async_op=True)

    def finish(self, d_fwd, d_canon, events=None):
        if events is not None:
            events[0].record()
        self.work.wait()
        self.work = None
        if events is not None:
            events[1].record()
        finalize_tables(d_fwd, self.kmin, self.kmax, d_canon)
