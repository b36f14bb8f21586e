"""Multi-GPU decomposition of the hot path (one process per GPU, torch.distributed).

Counting shards the linearised genome into contiguous chunks of START positions; each rank also
receives the kmax-1 bases after its last start (right halo), so a k-mer is owned by the rank that
owns its first base.  Record boundaries inside the chunk (and its halo) travel as chunk-local record
starts.  Per-rank tables are in the additive "top-k + tails" form (include/swga_b200.h), merged by ONE kernel per
rank over NVLink peer memory that also finalises them (PeerTableMerge, csrc/merge.cu) -- or, where symmetric memory
is not available, by one NCCL allreduce(sum) followed by the finalize pass on every rank (TableMerge).

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


class PeerTableMerge(TableMerge):
    """The merge as ONE kernel per rank over NVLink peer mappings (csrc/merge.cu, swga_merge_finalize): reduce-scatter
    by code range, finalize of the owned slice and all-gather in one pass over the blocks of all ranks, which live in
    a torch symmetric-memory allocation (torch supplies the allocation, the peer pointers and the device-side
    barrier; the data path is the kernel).  The kernel runs on a side stream between two barriers across the ranks,
    so the caller's stream is free for the replicated foreground count; finish() then folds locally."""

    name = "k_merge_finalize: fused reduce-scatter + finalize + all-gather over NVLink peer mappings; fold on every rank"

    def __init__(self, words, kmin, kmax, device, use_multicast=True):
        import ctypes
        import torch
        import torch.distributed as td
        import torch.distributed._symmetric_memory as symm
        TableMerge.__init__(self, words, kmin, kmax, device)
        self.lib = _lib.load()
        self.ctx = _lib.context(device.index)
        self.world, self.rank = td.get_world_size(), td.get_rank()
        ok = ctypes.c_int(0)
        _lib.check(self.lib.swga_merge_supported(kmin, kmax, self.world, ctypes.byref(ok)))
        if not ok.value:
            raise NotImplementedError("k_merge_finalize covers kmax - kmin <= 7 on 2, 4 or 8 ranks")
        try:
            symm.enable_symm_mem_for_group(td.group.WORLD.group_name)
        except Exception:
            pass
        self.t = symm.empty(words, dtype=torch.int32, device=device)
        self.hdl = symm.rendezvous(self.t, td.group.WORLD)
        self.ptrs = np.array([int(p) for p in self.hdl.buffer_ptrs], dtype=np.uint64)
        assert len(self.ptrs) == self.world and int(self.ptrs[self.rank]) == self.t.data_ptr()
        # NVSwitch multicast mapping of the blocks (0 when the fabric has none): the in-switch reduction variant
        self.multicast = 0
        if use_multicast:
            try:
                self.multicast = int(self.hdl.multicast_ptr or 0)
            except Exception:
                self.multicast = 0
        flag = torch.tensor([1 if self.multicast else 0], dtype=torch.int32, device=device)
        td.all_reduce(flag, op=td.ReduceOp.MIN)                 # every rank or none
        if int(flag.item()) == 0:
            self.multicast = 0
        if self.multicast:
            self.name = ("k_merge_finalize_mc: reduce-scatter INSIDE the NVSwitch (multimem.ld_reduce) + finalize + all-gather "
                         "(multimem.st) in one kernel; fold on every rank")
        self.side = torch.cuda.Stream(device=device, priority=-1)
        self.ev_in, self.ev_out = torch.cuda.Event(), torch.cuda.Event()
        self.time_kernel = False                               # bench.py: CUDA events around the kernel, on its stream
        self.kernel_events = []

    def block(self):
        return self.t

    def start(self, d_fwd, events=None):
        import torch
        assert d_fwd.data_ptr() == self.t.data_ptr(), "accumulate into TableMerge.block()"
        cur = torch.cuda.current_stream(self.device)
        self.ev_in.record(cur)
        self.side.wait_event(self.ev_in)
        with torch.cuda.stream(self.side):
            self.hdl.barrier(channel=0)                        # every rank has finished accumulating
            if self.time_kernel:
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record(self.side)
            if self.multicast:
                _lib.check(self.lib.swga_merge_finalize_multicast(self.ctx.h, _lib.ptr(self.multicast), self.world, self.rank,
                                                                  self.kmin, self.kmax, self.side.cuda_stream))
            else:
                _lib.check(self.lib.swga_merge_finalize(self.ctx.h, _lib.ptr(self.ptrs), self.world, self.rank, self.kmin,
                                                        self.kmax, self.side.cuda_stream))
            if self.time_kernel:
                k1.record(self.side)
                self.kernel_events.append((k0, k1))
            self.hdl.barrier(channel=1)                        # every rank has written its slice everywhere
            self.ev_out.record(self.side)

    def finish(self, d_fwd, d_canon, events=None):
        import torch
        cur = torch.cuda.current_stream(self.device)
        if events is not None:
            events[0].record()
        cur.wait_event(self.ev_out)
        if events is not None:
            events[1].record()
        _lib.check(self.lib.swga_fold_canonical(self.ctx.h, self.kmin, self.kmax, _lib.ptr(d_fwd), _lib.ptr(d_canon),
                                                cur.cuda_stream))


def table_merge(words, kmin, kmax, device, prefer="peer"):
    """The best merge this job supports: the fused peer-memory kernel, else the NCCL all_reduce.  The choice is made
    TOGETHER: a rank whose symmetric-memory set-up fails must not leave the others waiting in a barrier, so every rank
    reports its outcome and all fall back if any did."""
    import torch
    import torch.distributed as td
    if prefer in ("peer", "multicast"):
        m, err = None, ""
        try:
            m = PeerTableMerge(words, kmin, kmax, device, use_multicast=(prefer == "multicast"))
        except Exception as e:                                  # no symmetric memory / no P2P on this box
            err = repr(e)[:200]
        ok = torch.tensor([1 if m is not None else 0], dtype=torch.int32, device=device)
        td.all_reduce(ok, op=td.ReduceOp.MIN)
        if int(ok.item()) == 1:
            return m
        m = TableMerge(words, kmin, kmax, device)
        m.name += " (peer merge unavailable%s)" % (": " + err if err else " on another rank")
        return m
    return TableMerge(words, kmin, kmax, device)
