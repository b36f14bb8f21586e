"""Worker for tests/test_dist_gpu.py: one rank of a world_size-N job with the REAL kernels.  Ranks share
cuda:0 when the box has fewer GPUs than ranks (then the collective is gloo on CUDA tensors; with enough GPUs it
is NCCL): chunk + halo sharding, swga_count_accumulate per shard (partitioned kernel), allreduce of the additive
tables, finalize + fold on every rank, oracle comparison on rank 0; set scoring sharded by set index."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import swga_oracle as O                         # noqa: E402
from swga_b200 import dist, fasta, kmers, locate, score, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    n_dev = torch.cuda.device_count()
    dev = rank % n_dev
    torch.cuda.set_device(dev)
    backend = "nccl" if n_dev >= world else "gloo"
    td.init_process_group(backend, rank=rank, world_size=world)
    # ---- counting: 9.4 Mbp, shards above the partitioned kernel's threshold, an N run across the 2-rank cut ----
    n = 9_400_011
    seq = synth.bases(91, 0, n, 0.38)
    cut = dist.shard_bounds(n, 1, 2)[0]
    seq[cut - 5:cut + 9] = ord("N")
    seq[1_000_000:1_000_700] = ord("A")
    rs = np.array([0, 31, 2_000_000, cut + 3, 7_000_001, n], dtype=np.uint64)
    from swga_b200 import _lib
    _lib.check(_lib.load().swga_ctx_set_count_impl(_lib.context(dev).h, 2))      # 4.7 Mbp shards: force the partitioned kernel
    for mode_b in (False, True):
        sh = dist.shard_genome(seq, rs, rank, world, kmax=12)
        d = torch.from_numpy(np.ascontiguousarray(sh.seq)).cuda()
        fwd, _ = kmers.count_all_device(d, sh.rec_starts, mode_b=mode_b, n_starts=sh.n_starts, finalize=False)
        dist.allreduce_tables(fwd)
        canon = dist.finalize_tables(fwd, 5, 12)
        torch.cuda.synchronize()
        if rank == 0:
            want = O.count_dense_allk(seq, rs, 5, 12, mode_b=int(mode_b))
            assert np.array_equal(canon.cpu().numpy().view(np.uint32), want), mode_b
    # ---- the same merge through dist.table_merge: on a box with one GPU per rank this is the fused peer-memory
    # kernel (csrc/merge.cu: reduce-scatter + finalize + all-gather over NVLink mappings), else the NCCL / gloo path ----
    lib = _lib.load()
    words = int(lib.swga_tables_words(5, 12))
    names = []
    for prefer in (("multicast", "peer") if backend == "nccl" else ("nccl",)):
        merge = dist.table_merge(words, 5, 12, torch.device("cuda", dev), prefer=prefer)
        names.append(merge.name)
        if backend == "nccl":
            assert isinstance(merge, dist.PeerTableMerge), merge.name
            if prefer == "peer":
                assert not merge.multicast
        for it in range(3):                                        # repeated: the block is re-zeroed and re-merged in place
            sh = dist.shard_genome(seq, rs, rank, world, kmax=12)
            d = torch.from_numpy(np.ascontiguousarray(sh.seq)).cuda()
            fwd = merge.block()
            fwd.zero_()
            kmers.count_all_device(d, sh.rec_starts, mode_b=True, n_starts=sh.n_starts, finalize=False, d_fwd=fwd)
            canon = torch.empty_like(fwd)
            merge.start(fwd)
            merge.finish(fwd, canon)
            torch.cuda.synchronize()
            want = O.count_dense_allk(seq, rs, 5, 12, mode_b=1)
            assert np.array_equal(canon.cpu().numpy().view(np.uint32), want), (rank, it, merge.name)
            td.barrier()
    _lib.check(_lib.load().swga_ctx_set_count_impl(_lib.context(dev).h, 0))
    # ---- scoring: every rank scores its contiguous share of the sets; together they equal the unsharded result ----
    g = fasta.Genome(seq[:3_000_000].copy(), np.array([0, 1_200_000, 3_000_000], dtype=np.uint64), ["a", "b"])
    gi = locate.GenomeIndex(g)
    rng = np.random.RandomState(7)
    primers = []
    while len(primers) < 24:
        k = int(rng.randint(6, 11))
        p = int(rng.randint(0, g.n_bases - k))
        w = g.seq[p:p + k].tobytes().decode()
        if set(w) <= set("ACGT") and w not in primers:
            primers.append(w)
    sl = score.SiteLists.from_index(gi, primers)
    S = 501
    sets_ = np.full((S, 6), -1, dtype=np.int32)
    for i in range(S):
        m = int(rng.randint(1, 7))
        sets_[i, :m] = rng.choice(len(primers), m, replace=False)
    bgs = rng.uniform(10, 1e4, S)
    lo, hi = score.shard_sets(S, rank, world)
    mine = score.score_sets(sl, sets_[lo:hi], bgs[lo:hi], 36000)
    part = torch.zeros(S, 3, dtype=torch.float64, device="cuda")
    part[lo:hi, 0] = torch.from_numpy(mine.fg_max_dist.astype(np.float64)).cuda()
    part[lo:hi, 1] = torch.from_numpy(np.nan_to_num(mine.fg_dist_gini, nan=-7.0)).cuda()
    part[lo:hi, 2] = torch.from_numpy(mine.passed.astype(np.float64)).cuda()
    td.all_reduce(part)
    if rank == 0:
        full = score.score_sets(sl, sets_, bgs, 36000)
        got = part.cpu().numpy()
        assert np.array_equal(got[:, 0], full.fg_max_dist.astype(np.float64))
        assert np.array_equal(got[:, 1], np.nan_to_num(full.fg_dist_gini, nan=-7.0))
        assert np.array_equal(got[:, 2], full.passed.astype(np.float64))
    td.barrier()
    td.destroy_process_group()
    if rank == 0:
        print("dist gpu worker ok (%s; merges: %s)" % (backend, " | ".join(names)))


if __name__ == "__main__":
    main()
