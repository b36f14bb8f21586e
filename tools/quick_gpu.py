"""Scratch timing on the GPU box: stage timings of the counting path + atomic microbenchmarks."""
import ctypes, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swga_b200 import _lib, kmers, synth

lib, ctx = _lib.load(), _lib.context(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
seed, _, gc, n_rec = synth.CONFIGS["C2_bg"]
ta, tc, tg = synth.thresholds(gc)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
_lib.check(lib.swga_synth_bases(ctx.h, seed, 0, n, ta, tc, tg, _lib.ptr(d), None))
rs = synth.equal_records(n, n_rec)
drs = torch.from_numpy(rs.view(np.int64)).cuda()
words = int(lib.swga_tables_words(5, 12))
fwd = torch.empty(words, dtype=torch.int32, device="cuda")
canon = torch.empty(words, dtype=torch.int32, device="cuda")
packed = torch.empty(int(lib.swga_packed_words(n)), dtype=torch.int32, device="cuda")
brk = torch.empty(int(lib.swga_brk_words(n)), dtype=torch.int32, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
def step():
    ev[0].record(); fwd.zero_()
    ev[1].record(); _lib.check(lib.swga_pack(ctx.h, _lib.ptr(d), n, 0, _lib.ptr(packed), _lib.ptr(brk), None))
    _lib.check(lib.swga_mark_records(ctx.h, _lib.ptr(drs[1:]), n_rec - 1, n, _lib.ptr(brk), None))
    ev[2].record(); _lib.check(lib.swga_count_accumulate(ctx.h, _lib.ptr(packed), _lib.ptr(brk), n, 5, 12, _lib.ptr(fwd), None))
    ev[3].record(); _lib.check(lib.swga_count_finalize(ctx.h, 5, 12, _lib.ptr(fwd), None))
    ev[4].record(); _lib.check(lib.swga_fold_canonical(ctx.h, 5, 12, _lib.ptr(fwd), _lib.ptr(canon), None))
    ev[5].record()
cfgs = [(1, 0), (2, 0)] if len(sys.argv) < 3 else [(int(sys.argv[2]), 0)]
for impl, pm in cfgs:
  _lib.check(lib.swga_ctx_set_count_impl(ctx.h, impl));
  print("impl", impl)
  for i in range(3):
    step(); torch.cuda.synchronize()
    t = [ev[j].elapsed_time(ev[j + 1]) for j in range(5)]
    if i == 2: print("n=%d  zero %.3f  pack %.3f  count %.3f  finalize %.3f  fold %.3f  total %.3f ms  -> %.1f Gbases/s (count alone %.1f G atomics/s)"
          % (n, *t, sum(t), n / sum(t) / 1e6, n / t[2] / 1e6), flush=True)
  st = np.zeros(4, dtype=np.uint64)
  _lib.check(lib.swga_count_partition_stats(ctx.h, st.ctypes.data_as(_lib.c_u64p)))
  print("partition stats [staging-full, bucket-full, uniform groups, break groups]:", st.tolist(), flush=True)
  print("digest", int(canon.to(torch.int64).sum()), int((canon.to(torch.int64) * torch.arange(words, device="cuda") % 1000003).sum()), flush=True)
sys.exit(0)
ms = ctypes.c_float()
for tw in (1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 26):
    tab = torch.zeros(tw, dtype=torch.int32, device="cuda")
    _lib.check(lib.swga_bench_red_global(ctx.h, _lib.ptr(tab), tw, 1 << 31, ctypes.byref(ms)))
    print("red.global random, table %d words (%.1f MB): %.2f ms -> %.1f Gops/s" % (tw, tw * 4 / 1e6, ms.value, (1 << 31) / ms.value / 1e6), flush=True)
for tw in (1 << 10, 1 << 12, 1 << 14, 32768):
    _lib.check(lib.swga_bench_atom_shared(ctx.h, tw, 1 << 32, ctypes.byref(ms)))
    print("atom.shared random, table %d words: %.2f ms -> %.1f Gops/s" % (tw, ms.value, (1 << 32) / ms.value / 1e6), flush=True)
