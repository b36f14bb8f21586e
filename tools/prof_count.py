"""One background-sized counting pass for ncu: pack + mark + swga_count_accumulate + finalize + fold of the C2
background (or its first N bases), twice (the first pass allocates the workspaces).  `python tools/prof_count.py [N]`."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from swga_b200 import _lib, kmers, synth

lib, ctx = _lib.load(), _lib.context(0)
if os.environ.get("SWGA_VIA_PACK"):
    _lib.check(lib.swga_ctx_set_count_via_pack(ctx.h, 1))
seed, n, gc, n_rec = synth.CONFIGS["C2_bg"]
if len(sys.argv) > 1:
    n = int(float(sys.argv[1]))
ta, tc, tg = synth.thresholds(gc)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
_lib.check(lib.swga_synth_bases(ctx.h, seed, 0, n, ta, tc, tg, _lib.ptr(d), None))
rs = synth.equal_records(n, n_rec)
for _ in range(2):
    fwd, canon = kmers.count_all_device(d, rs, mode_b=True)
    torch.cuda.synchronize()
print("digest", int(canon.to(torch.int64).sum()))
