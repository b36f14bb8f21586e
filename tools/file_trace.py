"""Scratch: stage times of swga_count_fasta_host on a synthetic FASTA file (SWGA_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SWGA_TRACE"] = "1"
import numpy as np, torch
from swga_b200 import kmers, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
path = "/dev/shm/swga_e2e.fa" if os.path.isdir("/dev/shm") else "/tmp/swga_e2e.fa"
synth.write_fasta(path, synth.bases(4, 0, n, 0.41), synth.equal_records(n, 4))
torch.zeros(1).cuda()
for it in range(3):
    t0 = time.perf_counter(); tabs = kmers.count_all(path, 5, 12); t1 = time.perf_counter()
    print("file -> tables %.3fs (%.2f Gbases/s)" % (t1 - t0, n / (t1 - t0) / 1e9), flush=True)
os.remove(path)
