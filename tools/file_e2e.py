"""Scratch: wall time of the file-based path (FASTA file -> canonical tables) and its parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swga_b200 import fasta, kmers, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
path = "/dev/shm/swga_e2e.fa" if os.path.isdir("/dev/shm") else "/tmp/swga_e2e.fa"
seq = synth.bases(4, 0, n, 0.41)
synth.write_fasta(path, seq, synth.equal_records(n, 4))
print("file bytes", os.path.getsize(path), "cpus", os.cpu_count(), flush=True)
torch.zeros(1).cuda()
for it in range(2):
    t0 = time.perf_counter(); raw = np.memmap(path, dtype=np.uint8, mode="r"); t1 = time.perf_counter()
    g = fasta.parse_fasta_bytes(raw, pin=True); t2 = time.perf_counter()
    tabs = kmers.count_all(g, 5, 12); torch.cuda.synchronize(); t3 = time.perf_counter()
    g2 = fasta.parse_fasta_bytes(raw, pin=False); t4 = time.perf_counter()
    tabs = kmers.count_all(g2, 5, 12); torch.cuda.synchronize(); t5 = time.perf_counter()
    t6 = time.perf_counter(); tabs2 = kmers.count_all(path, 5, 12); t7 = time.perf_counter()
    assert np.array_equal(tabs.block, tabs2.block)
    print("   file image -> tables (swga_count_fasta_host): %.3fs -> %.2f Gbases/s" % (t7 - t6, n / (t7 - t6) / 1e9), flush=True)
    tp = time.perf_counter(); x = torch.empty(n, dtype=torch.uint8, pin_memory=True); tq = time.perf_counter()
    print("   pageable (library staging): parse %.3fs count %.3fs ; pinned alloc alone %.3fs" % (t4 - t3, t5 - t4, tq - tp), flush=True); del x
    print("read %.3fs  parse %.3fs (%.2f GB/s)  count(host buffers) %.3fs  total %.3fs -> %.2f Gbases/s"
          % (t1 - t0, t2 - t1, len(raw) / (t2 - t1) / 1e9, t3 - t2, t3 - t0, n / (t3 - t0) / 1e9), flush=True)
os.remove(path)
