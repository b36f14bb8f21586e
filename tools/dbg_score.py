import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from swga_b200 import fasta, locate, score, synth
n = 1_200_000
g = fasta.Genome(synth.bases(31, 0, n, 0.3), np.array([0, 400_000, n], dtype=np.uint64), ["a", "b"])
gi = locate.GenomeIndex(g)
primers = ["AAAAT", "ATATAT", "AAATTTAA", "ACGTACGT", "AATT"]
sl = score.SiteLists.from_index(gi, primers)
torch.cuda.synchronize(); print("sitelists ok", sl.counts(), flush=True)
sets = np.array([[0, 1, 2, -1], [3, 4, -1, -1], [0, 1, 2, 3]], dtype=np.int32)
res = score.score_sets(sl, sets, np.array([10.0, 20.0, 30.0]), 36000)
print({k: v for k, v in res.items()})
