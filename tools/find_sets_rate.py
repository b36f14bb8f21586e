"""Scratch: candidate sets per second through pipeline.find_sets (host set_finder -> scoring), block path vs
the per-line path, on the BASELINE config-5 workspace."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swga_b200 import fasta, locate, pipeline, sets, synth

class Stop(Exception):
    pass

genomes = {}
for name in ("C5_fg", "C5_bg"):
    seed, n, gc, n_rec = synth.CONFIGS[name]
    genomes[name] = fasta.Genome(synth.bases(seed, 0, n, gc), synth.equal_records(n, n_rec), ["%s%d" % (name, i) for i in range(n_rec)])
fg, bg = genomes["C5_fg"], genomes["C5_bg"]
rows = pipeline.count_primers(fg, bg, None, min_size=5, max_size=12, min_fg_bind=13, max_bg_bind=938, max_dimer_bp=3)
primers = [pipeline.Primer(r["seq"], r["fg_freq"], r["bg_freq"], r["ratio"]) for r in rows]
gi = locate.GenomeIndex(fg)
active = pipeline.filter_primers(primers, gi, 13, 938, max_primers=200, max_gini=0.6)
print("active primers", len(active), flush=True)
graph = os.path.join(tempfile.mkdtemp(), "g.dimacs")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
for mode in ("blocks", "lines"):
    t0 = time.perf_counter()
    last = [0, 0]
    def status(processed, passed, smallest):
        last[0], last[1] = processed, passed
        if time.perf_counter() - t0 > budget:
            raise Stop()
    try:
        if mode == "blocks":
            pipeline.find_sets(active, gi, int(bg.n_bases), 30000, 36000, 2, 7, max_sets=-1, graph_fp=graph, on_status=status)
        else:
            pipeline.build_graph(pipeline.assign_ids(active), 3, graph)
            gen = sets.find(min_bg_bind_dist=30000, min_size=2, max_size=7, bg_length=int(bg.n_bases), graph_fp=graph)
            try:
                pipeline.find_sets(active, gi, int(bg.n_bases), 30000, 36000, 2, 7, max_sets=-1, graph_fp=graph, lines=gen, on_status=status)
            finally:
                gen.close()
    except Stop:
        pass
    dt = time.perf_counter() - t0
    print("%s: %d candidate sets in %.2f s -> %.0f sets/s (%d passed)" % (mode, last[0], dt, last[0] / dt, last[1]), flush=True)
