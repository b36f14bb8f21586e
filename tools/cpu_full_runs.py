#!/usr/bin/env python
"""the reference's counting path -- `dsk <fasta> k -o ... -t 1` for k = 5..12, foreground then
background, ONE AFTER ANOTHER on one core as swga/commands/count.py:84-91 runs them, each followed by the
(restated) Python record parser swga/kmers.py:94-115 -- on the FULL configs C1 (1 Mbp / 10 Mbp) and C5
(1.3 Mbp / 140 Mbp).  Too long for bench.py's default run (tens of minutes), so it is run once and its output
committed as profiles/r02_cpu_full_runs.json; bench.py embeds it as cpu_baseline.full_runs.

    python tools/cpu_full_runs.py [C1 C5] > profiles/r02_cpu_full_runs.json
"""
import json
import os
import platform
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import swga_oracle as O          # noqa: E402  (checker / CPU baseline only)
from swga_b200 import synth                  # noqa: E402


def main():
    configs = [a for a in sys.argv[1:] if not a.startswith("-")] or ["C1", "C5"]
    out = {"what": "dsk k=5..12 sequential (1 core) + Python record parser, full genomes",
           "host": platform.node(), "cpus": os.cpu_count(), "runs": {}}
    for cfg in configs:
        rows, total = [], 0.0
        with tempfile.TemporaryDirectory(dir=os.environ.get("SWGA_TMP")) as tmp:
            for which in ("fg", "bg"):
                seed, n, gc, n_rec = synth.CONFIGS["%s_%s" % (cfg, which)]
                fa = os.path.join(tmp, "%s_%s.fa" % (cfg, which))
                synth.write_fasta(fa, synth.bases(seed, 0, n, gc), synth.equal_records(n, n_rec))
                for k in range(5, 13):
                    t0 = time.perf_counter()
                    res = O.run_dsk(fa, k, tmp)
                    t1 = time.perf_counter()
                    distinct = sum(1 for _ in O.parse_kmer_binary(res))
                    t2 = time.perf_counter()
                    os.remove(res)
                    rows.append({"genome": which, "bases": n, "k": k, "dsk_s": round(t1 - t0, 2),
                                 "parse_s": round(t2 - t1, 2), "distinct": distinct})
                    total += t2 - t0
                    sys.stderr.write("%s %s k=%d dsk %.1f s parse %.1f s\n" % (cfg, which, k, t1 - t0, t2 - t1))
                os.remove(fa)
        bases = synth.CONFIGS[cfg + "_fg"][1] + synth.CONFIGS[cfg + "_bg"][1]
        out["runs"][cfg] = {"total_s": round(total, 1), "bases": bases, "gbases_per_s": bases / total / 1e9,
                            "per_run": rows}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
