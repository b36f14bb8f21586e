"""Samples, instructions and shared-memory wavefronts of a kernel per PHASE (the code between two CTA barriers),
from an ncu source-page CSV (`ncu -i x.ncu-rep --page source --csv --print-source sass`).  `python tools/ncu_phases.py
file.csv [n_lines_of_one_launch]`; also lists the hottest lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
if len(sys.argv) > 2:
    body = body[:int(sys.argv[2])]


def I(r, k):
    try:
        return int(r[ix[k]] or 0)
    except (ValueError, KeyError):
        return 0


tot = sum(I(r, "# Samples") for r in body)
toti = sum(I(r, "Instructions Executed") for r in body)
print("lines", len(body), "samples", tot, "warp-instructions", toti)
bars = [n for n, r in enumerate(body) if "BAR." in r[ix["Source"]]]
edges = [0] + bars + [len(body)]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for a, b in zip(edges[:-1], edges[1:]):
    seg = body[a:b]
    s = sum(I(r, "# Samples") for r in seg)
    if not s:
        continue
    ins = sum(I(r, "Instructions Executed") for r in seg)
    wf = sum(I(r, "L1 Wavefronts Shared") for r in seg)
    st = sorted(((sum(I(r, h) for r in seg), h[6:]) for h in stall_cols), reverse=True)[:4]
    print("lines %4d-%4d  samples %5.1f%%  instructions %5.1f%%  shared wavefronts %11d  stalls %s" % (
        a, b, 100.0 * s / tot, 100.0 * ins / max(toti, 1), wf, [(k, round(100.0 * v / s)) for v, k in st]))
print("hottest lines:")
for n, r in sorted(enumerate(body), key=lambda x: -I(x[1], "# Samples"))[:30]:
    s = I(r, "# Samples")
    st = sorted(((I(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
    print("%4d %6d %5.1f%% inst=%10s wf=%-10s %-70s %s" % (n, s, 100.0 * s / tot, r[ix["Instructions Executed"]],
                                                          r[ix["L1 Wavefronts Shared"]], r[ix["Source"]].strip()[:70], st))
