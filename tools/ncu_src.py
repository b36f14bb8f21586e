"""Print the hottest SASS lines of an ncu source-page CSV (--page source --csv --print-source sass)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
toti = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print("total samples", tot, "warp-instr", toti)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > tot * 0.01})
mode = sys.argv[2] if len(sys.argv) > 2 else "top"
if mode == "all":
    for n, r in enumerate(body):
        s = int(r[ix["# Samples"]] or 0)
        top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print("%4d %6d %5.1f%% %9s wf=%-9s %-70s %s" % (n, s, 100.0 * s / tot, r[ix["Instructions Executed"]], r[ix["L1 Wavefronts Shared"]], r[ix["Source"]].strip()[:70], top if s > tot * 0.003 else ""))
else:
    for n, r in sorted(enumerate(body), key=lambda x: -int(x[1][ix["# Samples"]] or 0))[:40]:
        s = int(r[ix["# Samples"]] or 0)
        top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print("%4d %6d %5.1f%% %9s wf=%-9s %-70s %s" % (n, s, 100.0 * s / tot, r[ix["Instructions Executed"]], r[ix["L1 Wavefronts Shared"]], r[ix["Source"]].strip()[:70], top))
