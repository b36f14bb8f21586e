"""Summarise an `ncu --csv --metrics ...` log: per kernel, last launch's metrics (or mean)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault((d["Kernel Name"].split("(")[0][:28], d["Grid Size"]), collections.OrderedDict()).setdefault(d["Metric Name"], []).append(v)
for (k, g), m in agg.items():
    print(k, g)
    for name, v in m.items():
        print("    %-58s n=%-3d last=%-16.6g mean=%.6g" % (name, len(v), v[-1], sum(v) / len(v)))
