"""Aggregate an ncu --metrics gpu__time_duration.sum CSV launch list.  python tools/launch_table.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; seq = []
for r in rows:
    if len(r) > 10 and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        name = r[4].split("(")[0].replace("void ", "").replace("f2g::", "")
        seq.append((name[:44], r[hdr.index("Grid Size")], float(r[-1]) / 1e3))
agg = collections.OrderedDict()
for n, g, t in seq:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print("launches", len(seq), "total us %.1f" % tot)
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k:46s} {v[0]:4d} {v[1]:9.1f} {v[1]/v[0]:8.1f} {100*v[1]/tot:5.1f}%")
