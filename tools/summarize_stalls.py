"""Warp-stall samples by SASS instruction for one kernel, from `ncu -i X.ncu-rep --page source --csv` (optionally gzipped).
   python tools/summarize_stalls.py gpurun_out/profiles_out/r02_gemm_step_source.csv.gz "gemm_pair_kernel<(int)0, (int)0, (int)4, (int)1>" r02"""
import collections, csv, gzip, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, kname, tag = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "r02"
op = gzip.open if path.endswith(".gz") else open
rows = list(csv.reader(op(path, "rt")))
agg, reasons, launches, hdr, on = collections.OrderedDict(), collections.Counter(), 0, None, False
for r in rows:
    if r and r[0] == "Kernel Name":
        on = kname in r[1]; launches += on; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if not on or not hdr or len(r) < len(hdr) - 2: continue
    d = dict(zip(hdr, r))
    n = int(d["# Samples"] or 0)
    if not n: continue
    key = d["Source"].strip()
    a = agg.setdefault((key, d["Address"][-5:]), [0, collections.Counter()])
    a[0] += n
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            a[1][k[6:]] += int(v); reasons[k[6:]] += int(v)
# the same instruction appears once per launch at a different load address: merge by text + address tail
tot = sum(v[0] for v in agg.values())
md = [f"# Dominant kernel: warp stall samples by instruction (`ncu --set full --import-source on`, round {tag})", "",
      f"`{kname}`: {launches} launches of one inference step (`tools/one_step.py`), {tot} samples.  Stall reasons over all samples: "
      + ", ".join(f"{k} {100*v/max(1,sum(reasons.values())):.0f}%" for k, v in reasons.most_common(8)) + ".", "",
      "| samples | share | SASS (address tail) | top stall reasons |", "|---|---|---|---|"]
for (src, ad), (n, rs) in sorted(agg.items(), key=lambda x: -x[1][0])[:28]:
    md.append(f"| {n} | {100*n/tot:.1f}% | `{src}` ({ad}) | " + ", ".join(f"{k} {v}" for k, v in rs.most_common(2)) + " |")
out = os.path.join(ROOT, "profiles", f"{tag}_gemm_step_stalls.md")
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md[:24]))
