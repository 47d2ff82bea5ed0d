"""Turn the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/.
   python tools/summarize_profiles.py r01"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.environ.get("F2G_PROFILES_OUT", os.path.join(ROOT, "profiles")); os.makedirs(OUT, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

def launch_table(path):
    rows = list(csv.reader(open(path)))
    hdr, agg, n = None, collections.OrderedDict(), 0
    for r in rows:
        if len(r) > 10 and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            name = r[4].split("(")[0].replace("void ", "").replace("f2g::", "")
            name = name if len(name) < 70 else name[:67] + "..."
            a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]); n += 1
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1]/1e3:.1f} | {100*v[1]/tot:.1f}% | {v[1]/v[0]/1e3:.1f} |")
    return n, tot / 1e3, "\n".join(lines)

def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = {"Kernel Name": "kernel", "Grid Size": "grid", "gpu__time_duration.sum": "time_us",
            "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "hmma_pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
            "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct"}
    idx = {h: i for i, h in enumerate(hdr)}
    units = rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for k, nm in want.items():
            if k in idx:
                d[nm] = r[idx[k]]
                if nm in ("dram_read", "dram_write"): d[nm + "_unit"] = units[idx[k]]
        res.append(d)
    return res

md = [f"# ncu summaries, round {tag}", ""]
for name, title in (("launches_step.csv", "1-step inference (`tools/one_step.py`, eager, bench shape): every launch, `--metrics gpu__time_duration.sum --clock-control none`"),
                    ("launches_step_warm.csv", "the same step with `--cache-control none` (caches left warm between kernels: closer to the in-graph times)"),
                    ("launches_train.csv", "GAN D+G iteration pair (`tools/one_train_pair.py`, bs 16 x 24000), `--cache-control none`")):
    p = os.path.join(G, name)
    if os.path.exists(p):
        n, tot, tab = launch_table(p)
        md += [f"## {title}", "", f"{n} launches, {tot:.1f} us summed (serialised, cold-cache: compare shares)", "", tab, ""]
rep = os.path.join(G, "prof_gemm_step.ncu-rep")
if os.path.exists(rep):
    ms = raw_metrics(rep)
    md += ["## `ncu --set full` of the tcgen05 GEMM launches of one inference step", "",
           "| grid | time us | dram read | dram write | tensor pipe % (active) | dram % | L2 % | regs |", "|---|---|---|---|---|---|---|---|"]
    for d in ms:
        md.append(f"| {d.get('grid')} | {d.get('time_us')} | {d.get('dram_read')} {d.get('dram_read_unit','')} | {d.get('dram_write')} {d.get('dram_write_unit','')} | {d.get('tensor_pct_active')} | {d.get('dram_pct')} | {d.get('l2_pct')} | {d.get('regs')} |")
    json.dump(ms, open(os.path.join(OUT, f"{tag}_gemm_metrics.json"), "w"), indent=1)
    # bench.py's roofline.traffic: dram bytes per launch of the dominant kernel (fp16-operand
    # instantiation: last template argument 1; falls back to all pair-GEMM launches)
    def _bytes(v, unit):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        return float(v.replace(",", "")) * scale
    import re
    dom = [d for d in ms if re.search(r"gemm_pair_kernel<\d+, \d+, \d+, 1>", d.get("kernel", ""))] or ms
    tot = [_bytes(d["dram_read"], d.get("dram_read_unit", "byte")) + _bytes(d["dram_write"], d.get("dram_write_unit", "byte")) for d in dom]
    json.dump({"kernel": dom[0].get("kernel"), "launches": len(dom), "dram_bytes_per_launch_avg": sum(tot) / len(tot),
               "source": f"profiles/{tag}_gemm_metrics.json (ncu --set full --clock-control none, tools/one_step.py; "
                         "ncu flushes L2 before every kernel, so weights AND activations are re-read from HBM)"},
              open(os.path.join(OUT, "gemm_traffic.json"), "w"), indent=1)
open(os.path.join(OUT, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
print("\n".join(md)[:6000])
