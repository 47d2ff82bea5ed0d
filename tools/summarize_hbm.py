"""ncu --set full rows of the HBM-bound kernels -> profiles/<tag>_hbm_kernels.md (+ .json).

    python tools/summarize_hbm.py <tag> gpurun_out/prof_hbm_step.ncu-rep [more .ncu-rep ...]

Per launch: duration, dram__bytes_read/write.sum, achieved DRAM GB/s (= those bytes / duration) against the
measured copy bandwidth (MEASURED_PEAKS.json), L2 throughput %, achieved occupancy (warps active %), registers,
and the top three warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, reps = sys.argv[1], sys.argv[2:]
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
hbm = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
def val(r, hdr, units, k, default=0.0):
    if k not in hdr: return default
    i = hdr.index(k)
    try: return float(r[i].replace(",", "")) * SC.get(units[i], 1)
    except ValueError: return default
out = []
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("f2g::", "")
        us = val(r, hdr, units, "gpu__time_duration.sum")
        rd, wr = val(r, hdr, units, "dram__bytes_read.sum"), val(r, hdr, units, "dram__bytes_write.sum")
        stalls = sorted(((val(r, hdr, units, c), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                         for c in stall_cols), reverse=True)
        stalls = [s for s in stalls if s[1] not in ("selected",)][:3]
        out.append({"kernel": name, "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")], "us": us,
                    "dram_read_mb": rd / 1e6, "dram_write_mb": wr / 1e6, "dram_gbs": (rd + wr) / us / 1e3 if us else 0.0,
                    "frac_of_hbm_peak": (rd + wr) / us / 1e3 / hbm if us else 0.0,
                    "dram_pct": val(r, hdr, units, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    "l2_pct": val(r, hdr, units, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "warps_active_pct": val(r, hdr, units, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "regs": int(val(r, hdr, units, "launch__registers_per_thread")),
                    "top_stalls": [f"{n} {v:.1f}" for v, n in stalls], "source": os.path.basename(rep)})
md = [f"# HBM-bound kernels, `ncu --set full --clock-control none`, round {tag}", "",
      f"Achieved GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum; peak = {hbm:.0f} GB/s "
      "(MEASURED_PEAKS.json, copy).  ncu flushes the caches before every kernel replay, so inputs that the step keeps "
      "L2-resident are read from DRAM here (cold figures; the in-step, warm-L2 times are in the bench line).  Stall columns: "
      "average warps stalled per issue-active cycle, top three reasons.", "",
      "| kernel | grid x block | us | dram read MB | dram write MB | GB/s | of peak | L2 % | warps active % | regs | top stalls |",
      "|---|---|---|---|---|---|---|---|---|---|---|"]
for d in out:
    md.append(f"| `{d['kernel'][:48]}` | {d['grid']} x {d['block']} | {d['us']:.1f} | {d['dram_read_mb']:.2f} | {d['dram_write_mb']:.2f} | "
              f"{d['dram_gbs']:.0f} | {d['frac_of_hbm_peak']:.2f} | {d['l2_pct']:.0f} | {d['warps_active_pct']:.0f} | {d['regs']} | {'; '.join(d['top_stalls'])} |")
OUTD = os.environ.get("F2G_PROFILES_OUT", os.path.join(ROOT, "profiles"))
os.makedirs(OUTD, exist_ok=True)
open(os.path.join(OUTD, f"{tag}_hbm_kernels.md"), "w").write("\n".join(md) + "\n")
json.dump(out, open(os.path.join(OUTD, f"{tag}_hbm_kernels.json"), "w"), indent=1)
print("\n".join(md))
