"""SASS opcode census of the built library, per object file (run here, no GPU needed):

    python tools/sass_census.py [tag]          -> profiles/<tag>_sass_census.md

Counts the mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
SHFL = warp shuffles (the FFT butterflies), HMMA = legacy mma.sync (must be 0), plus per-kernel
registers / spills from the -Xptxas -v build log."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "flow2gan_b200", "csrc")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "LDTM", "STTM", "SYNCS", "SHFL",
       "HMMA", "LDGSTS", "ATOMG", "REDG", "RED", "MEMBAR", "FFMA", "MUFU"]
rows = []
for fn in sorted(os.listdir(CSRC)):
    if not fn.endswith(".o"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, fn)], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    kernels = 0
    variants = collections.Counter()
    for ln in sass.splitlines():
        if "Function :" in ln:
            kernels += 1
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        if base in OPS:
            cnt[base] += 1
        if base in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "LDGSTS", "SHFL"):
            variants[op] += 1
    rows.append((fn, kernels, cnt, variants))
md = [f"# SASS opcode census, round {tag}", "",
      "`cuobjdump -sass flow2gan_b200/csrc/*.o` (sm_100a), instruction counts per object file.  "
      "`UTCHMMA` = `tcgen05.mma` (kind::f16 / kind::tf32), `LDTM` = `tcgen05.ld`, `UTMALDG` = TMA tensor load, "
      "`UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier ops, `SHFL` = warp shuffles, `HMMA` = legacy `mma.sync` (none).", "",
      "| object | kernels | " + " | ".join(OPS) + " |", "|---|---|" + "---|" * len(OPS)]
for fn, k, cnt, _ in rows:
    md.append(f"| `{fn}` | {k} | " + " | ".join(str(cnt.get(o, 0)) for o in OPS) + " |")
md += ["", "## Variants of the Blackwell-specific opcodes", ""]
for fn, k, cnt, var in rows:
    if var:
        md.append(f"* `{fn}`: " + ", ".join(f"`{o}` x{n}" for o, n in sorted(var.items(), key=lambda x: -x[1])[:14]))
log = os.path.join(CSRC, "build.log")
if os.path.exists(log):
    md += ["", "## Registers / spills per kernel (`-Xptxas -v`, kernels using > 64 registers or any spill)", "",
           "| kernel | registers | spill stores B | spill loads B | smem B |", "|---|---|---|---|---|"]
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n[^\n]*Used (\d+) registers(?:, used \d+ barriers)?(?:, \d+ bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        regs, ss, sl = int(m.group(5)), int(m.group(3)), int(m.group(4))
        if regs > 64 or ss or sl:
            md.append(f"| `{name[:90]}` | {regs} | {ss} | {sl} | {m.group(6) or 0} |")
out = os.path.join(ROOT, "profiles", f"{tag}_sass_census.md")
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md[:40]))
