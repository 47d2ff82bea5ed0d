"""One-line digest of gpurun_out/bench_quick.json (tools/gpu_quick.sh)."""
import json
for l in open("gpurun_out/bench_quick.json"):
    if l.startswith("{"):
        d = json.loads(l); t = d.get("gan_train") or {}; r = t.get("roofline") or {}
        print("ms/step %.4f e2e %.1fM gemm frac %.3f (%.1f us/launch) hbm %.3f | train %.2f ms/pair gemm_ms %s busy %s glue %s launches %s" % (
            d["ms_per_step"], d["e2e"]["value"] / 1e6, d["roofline"]["frac"], d["roofline"]["us_per_launch_avg"],
            (d.get("roofline_hbm") or {}).get("frac", 0), t.get("ms_per_pair", 0), r.get("gemm_ms"), r.get("gpu_busy_ms"),
            r.get("torch_glue_share"), r.get("launches_per_pair")))
