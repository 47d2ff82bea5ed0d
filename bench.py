#!/usr/bin/env python
"""Headline benchmark of the Flow2GAN hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-timesteps 1]

A "step" = one `model.infer` of mel_24k_base on a synthetic (16, 100, 94) mel batch (bs=16 x
~1 s of 24 kHz audio, 16 x 24064 samples), n_timesteps ODE steps (default 1 = the metric's
configuration).  Prints ONE JSON line (rank 0).

  value      : audio samples/s, inputs resident in HBM, K CUDA-graph replays timed with CUDA
               events (max over ranks; N>1 = N independent replicas, weak scaling).
  e2e        : the same through the public API with HOST buffers: pinned mel -> H2D ->
               model.infer (draws its noise like the reference) -> D2H pinned audio, per step.
  roofline   : the dominant kernel = the fp16-operand tcgen05 GEMM of the ConvNeXt blocks
               (pwconv1/pwconv2): sum(2*M*N*K) of its launches in one step / their time when
               replayed back to back from a CUDA graph (CUDA events), against the measured dense
               bf16/fp16 tensor peak; `roofline_other` = the remaining TF32 GEMM launches;
               `roofline_hbm` = the block-prologue launches (dominant HBM-bound kernel) the same way;
               `traffic` = dram bytes per launch from the committed ncu capture (profiles/).
  cpu_baseline / --impl reference : the UNMODIFIED reference (staged by tools/stage_reference.sh under
               the git-ignored baseline/_ref/, which ships to the GPU box) through its own public API
               get_model(checkpoint=...) -> model.infer on the host cores, bounded sample; falls back to
               the oracle port (kind "port") only when baseline/_ref is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODEL = "mel_24k_base"
B, N_MELS, FRAMES, HOP = 16, 100, 94, 256
T = FRAMES * HOP
SAMPLES_PER_STEP = B * T
METRIC = "audio samples/sec (bs=16, 1s, 24 kHz) 1-step infer + GAN train step @1/2/4/8 B200"   # BASELINE.json
REF_FLOPS = {1: 347.6e9, 2: 675.8e9, 4: 1332.2e9}   # SURVEY.md section 8(d), conv/matmul FLOPs


def workload_name(n_timesteps: int) -> str:
    """`config.workload` of both arms (BASELINE.json configs[1])."""
    return f"{MODEL} {n_timesteps}-step inference, synthetic mel (16,100,94) -> (16,24064) per GPU"


def synth_inputs():
    from _cases import mel_input, noise_input
    return mel_input(B, N_MELS, FRAMES, seed=0), noise_input(B, T, seed=1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md 'clocks line')."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, sm_load, mx, reasons, pw = [], [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
                if len(f) > 9 and float(f[9]) >= 50.0:
                    sm_load.append(float(f[1]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                               f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        use = sorted(sm_load if sm_load else sm)         # median over the samples taken under load
        return {"sm_mhz": use[len(use) // 2] if use else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_under_load": len(sm_load), "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons)}


def build_model(device):
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from _synth import synth_state_dict
    torch.manual_seed(0)
    m = MelAudioGenerator(**get_generator_config(MODEL))
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    m.load_state_dict(synth_state_dict(spec, 99), strict=False)
    return m.to(device).eval()


REF_DIR = os.path.join(ROOT, "baseline", "_ref")       # tools/stage_reference.sh (git-ignored, ships with gpurun)


def import_reference():
    """The UNMODIFIED reference package from baseline/_ref (lhotse stubbed as in SURVEY.md App. B:
    it is imported for type names / seed helpers only).  Returns the `flow2gan` module or None."""
    if not os.path.isdir(os.path.join(REF_DIR, "flow2gan")):
        return None
    import types
    l = types.ModuleType("lhotse")
    l.__version__, l.__file__, l.RecordingSet = "stub", "/dev/null", object
    for nm in ("lhotse.dataset", "lhotse.dataset.sampling", "lhotse.dataset.sampling.base", "lhotse.utils"):
        sys.modules.setdefault(nm, types.ModuleType(nm))
    sys.modules.setdefault("lhotse", l)
    sys.modules["lhotse.dataset.sampling.base"].CutSampler = object
    sys.modules["lhotse.utils"].fix_random_seed = lambda s: None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import flow2gan
    return flow2gan


def _best_threads(call, candidates=(8, 12, 16, 24, 32)):
    """The reference's CPU path (mkldnn convs on short sequences) stops scaling well before a big
    host's core count (measured on the 128-core B200 host: 16 threads 0.21-0.27 s per bs-16 call, 8 threads
    0.40 s, 64 threads 0.88 s, 128 threads > 20 s) -- give the baseline its best thread count, chosen on the
    full-size call (best of two timed calls per candidate after one warm-up)."""
    ncpu = os.cpu_count() or 1
    best, cores = None, 1
    for th in sorted({min(ncpu, c) for c in candidates}):
        torch.set_num_threads(th)
        call()
        dt = None
        for _ in range(2):
            t0 = time.perf_counter()
            call()
            d = time.perf_counter() - t0
            dt = d if dt is None or d < dt else dt
        if best is None or dt < best:
            best, cores = dt, th
    torch.set_num_threads(cores)
    return cores


def cpu_reference_rate(n_timesteps: int, iters: int, warmup: int = 1):
    """CPU arm.  kind == "reference": the reference's own code through its own public API --
    flow2gan.get_model(checkpoint=...) -> model.infer(cond, n_timesteps) (flow2gan/__init__.py:29-47,
    models/generator.py:327-366), same synthetic weights (saved as a {"model": state_dict} checkpoint,
    the released-checkpoint format) and the same mel batch as the GPU arm.  kind == "port": the
    oracle restatement, only when baseline/_ref is absent."""
    from _synth import synth_state_dict
    mel, noise = synth_inputs()
    ref = import_reference()
    if ref is not None:
        import tempfile
        from flow2gan.models.config import get_generator_config as ref_cfg
        from flow2gan.models.generator import MelAudioGenerator as RefGen
        import contextlib, io
        m0 = RefGen(**ref_cfg(MODEL))
        like = m0.state_dict()
        sd = synth_state_dict([(k, tuple(v.shape)) for k, v in like.items()], 99, like=like)
        with tempfile.TemporaryDirectory() as td:
            ck = os.path.join(td, "synthetic.pt")
            torch.save({"model": sd}, ck)
            with contextlib.redirect_stdout(io.StringIO()):          # get_model prints the checkpoint path
                model, _ = ref.get_model(model_name=MODEL, hf_model_name=None, checkpoint=ck)
        model.eval()
        call = lambda b=B: model.infer(cond=mel[:b], n_timesteps=n_timesteps)
        kind = "reference"
    else:
        from oracle import flow2gan_oracle as O
        from flow2gan_b200 import get_generator_config
        from flow2gan_b200.generator import MelAudioGenerator
        m = MelAudioGenerator(**get_generator_config(MODEL))
        sd = synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 99)
        cfg = O.generator_config(MODEL)
        call = lambda b=B: O.generator_infer(sd, cfg, mel[:b], noise[:b], None, n_timesteps, False)
        kind = "port"
    times = []
    with torch.inference_mode():
        cores = _best_threads(call)
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            call()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    tot = sum(times)
    return SAMPLES_PER_STEP * len(times) / tot, tot / len(times) * 1e3, cores, kind


def build_model_named(device, model_name: str, seed: int = 99):
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from _synth import synth_state_dict
    torch.manual_seed(0)
    m = MelAudioGenerator(**get_generator_config(model_name))
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    m.load_state_dict(synth_state_dict(spec, seed), strict=False)
    return m.to(device).eval()


# reference-equivalent conv/matmul FLOPs of one bs-16 call
LEGS = {  # key: (model, n_mels, frames, hop, n_timesteps, mel scale, mel shift, reference-equivalent FLOPs)
    "infer_n2": ("mel_24k_base", 100, 94, 256, 2, 1.7, -1.6, 675.8e9),
    "infer_n4": ("mel_24k_base", 100, 94, 256, 4, 1.7, -1.6, 1332.2e9),
    "infer_44k_n4": ("mel_44k_128band_512x_base", 128, 87, 512, 4, 1.4, 0.2, 1252.9e9),
}


def extra_infer_leg(dev, key: str, steps: int, warmup: int, model=None):
    """BASELINE.json configs[1] (2 / 4 ODE steps) and configs[3] (44.1 kHz family, 4 steps): K graph
    replays between two CUDA events, inputs resident in HBM, plus the same call end to end through
    model.infer with pinned host buffers."""
    name, n_mels, frames, hop, n, sc, sh, flops = LEGS[key]
    m = model if model is not None else build_model_named(dev, name)
    T_ = frames * hop
    g = torch.Generator().manual_seed(0)
    mel_h = torch.randn(B, n_mels, frames, generator=g) * sc + sh          # SURVEY 8(d) input statistics
    noise = (torch.randn(B, T_, generator=torch.Generator().manual_seed(1)) * 0.1).to(dev)
    mel = mel_h.to(dev)
    with torch.no_grad():
        plan = m.plan(B, frames, T_, False)
        plan.infer(mel, noise, None, n, False)
        plan.infer(mel, noise, None, n, False)
        graph = plan.graphs[(n, False)][0]
        for _ in range(warmup):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        mel_pin, out_pin = mel_h.pin_memory(), torch.empty(B, T_).pin_memory()
        for _ in range(warmup):
            m.infer(mel_pin, n_timesteps=n, out=out_pin)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            m.infer(mel_pin, n_timesteps=n, out=out_pin)
            torch.cuda.current_stream().synchronize()
        wall = (time.perf_counter() - t0) / steps
    return {"workload": f"{name} {n}-step inference, synthetic mel ({B},{n_mels},{frames}) -> ({B},{T_})",
            "ms_per_step": ms, "value": B * T_ / (ms * 1e-3), "unit": "samples/s",
            "e2e_value": B * T_ / wall, "steps": steps,
            "ref_equiv_gflop": flops / 1e9, "ref_equiv_tflops": flops / (ms * 1e-3) / 1e12}


def tf32_operand_leg(steps: int, warmup: int):
    """The same bench shape with fp32-container TF32 operands in the ConvNeXt-block GEMMs
    (F2G_BLOCK_OPERANDS=tf32 is read at import, hence a child interpreter): the line that sits beside
    the fp16-operand headline (SURVEY.md section 8(d) names TF32 as the precision contract)."""
    env = dict(os.environ, F2G_BLOCK_OPERANDS="tf32")
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--no-train", "--no-legs", "--no-cpu",
                            "--steps", str(steps), "--warmup", str(warmup)], env=env, capture_output=True,
                           text=True, timeout=600)
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                return {"dtype": "tf32 operands -> f32 accumulate", "ms_per_step": d["ms_per_step"],
                        "value": d["value"], "unit": "samples/s", "e2e_value": d["e2e"]["value"],
                        "roofline_frac_of_tf32_peak": d["roofline"]["frac"],
                        "roofline_achieved_tflops": d["roofline"]["achieved"]}
        return {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:
        return {"error": repr(e)[:300]}


def kernel_census(run_once, match=("gemm_pair_kernel", "gemm_tf32_kernel")):
    """GPU time of one call of `run_once` by kernel family, from CUPTI activity records
    (torch.profiler; the records are timestamps taken by the device, not a replay profiler)."""
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run_once()
        torch.cuda.synchronize()
    tot = gemm = glue = 0.0
    n_all = n_gemm = n_glue = 0
    for ev in prof.events():
        if ev.device_type.name != "CUDA" or "memcpy" in ev.name.lower() or "memset" in ev.name.lower():
            if ev.device_type.name == "CUDA":                       # memcpy / memset nodes: glue
                glue += ev.device_time
                n_glue += 1
                tot += ev.device_time
                n_all += 1
            continue
        dt = ev.device_time
        tot += dt
        n_all += 1
        if any(k in ev.name for k in match):
            gemm += dt
            n_gemm += 1
        elif "f2g::" not in ev.name:
            glue += dt
            n_glue += 1
    return {"gpu_busy_us": tot, "gemm_us": gemm, "torch_glue_us": glue, "launches": n_all,
            "gemm_launches": n_gemm, "torch_glue_launches": n_glue}


def gan_train_bench(dev, dist, world, pairs=3, n_timesteps=1):
    """GAN fine-tune step pair (D-iteration + G-iteration on bs=16 x 24000 samples per GPU, incl.
    ScaledAdam updates and, for world > 1, the NCCL gradient all-reduce of the stepped half) --
    the second half of BASELINE.json's metric (finetune.py:590-626)."""
    from _cases import audio_input
    from flow2gan_b200 import get_gan_config, get_generator_config
    from flow2gan_b200.gan import GAN
    from flow2gan_b200.generator import MelAudioGenerator
    from flow2gan_b200.trainer import GANTrainer
    from _synth import synth_state_dict
    torch.manual_seed(0)
    gen = MelAudioGenerator(**get_generator_config(MODEL))
    gen.branch_dropout = 0.0
    gan = GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
    spec = [(k, tuple(v.shape)) for k, v in gan.state_dict().items()]
    gan.load_state_dict(synth_state_dict(spec, 4321), strict=False)
    gan = gan.to(dev)
    tr = GANTrainer(gan, n_timesteps=n_timesteps)
    Tt = 24000
    audio = audio_input(B, Tt, seed=2 + int(os.environ.get("RANK", "0"))).to(dev)
    lens = torch.full((B,), Tt, device=dev, dtype=torch.int64)
    torch.manual_seed(1 + int(os.environ.get("RANK", "0")))
    for _ in range(6):                    # warm-up: eager pair, graph-capture pair, one replayed pair
        tr.step(audio, lens)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2 * pairs):
        info = tr.step(audio, lens)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    roof = None
    if world == 1:          # single process only: the extra eager / profiled pairs would leave the other ranks'
        try:                # collectives without a partner
            from flow2gan_b200 import _lib as L
            # (a) FLOPs as executed: the pair's GEMM descriptors, recorded on one eager D + G iteration
            was = tr.use_graph
            tr.use_graph = False
            L.PROFILE = []
            tr.step(audio, lens)
            tr.step(audio, lens)
            torch.cuda.synchronize()
            rec, L.PROFILE = L.PROFILE, None
            tr.use_graph = was
            flops = sum(r[2] for r in rec)
            # (b) GPU time by kernel family of one replayed pair (CUPTI activity records)
            cen = kernel_census(lambda: (tr.step(audio, lens), tr.step(audio, lens)))
            pk, how = peaks()
            peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / 2.0
            ach = flops / (cen["gemm_us"] * 1e-6) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_pair_kernel / gemm_tf32_kernel (kind::tf32), all launches of one D+G pair",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "gemm_gflop_as_executed": flops / 1e9, "gemm_launches": cen["gemm_launches"],
                    "gemm_ms": cen["gemm_us"] / 1e3, "gpu_busy_ms": cen["gpu_busy_us"] / 1e3,
                    "gemm_share": cen["gemm_us"] / cen["gpu_busy_us"],
                    "torch_glue_share": cen["torch_glue_us"] / cen["gpu_busy_us"],
                    "torch_glue_launches": cen["torch_glue_launches"], "launches_per_pair": cen["launches"],
                    "pair_ref_equiv_tflops": 7261e9 / (ms / pairs * 1e-3) / 1e12,
                    "how": "FLOPs = sum(2MNK) of the pair's GEMM descriptors (one eager pair); times = CUPTI kernel "
                           "records of one graph-replayed pair (torch.profiler)",
                    "peak_source": f"{how}: bf16_tflops_sustained / 2 (TF32 issues at half the bf16 rate; kernels timed "
                                   "inside a long step)"}
        except Exception as e:
            roof = {"error": repr(e)[:300]}
    return {"value": world * 2 * B * Tt * pairs / (ms * 1e-3), "unit": "samples/s", "ms_per_pair": ms / pairs,
            "roofline": roof,
            "pairs": pairs, "n_timesteps": n_timesteps, "global_batch": B * world,
            "allreduce_bytes_per_pair": 0 if world == 1 else (78949542 + 42503752) * 4,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "step_graphs": sum("graph" in e for e in tr._graphs.values()),
            "includes": "LogMel front-end, GAN.forward, backward (one CUDA graph per phase), grad all-reduce "
                        "(world>1), ScaledAdam.step, Eden2"}


def config_for(n_timesteps: int, world: int) -> dict:
    """`config` of BOTH arms (identical keys and values: the workload, nothing arm specific)."""
    return {"workload": workload_name(n_timesteps), "model": MODEL, "n_timesteps": n_timesteps,
            "batch_per_gpu": B, "global_batch": B * world, "mel_shape": [B, N_MELS, FRAMES],
            "samples_per_step_per_gpu": SAMPLES_PER_STEP,
            "parallelism": f"replicas x{world} (no data-path collective)",
            "weights": "synthetic (seeded, tests/_synth.py seed 99), reference state_dict layout",
            "l2": "no explicit flush: every step streams > 126 MB (packed weights ~185 MB + ~90 MB of activations)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every requested step is timed (one step = one bs-16 x 1 s call, 0.3-0.9 s on the host cores);
    # the cap only guards against a K that would run for many minutes
    steps = max(1, min(args.steps, 200))
    W = max(3, args.warmup)
    rate, ms, cores, kind = cpu_reference_rate(args.n_timesteps, steps, warmup=W)
    what = ("the UNMODIFIED reference (baseline/_ref/flow2gan): get_model(checkpoint=synthetic) -> model.infer"
            if kind == "reference" else
            "oracle/flow2gan_oracle.py (baseline/_ref absent: run tools/stage_reference.sh where /root/reference is mounted)")
    line = {
        "impl": "reference", "metric": METRIC, "metric_part": "%d-step infer" % args.n_timesteps,
        "value": rate, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_for(args.n_timesteps, args.gpus),
        "notes": {"arm": what, "torch_threads": cores},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{steps} x (bs=16, 1 s) {args.n_timesteps}-step calls after {W} warm-up calls"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a rank that stops participating must fail the run in minutes, not hold 8 GPUs for the 10-minute default
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    from flow2gan_b200 import _lib as L
    L.lib()
    model = build_model(dev)
    mel_h, noise_h = synth_inputs()
    mel, noise = mel_h.to(dev), noise_h.to(dev)
    n = args.n_timesteps
    W = max(3, args.warmup)
    K = args.steps

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        plan = model.plan(B, FRAMES, T, False)
        plan.infer(mel, noise, None, n, False)          # 1st call: eager
        plan.infer(mel, noise, None, n, False)          # 2nd call: captures the CUDA graph
        graph = plan.graphs[(n, False)][0]
        # launches per replay: count C-ABI calls of one eager pass
        L.COUNT = 0
        plan._steps = plan._prepare_steps(n)
        plan._run(n, False)
        launches_per_step = L.COUNT
        torch.cuda.synchronize()

        # ---------------- kernel-only (HBM-resident inputs), K graph replays ----------------
        # clocks are sampled from the warm-up through the end-to-end loop: the device-timed region
        # alone (K x 0.7 ms) is shorter than nvidia-smi's sampling period
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(W):
            plan.x_audio.copy_(noise)
            graph.replay()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(K):
            graph.replay()      # x_audio keeps evolving; the work per replay is data independent
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        t_dev = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        ms_total = float(t_dev.item())
        value = world * SAMPLES_PER_STEP * K / (ms_total * 1e-3)

        # ---------------- end to end through the public API with host buffers ---------------
        mel_pin = mel_h.pin_memory()
        out_pin = torch.empty(B, T).pin_memory()
        for _ in range(W):
            model.infer(mel_pin, n_timesteps=n, out=out_pin)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e2.record()
        for _ in range(K):
            # pinned host mel -> (H2D into the launch graph's input buffer) -> noise draw -> graph
            # replay -> (D2H into the pinned result buffer); the caller then consumes the audio
            model.infer(mel_pin, n_timesteps=n, out=out_pin)
            torch.cuda.current_stream().synchronize()
        e3.record()
        barrier()
        wall = time.perf_counter() - t0
        ms_e2e = max(e2.elapsed_time(e3), wall * 1e3)
        t_dev = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        e2e_value = world * SAMPLES_PER_STEP * K / (float(t_dev.item()) * 1e-3)

        # ---------------- roofline of the dominant kernel (tcgen05 TF32 GEMM) ---------------
        # The step's GEMM launches (same descriptors, same buffers) are re-issued alone, back to
        # back, from a CUDA graph so that no host submission latency sits between the two events:
        # achieved = sum(2MNK) / (event time / launches) -- the in-step average launch duration.
        roof = roof_other = roof_hbm = None
        if rank == 0:
            L.PROFILE, L.PROFILE_PRE = [], []
            plan.x_audio.copy_(noise)
            plan._run(n, False)
            torch.cuda.synchronize()
            rec, L.PROFILE = L.PROFILE, None
            rec_pre, L.PROFILE_PRE = L.PROFILE_PRE, None
            pk, how = peaks()

            def replay_timed(sub, issue=L.gemm_replay):
                """Launches `sub` re-issued alone, back to back, from a CUDA graph on a side
                stream (no host submission latency between the two events)."""
                side = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    for r in sub:
                        issue(r[0], r[1])
                    side.synchronize()
                    gg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gg, stream=side):
                        for r in sub:
                            issue(r[0], r[1])
                    reps = 10
                    for _ in range(3):
                        gg.replay()
                    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    g0.record()
                    for _ in range(reps):
                        gg.replay()
                    g1.record()
                    side.synchronize()
                return g0.elapsed_time(g1) / reps

            def roof_entry(sub, kernel, peak, peak_note):
                fl = sum(r[2] for r in sub)
                ms = replay_timed(sub)
                ach = fl / (ms * 1e-3) / 1e12
                return {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "traffic": None, "launches_timed": len(sub),
                        "flops_per_launch_avg": fl / len(sub), "us_per_launch_avg": ms * 1e3 / len(sub),
                        "share_of_step_flops": fl / sum(r[2] for r in rec),
                        "how": "the step's launches of this kernel replayed back to back from a CUDA graph, "
                               "CUDA events on the launching stream",
                        "peak_source": f"{how}: {peak_note}"}

            f16 = [r for r in rec if r[3]]
            t32 = [r for r in rec if not r[3]]
            if f16:      # dominant kernel: the ConvNeXt-block contractions (fp16 operands)
                roof = roof_entry(f16, "gemm_pair_kernel<F16> (tcgen05.mma.cta_group::2 kind::f16, fp32 accumulate)",
                                  pk["bf16_tflops"], "bf16_tflops (burst); fp16 issues at the bf16 rate")
                if t32:
                    roof_other = roof_entry(t32, "gemm_pair_kernel (tcgen05.mma.cta_group::2 kind::tf32)",
                                            pk["bf16_tflops"] / 2.0,
                                            "bf16_tflops (burst)/2 -- TF32 issues at half the bf16 rate")
            else:
                roof = roof_entry(t32, "gemm_pair_kernel (tcgen05.mma.cta_group::2 kind::tf32)",
                                  pk["bf16_tflops"] / 2.0, "bf16_tflops (burst)/2 -- TF32 issues at half the bf16 rate")
            # The generator's dominant HBM-bound kernel: the block prologue (mask -> dwconv7 -> BiasNorm ->
            # + cond -> x(1 + time scale) -> fp16), 19 % of the step.  Algorithmic bytes = residual
            # stream read (fp32) + prologue output written (6 B per row x channel).
            try:
                if rec_pre:
                    by = sum(r[2] for r in rec_pre)
                    ms = replay_timed(rec_pre, issue=L.block_pre_replay)
                    ach = by / (ms * 1e-3) / 1e9
                    roof_hbm = {"bound": "hbm", "kernel": "block_pre_kernel (grouped ConvNeXt block prologue)",
                                "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                                "traffic": None, "launches_timed": len(rec_pre),
                                "bytes_per_launch_avg": by / len(rec_pre), "us_per_launch_avg": ms * 1e3 / len(rec_pre),
                                "how": "the step's launches of this kernel replayed back to back from a CUDA graph, "
                                       "CUDA events on the launching stream; inputs stay L2-resident between "
                                       "replays, so this is the kernel's own ceiling, not a DRAM measurement",
                                "peak_source": f"{how}: hbm_gbs"}
                    try:       # the same launches with L2 flushed before each one (256 MB written): DRAM-fed figure
                        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
                        side, evs = torch.cuda.Stream(), []
                        with torch.cuda.stream(side):
                            for rep in range(2):             # first pass warms the instruction cache / allocator
                                evs = []
                                for r in rec_pre:
                                    flush.zero_()
                                    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                                    c0.record()
                                    L.block_pre_replay(r[0], r[1])
                                    c1.record()
                                    evs.append((c0, c1))
                                side.synchronize()
                        cold_ms = sum(a.elapsed_time(b) for a, b in evs)
                        cach = by / (cold_ms * 1e-3) / 1e9
                        roof_hbm["cold_l2"] = {"achieved": cach, "frac": cach / pk["hbm_gbs"],
                                               "us_per_launch_avg": cold_ms * 1e3 / len(rec_pre),
                                               "how": "each launch alone between two CUDA events, after a 256 MB "
                                                      "write that evicts L2 (residual stream read from DRAM)"}
                        del flush
                    except Exception as e:
                        roof_hbm["cold_l2"] = {"error": repr(e)[:200]}
            except Exception as e:                      # never lose the headline line to the extra leg
                roof_hbm = {"error": repr(e)[:200]}
            roof["step_ref_equiv_tflops"] = REF_FLOPS[n] * K / (ms_total * 1e-3) / 1e12
            tr_path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
            if os.path.exists(tr_path):       # dram__bytes_{read,write}.sum of the same launches (ncu --set full)
                tj = json.load(open(tr_path))
                roof["traffic"] = tj.get("dram_bytes_per_launch_avg")
                roof["traffic_source"] = tj.get("source")

    # the other configurations of BASELINE.json (rank 0, N=1 only: they are per-GPU numbers)
    legs = {}
    if rank == 0 and world == 1 and not args.no_legs:
        for key in LEGS:
            try:
                legs[key] = extra_infer_leg(dev, key, steps=max(10, K // 2), warmup=W,
                                            model=model if LEGS[key][0] == MODEL else None)
            except Exception as e:
                legs[key] = {"error": repr(e)[:300]}
    train = None
    if not args.no_train:
        try:
            train = gan_train_bench(dev, dist, world, pairs=args.train_pairs, n_timesteps=n)
        except Exception as e:                       # keep the headline line alive
            train = {"error": repr(e)[:300]}
    clocks = sampler.stop()      # sampled from the warm-up of the headline leg through the train pairs
    if rank == 0 and world == 1 and not args.no_legs:
        legs["tf32_operands"] = tf32_operand_leg(K, W)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu_rate, cpu_ms, cores, cpu_kind = (cpu_reference_rate(n, iters=8 if n == 1 else 3)
                                         if world == 1 and not args.no_cpu else (None, None, None, None))
    line = {
        "metric": METRIC, "metric_part": "%d-step infer (`value`, `e2e`); GAN train step pair under `gan_train`" % n,
        "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16/tf32 operands -> f32 accumulate", "data": "synthetic",
        "config": config_for(n, world),
        "notes": {"timed": "K CUDA-graph replays between two CUDA events"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": B * N_MELS * FRAMES * 4,
                "d2h_bytes_per_step": B * T * 4},
        "gpu_launches": launches_per_step * K,
        "clocks": clocks,
        "roofline": roof,
    }
    if roof_other is not None:
        line["roofline_other"] = roof_other
    if roof_hbm is not None:
        line["roofline_hbm"] = roof_hbm
    for k_, v_ in legs.items():
        line[k_] = v_
    if train is not None:
        line["gan_train"] = train
    if cpu_rate is not None:
        line["cpu_baseline"] = {"value": cpu_rate, "unit": "samples/s", "cores": cores, "kind": cpu_kind,
                                "sample": f"{8 if n == 1 else 3} x (bs=16, 1 s) {n}-step calls, {cpu_ms:.0f} ms each"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-timesteps", type=int, default=1, choices=[1, 2, 4])
    ap.add_argument("--no-train", action="store_true", help="skip the GAN train-step measurement")
    ap.add_argument("--train-pairs", type=int, default=3)
    ap.add_argument("--no-legs", action="store_true", help="skip the 2/4-step, 44.1 kHz and TF32-operand legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
