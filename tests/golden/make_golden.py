"""Generate golden vectors by running the REAL reference (imported from /root/reference,
lhotse stubbed as in SURVEY.md App. B) on small seeded inputs.  Run in the build container:

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/ref_*.pt, tests/golden/mel_24k_short.pt.
The reference cannot travel to the GPU box; these files can.  Weights are not stored: they
are re-derived from the reference-layout state_dict *key/shape list* stored in each file
(`sd_spec`) through oracle.perturb-style seeded generation (`synth_state_dict`).
"""
import math
import os
import random
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _cases import audio_input, mel_input, noise_input  # noqa: E402
from _synth import synth_state_dict  # noqa: E402


def import_reference():
    l = types.ModuleType("lhotse")
    l.__version__ = "stub"
    l.__file__ = "/dev/null"
    l.RecordingSet = object
    for n in ("lhotse.dataset", "lhotse.dataset.sampling", "lhotse.dataset.sampling.base", "lhotse.utils"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["lhotse"] = l
    sys.modules["lhotse.dataset.sampling.base"].CutSampler = object
    sys.modules["lhotse.utils"].fix_random_seed = lambda s: None
    sys.path.insert(0, "/root/reference")


def spec_of(sd):
    return [(k, tuple(v.shape)) for k, v in sd.items()]


def grad_summary(named_grads, full_numel=4096):
    out = {}
    for k, g in named_grads:
        if g is None:
            out[k] = None
            continue
        g = g.detach().float()
        e = {"sum": float(g.double().sum()), "l2": float(g.double().pow(2).sum().sqrt()),
             "abs": float(g.double().abs().sum())}
        if g.numel() <= full_numel:
            e["full"] = g.clone()
        else:
            # strided sample over the WHOLE tensor (a contiguous head would be one output channel)
            stride = max(1, g.numel() // 2048)
            e["stride"] = stride
            e["sample"] = g.flatten()[::stride][:2048].clone()
        out[k] = e
    return out


def main():
    torch.set_num_threads(8)
    import_reference()
    from flow2gan.models.config import get_gan_config, get_generator_config
    from flow2gan.models.gan import GAN
    from flow2gan.models.generator import BaseAudioGenerator, MelAudioGenerator
    from flow2gan.models.modules import LogMelSpectrogram
    from flow2gan.optim import Eden2, ScaledAdam
    import flow2gan.models.modules as ref_modules

    # ---- (1) the reference's own fixture: wav <-> mel (pins the mel front-end, a1) -------
    import wave
    import numpy as np
    wav_path = "/root/reference/test_data/wav/1089_134686_000002_000000.wav"
    with wave.open(wav_path, "rb") as w:
        assert w.getframerate() == 24000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).copy()
    mel_ref = torch.load("/root/reference/test_data/mel/1089_134686_000002_000000.pt")
    torch.save({"pcm_int16": torch.from_numpy(pcm), "mel": mel_ref.clone(),
                "sampling_rate": 24000, "n_fft": 1024, "hop": 256, "n_mels": 100,
                "source": "reference test_data/wav|mel/1089_134686_000002_000000"},
               os.path.join(HERE, "mel_24k_short.pt"))
    chk = LogMelSpectrogram(24000, 1024, 256, 100)(torch.from_numpy(pcm).float()[None] / 32768.0)
    print("fixture self-check max-abs", float((chk - mel_ref).abs().max()))

    # ---- (2) generator inference, cfg A, N = 1/2/4, with and without audio_lens ---------
    for name, b, frames, tag, cfg44 in (("mel_24k_base", 2, 24, "24k", False),
                                        ("mel_44k_128band_512x_base", 1, 12, "44k", True)):
        cfg = get_generator_config(name)
        model = MelAudioGenerator(**cfg).eval()
        sd = synth_state_dict(spec_of(model.state_dict()), seed=1234, like=model.state_dict())
        model.load_state_dict(sd)
        mel = mel_input(b, cfg["n_mels"], frames, seed=0, cfg44=cfg44)
        T = frames * cfg["mel_hop_length"]
        noise = noise_input(b, T, seed=1)
        out = {"sd_spec": spec_of(model.state_dict()), "sd_seed": 1234, "model_name": name,
               "mel": mel, "noise": noise}
        with torch.no_grad():
            cond = model.cond_encoder(mel)
            out["cond"] = cond.clone()
            for n in (1, 2, 4):
                out[f"audio_n{n}"] = BaseAudioGenerator.infer(model, noise, cond, None, n, False).clone()
            out["audio_n2_clamp"] = BaseAudioGenerator.infer(model, noise * 30, cond, None, 2, True).clone()
            # per-branch outputs at t=0.25 for finer-grained checks
            t = torch.full((b,), 0.25)
            out["branch_t025"] = torch.stack(
                [est(audio=noise, cond=cond, t=t, audio_lens=None) for est in model.estimators], 1).clone()
            if tag == "24k":
                lens = torch.tensor([T - 150, T - 1300])
                nz = noise[:, : int(lens.max())]
                out["lens"] = lens
                out["audio_lens_n1"] = BaseAudioGenerator.infer(model, nz, cond, lens, 1, False).clone()
        torch.save(out, os.path.join(HERE, f"ref_infer_{tag}.pt"))
        print("infer", tag, {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})

        # ---- (3) stage-1 FM loss (eval mode => no branch dropout / limit hooks) + grads ---
        if tag == "24k":
            audio = audio_input(b, T, seed=2)
            lens = torch.tensor([T, T - 1000])
            torch.manual_seed(77)
            noise_fm = torch.randn_like(audio) * cfg["init_noise_scale"]
            t_fm = torch.rand((b, 1))
            torch.manual_seed(77)
            model.zero_grad()
            loss = model(cond=mel, audio=audio, audio_lens=lens)
            loss.backward()
            keep = [k for k, _ in model.named_parameters()]
            g = grad_summary([(k, p.grad) for k, p in model.named_parameters()])
            torch.save({"sd_spec": spec_of(model.state_dict()), "sd_seed": 1234, "model_name": name,
                        "mel": mel, "audio": audio, "lens": lens, "noise": noise_fm, "t": t_fm,
                        "loss": loss.detach().clone(), "grads": g, "param_order": keep},
                       os.path.join(HERE, "ref_fm_loss_24k.pt"))
            print("fm loss", float(loss))

    # ---- (4) GAN forward, D phase and G phase, losses + grad summaries --------------------
    cfg = get_generator_config("mel_24k_base")
    gen = MelAudioGenerator(**cfg)
    gen.branch_dropout = 0.0                       # finetune.py:414
    gan = GAN(generator=gen, **get_gan_config("gan_multi_scale_mel_recon"))
    sd = synth_state_dict(spec_of(gan.state_dict()), seed=4321, like=gan.state_dict())
    gan.load_state_dict(sd)
    b, T = 2, 6144
    audio = audio_input(b, T, seed=5)
    lens = torch.tensor([T, T])
    mel = LogMelSpectrogram(24000, 1024, 256, 100)(audio)
    out = {"sd_spec": spec_of(gan.state_dict()), "sd_seed": 4321, "audio": audio, "lens": lens, "mel": mel}
    for limit_draw, tag in ((0.0, "limit_on"), (0.99, "limit_off")):
        random_random = random.random
        ref_modules.random.random = lambda: limit_draw          # pin limit_param_value draw
        try:
            for train_disc in (True, False):
                torch.manual_seed(99)
                noise = torch.randn((b, T)) * cfg["init_noise_scale"]
                torch.manual_seed(99)
                gan.zero_grad()
                losses = gan(cond=mel, audio=audio, audio_lens=lens, n_timesteps=1, train_disc=train_disc)
                if train_disc:
                    total = losses[0] * 1.0 + losses[1] * 0.1          # finetune.py:453-454
                else:
                    total = (losses[0] * 1.0 + losses[1] * 0.1 + losses[2] * 1.0
                             + losses[3] * 0.1 + losses[4] * 45.0)      # finetune.py:478-482
                total.backward()
                ph = "d" if train_disc else "g"
                out[f"noise"] = noise
                out[f"{ph}_{tag}_losses"] = torch.stack([l.detach() for l in losses])
                sub = gan.discriminator if train_disc else gan.generator
                pre = "discriminator." if train_disc else "generator."
                out[f"{ph}_{tag}_grads"] = grad_summary(
                    [(pre + k, p.grad) for k, p in sub.named_parameters()], full_numel=2048)
                print("gan", ph, tag, [float(l) for l in losses])
        finally:
            ref_modules.random.random = random_random
    with torch.no_grad():
        gan.eval()
        torch.manual_seed(99)
        out["fake_audio"] = gan.generator.infer(cond=mel, audio_lens=lens, n_timesteps=1).clone()
    torch.save(out, os.path.join(HERE, "ref_gan_24k.pt"))

    # ---- (5) ScaledAdam + Eden2: 45 steps on a small mixed bag of tensors -----------------
    g = torch.Generator().manual_seed(7)
    shapes = [("a.weight", (6, 5, 3)), ("b.weight", (6, 5, 3)), ("a.bias", (6,)), ("b.bias", (6,)),
              ("s.log_scale", ()), ("u.log_scale", ()), ("c.weight", (4, 7)), ("d.scale", (9, 1))]
    params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.3) for _, s in shapes]
    init = [p.detach().clone() for p in params]
    opt = ScaledAdam([(n, p) for (n, _), p in zip(shapes, params)], lr=2e-3, clipping_scale=2.0)
    sched = Eden2(opt, lr_batches=20, warmup_batches=10, warmup_start=0.1)
    grads_all, lrs = [], []
    import logging
    logging.disable(logging.WARNING)
    for step in range(45):
        gs = [torch.randn(s, generator=g) * (5.0 if step in (13, 31) else 1.0) * (0.5 + 0.1 * i)
              for i, (_, s) in enumerate(shapes)]
        if step == 17:
            gs[2] = None
        grads_all.append(gs)
        for p, gr in zip(params, gs):
            p.grad = None if gr is None else gr.clone()
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step_batch()
    torch.save({"shapes": shapes, "init": init, "grads": grads_all, "lrs": lrs,
                "final": [p.detach().clone() for p in params],
                "hyper": dict(lr=2e-3, clipping_scale=2.0, lr_batches=20, warmup_batches=10, warmup_start=0.1)},
               os.path.join(HERE, "ref_scaled_adam.pt"))
    print("scaled adam final[0][0,0]:", params[0][0, 0].tolist(), "lrs", lrs[:3], lrs[-1])


if __name__ == "__main__":
    main()
