"""The GAN fine-tuning iteration of flow2gan/bin/finetune.py:427-492,590-626 as a small class:
alternating discriminator / generator phases, the recipe's loss weights, ScaledAdam + Eden2 per
half, gradient averaging across data-parallel ranks for the half being stepped."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from .dist import GradBuckets, broadcast_module_state
from .gan import GAN
from .modules import LogMelSpectrogram
from .optim import Eden2, ScaledAdam

LOSS_WEIGHTS = dict(mp=1.0, mr=0.1, fm_mp=1.0, fm_mr=0.1, mel=45.0)      # finetune.py defaults


class GANTrainer:
    def __init__(self, gan: GAN, lr_g: float = 2e-3, lr_d: float = 2e-2, lr_batches_g: float = 20000,
                 lr_batches_d: float = 5000, warmup_start: float = 0.1, n_timesteps: int = 1,
                 weights: Optional[Dict[str, float]] = None, use_graph: bool = True,
                 graph_warmup: int = 1):
        self.gan = gan
        broadcast_module_state(gan)              # DDP's construction-time sync of parameters + buffers
        g = gan.generator
        self.cond_module = LogMelSpectrogram(g.sampling_rate, g.mel_n_fft, g.mel_hop_length, g.n_mels) \
            .to(next(gan.parameters()).device)
        self.opt_g = ScaledAdam(gan.generator.named_parameters(), lr=lr_g, clipping_scale=2.0)
        self.opt_d = ScaledAdam(gan.discriminator.named_parameters(), lr=lr_d, clipping_scale=2.0)
        self.sched_g = Eden2(self.opt_g, lr_batches_g, warmup_start=warmup_start)
        self.sched_d = Eden2(self.opt_d, lr_batches_d, warmup_start=warmup_start)
        self.buckets_g = GradBuckets(gan.generator.parameters())
        self.buckets_d = GradBuckets(gan.discriminator.parameters())
        self.n_timesteps = n_timesteps
        self.w = dict(LOSS_WEIGHTS, **(weights or {}))
        self.train_disc = True
        self.use_graph = use_graph
        self.graph_warmup = graph_warmup          # eager calls per (phase, shape) before capture
        self._graphs: Dict[tuple, dict] = {}
        self._pool = torch.cuda.graph_pool_handle() if use_graph and torch.cuda.is_available() else None

    # ------------------------------------------------------------------ one phase, eager
    def _forward_backward(self, disc: bool, audio: Tensor, audio_lens: Tensor) -> Dict[str, Tensor]:
        with torch.no_grad():
            cond = self.cond_module(audio)                                   # finetune.py:441
        w = self.w
        if disc:
            d_mp, d_mr = self.gan(cond=cond, audio=audio, audio_lens=audio_lens,
                                  n_timesteps=self.n_timesteps, train_disc=True)
            loss = d_mp * w["mp"] + d_mr * w["mr"]
            loss.backward()
            return {"disc_loss": loss.detach(), "disc_loss_mp": d_mp.detach(), "disc_loss_mr": d_mr.detach()}
        g_mp, g_mr, fm_mp, fm_mr, mel = self.gan(cond=cond, audio=audio, audio_lens=audio_lens,
                                                 n_timesteps=self.n_timesteps, train_disc=False)
        loss = (g_mp * w["mp"] + g_mr * w["mr"] + fm_mp * w["fm_mp"] + fm_mr * w["fm_mr"]
                + mel * w["mel"])
        loss.backward()
        return {"gen_loss": loss.detach(), "mel_recon_loss": mel.detach()}

    def _update(self, disc: bool) -> None:
        if disc:
            self.buckets_d.allreduce_mean()
            self.opt_d.step()
            self.sched_d.step_batch()
        else:
            self.buckets_g.allreduce_mean()
            self.opt_g.step()
            self.sched_g.step_batch()

    # ------------------------------------------------------------------ one phase, CUDA graph
    def _capture(self, disc: bool, audio: Tensor, audio_lens: Tensor) -> Optional[dict]:
        """Captures cond + GAN.forward + backward of one phase for this batch shape.  What stays
        on the host: the LimitParamValue coin flips (refilled into a device vector before each
        replay), the gradient all-reduce, ScaledAdam / Eden2."""
        from . import train as T
        gen = self.gan.generator
        length = audio.shape[-1]
        if int(audio_lens.max()) != length:          # finetune batches are padded to their longest item
            return None
        ent = {"audio": audio.clone(), "lens": audio_lens.clone(), "draws": T.DrawBuffer(audio.device)}
        half = self.gan.discriminator if disc else self.gan.generator
        (self.opt_d if disc else self.opt_g).zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        T._draws, gen._static_length = ent["draws"], length
        try:
            with torch.cuda.graph(graph, pool=self._pool):
                ent["info"] = self._forward_backward(disc, ent["audio"], ent["lens"])
        except Exception as e:                       # keep training alive on the eager path
            import logging
            logging.warning("GANTrainer: step-graph capture failed (%r); staying eager", e)
            self.use_graph = False
            (self.opt_d if disc else self.opt_g).zero_grad(set_to_none=True)
            return None
        finally:
            T._draws, gen._static_length = None, None
        ent["draws"].count = ent["draws"].i
        ent["graph"] = graph
        # The graph writes this phase's gradients into the tensors that became `.grad` during the
        # capture (private-pool memory, kept alive here).  Any later zero_grad(set_to_none=True) --
        # the eager path, or the capture of another batch shape -- rebinds `.grad` elsewhere, so
        # every replay re-attaches exactly these tensors before the all-reduce / optimizer read them.
        ent["grads"] = [(p, p.grad) for p in half.parameters()]
        return ent

    def step(self, audio: Tensor, audio_lens: Tensor) -> Dict[str, Tensor]:
        """One iteration on one batch: D-phase or G-phase (they alternate, finetune.py:612-626).
        With use_graph (default) the third call of a phase with a given batch shape captures the
        phase's forward+backward (~2.6 k launches) into a CUDA graph; later calls replay it."""
        disc = self.train_disc
        ent = None
        # The generator's packed (TF32 / fp16) weight copies are refreshed HERE, eagerly and in place
        # (engine.PackedGenerator), never inside a phase graph: every graph of every batch shape
        # reads the same persistent buffers, whichever phase / path ran before it.
        self.gan.generator.packed()
        if self.use_graph:
            key = (disc, tuple(audio.shape))
            ent = self._graphs.get(key)
            if ent is None:
                ent = self._graphs[key] = {"calls": 0}
            if "graph" not in ent:
                ent["calls"] += 1
                if ent["calls"] > self.graph_warmup:
                    cap = self._capture(disc, audio, audio_lens)
                    if cap is not None:
                        ent.update(cap)
            if "graph" not in ent:
                ent = None
        if ent is not None:
            ent["audio"].copy_(audio)
            ent["lens"].copy_(audio_lens)
            ent["draws"].refill()
            ent["graph"].replay()
            for p, g in ent["grads"]:
                p.grad = g
            info = ent["info"]
        else:
            (self.opt_d if disc else self.opt_g).zero_grad()
            info = self._forward_backward(disc, audio, audio_lens)
        self._update(disc)
        self.train_disc = not disc
        return info
