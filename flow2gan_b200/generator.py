"""BaseAudioGenerator / MelAudioGenerator with the reference's constructor, attributes,
state_dict layout and forward / infer signatures (flow2gan/models/generator.py:30-366); the
computation is delegated to the CUDA engine (flow2gan_b200.engine for inference,
flow2gan_b200.train for the differentiable path)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from .engine import InferencePlan, PackedGenerator
from .modules import AudioConvNeXt, CondEncoder, LinearFilterSpectrogram


class BaseAudioGenerator(nn.Module):
    def __init__(
        self,
        sampling_rate: int = 24000,
        n_ffts: Tuple[int, ...] = (512, 256, 128),
        hop_lengths: Tuple[int, ...] = (256, 128, 64),
        channels: Tuple[int, ...] = (768, 512, 384),
        time_embed_channels: int = 512,
        hidden_factor: int = 3,
        conv_kernel_sizes: Tuple[int, ...] = (7, 7, 7),
        num_layers: Tuple[int, ...] = (8, 8, 8),
        use_cond_encoder: bool = True,
        cond_dim: int = 100,
        cond_hop_length: int = 256,
        cond_enc_channels: int = 512,
        cond_enc_hidden_factor: int = 3,
        cond_enc_conv_kernel_size: int = 7,
        cond_enc_num_layers: int = 4,
        residual_scale: Optional[float] = 1.0,
        init_noise_scale: float = 0.1,
        pred_x1: bool = True,
        branch_reduction: str = "mean",
        spec_scaling_loss: bool = True,
        loss_n_filters: int = 256,
        loss_n_fft: int = 1024,
        loss_hop_length: int = 256,
        loss_power: float = 0.5,
        loss_eps: float = 1e-7,
        loss_scale_min: float = 1e-2,
        loss_scale_max: float = 1e2,
        branch_dropout: float = 0.05,
    ):
        super().__init__()
        self.num_branches = len(n_ffts)
        assert len(hop_lengths) == len(channels) == len(conv_kernel_sizes) == len(num_layers) \
            == self.num_branches
        assert branch_reduction in ("mean", "sum")
        if not (use_cond_encoder and pred_x1 and branch_reduction == "mean" and spec_scaling_loss):
            raise NotImplementedError(
                "only the released configuration family (cond encoder, x1 prediction, mean fusion, "
                "spectral-scaled loss) is built")
        if len(set(num_layers)) != 1:
            raise NotImplementedError("branches must have the same depth (grouped launches)")
        self.sampling_rate = sampling_rate
        self.init_noise_scale = init_noise_scale
        self.pred_x1 = pred_x1
        self.branch_reduction = branch_reduction
        self.spec_scaling_loss = spec_scaling_loss
        self.loss_power, self.loss_eps = loss_power, loss_eps
        self.loss_scale_min, self.loss_scale_max = loss_scale_min, loss_scale_max
        self.branch_dropout = branch_dropout

        self.loss_spec = LinearFilterSpectrogram(sample_rate=sampling_rate, n_fft=loss_n_fft,
                                                 hop_length=loss_hop_length,
                                                 n_filter=loss_n_filters, center=True, power=2.0)
        self.cond_encoder = CondEncoder(cond_dim, cond_enc_channels, cond_enc_hidden_factor,
                                        cond_enc_conv_kernel_size, cond_enc_num_layers,
                                        residual_scale)
        self.estimators = nn.ModuleList([
            AudioConvNeXt(n_ffts[i], hop_lengths[i], cond_hop_length, channels[i],
                          cond_enc_channels, time_embed_channels, hidden_factor,
                          conv_kernel_sizes[i], num_layers[i], residual_scale)
            for i in range(self.num_branches)])
        self.apply(self._init_weights)
        self._packed: Optional[PackedGenerator] = None
        self._plans: Dict[tuple, InferencePlan] = {}
        self._block_operands: Optional[str] = None      # None = engine default (fp16); "tf32" after a range fallback

    @torch.no_grad()
    def _init_weights(self, m):
        # generator.py:122-127
        if isinstance(m, (nn.Conv1d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=0.015)
            if isinstance(getattr(m, "bias", None), Tensor):
                nn.init.constant_(m.bias, 0)

    # ---------------------------------------------------------------- engine plumbing
    def _apply(self, fn, *a, **k):           # .to()/.cuda()/.double() move storage
        self._packed, self._plans = None, {}
        return super()._apply(fn, *a, **k)

    def __deepcopy__(self, memo):
        # `copy.deepcopy(model).to(torch.float64)` is how the training scripts create model_avg
        # (finetune.py:902, pretrain.py:777).  Packed weights and captured CUDA graphs are derived
        # from the parameters and are not copyable: the copy starts without them.
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_packed" else {} if k == "_plans" else copy.deepcopy(v, memo)
        return new

    def _require_cuda(self) -> None:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("flow2gan_b200: the generator runs on CUDA only (sm_100a kernels, "
                               "no CPU fallback); call model.to('cuda') first")

    def packed(self) -> PackedGenerator:
        self._require_cuda()
        if self._packed is None:
            self._packed = PackedGenerator(self)
        elif self._packed.stale():
            # in place: plans and captured graphs keep pointing at the same buffers; plans notice
            # the new `version` and recompute their derived caches (time-scale vectors)
            self._packed.refresh()
        return self._packed

    def plan(self, B: int, Fm: int, T: int, masked: bool) -> InferencePlan:
        pk = self.packed()
        key = (B, Fm, T, masked)
        p = self._plans.get(key)
        if p is None:
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            p = InferencePlan(self, pk, B, Fm, T, masked)
            self._plans[key] = p
        return p

    # ---------------------------------------------------------------- reference surface
    def infer(self, noise: Tensor, cond: Tensor, audio_lens: Optional[Tensor] = None,
              n_timesteps: int = 1, clamp_pred: bool = False) -> Tensor:
        """Euler sampler on an already-encoded `cond` (generator.py:236-271)."""
        B, _, Fm = cond.shape
        p = self.plan(B, Fm, noise.shape[-1], audio_lens is not None)
        with torch.no_grad():
            return p.infer_from_cond(cond.float(), noise.float(), audio_lens, n_timesteps, clamp_pred)

    def forward(self, *args, **kwargs):
        from .train import generator_fm_loss
        return generator_fm_loss(self, *args, **kwargs)


class MelAudioGenerator(BaseAudioGenerator):
    """Mel-conditioned audio generator (generator.py:274-366)."""

    def __init__(self, n_mels: int = 100, mel_n_fft: int = 1024, mel_hop_length: int = 256,
                 max_add_noise_scale: float = 0.0, **kwargs):
        super().__init__(cond_dim=n_mels, cond_hop_length=mel_hop_length, **kwargs)
        self.n_mels = n_mels
        self.mel_n_fft = mel_n_fft
        self.mel_hop_length = mel_hop_length
        self.max_add_noise_scale = max_add_noise_scale

    def _maybe_noisy_cond(self, cond: Tensor) -> Tensor:
        if self.training and self.max_add_noise_scale > 0.0:      # generator.py:306-309,342-345
            e = torch.randn_like(cond) * torch.rand(cond.shape[0], 1, 1, device=cond.device) \
                * self.max_add_noise_scale
            cond = cond + e
        return cond

    def infer(self, cond: Tensor, audio_lens: Optional[Tensor] = None, n_timesteps: int = 1,
              clamp_pred: bool = False, noise: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
        """mel (B, n_mels, F) -> audio (B, T); T = F*mel_hop_length or audio_lens.max().
        `noise` (optional, extension) pins the initial noise; by default it is drawn from
        torch's global RNG exactly like the reference (generator.py:356).
        Extensions for host-side callers: `cond` may be a host tensor (copied straight into the
        launch graph's input buffer) and `out` (host, ideally pinned, or device) receives the audio
        asynchronously on the current stream instead of a new device tensor.
        With grad enabled and parameters requiring grad this call is differentiable (GAN
        G-phase, gan.py:138-143) and runs through flow2gan_b200.train instead of the graph."""
        self._require_cuda()
        dev = next(self.parameters()).device
        differentiable = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if differentiable or (self.training and self.max_add_noise_scale > 0.0):
            cond = cond.to(dev)
        cond = self._maybe_noisy_cond(cond)
        if audio_lens is None:
            length = cond.shape[2] * self.mel_hop_length
        elif getattr(self, "_static_length", None):
            length = self._static_length        # whole-step graph capture: no host read-back
        else:
            length = int(audio_lens.max().item())
        if differentiable:
            if noise is None:
                noise = torch.randn((cond.shape[0], length), device=dev, dtype=cond.dtype) * self.init_noise_scale
            from .train import generator_infer_with_grad
            return generator_infer_with_grad(self, cond, noise, audio_lens, n_timesteps, clamp_pred)
        p = self.plan(cond.shape[0], cond.shape[2], length, audio_lens is not None)
        capturing = torch.cuda.is_current_stream_capturing()
        with torch.no_grad():
            if p.f16 and not capturing and int(p.sat_host[0]) != 0:
                # deferred report of an EARLIER graph replay of this plan (the flag travels to pinned host
                # memory behind every launch sequence): that result is already with the caller, so say so
                # loudly, and serve this and all later calls with TF32 operands
                p = self._fp16_fallback(p, cond, length, audio_lens, "an earlier call")
            rng = None
            if noise is None and p.f16 and not capturing and not p._range_checked:
                rng = torch.cuda.get_rng_state(dev)       # the checked first call may have to be repeated
            # inside an outer stream capture the plan's launches become part of that graph;
            # noise=None: drawn from the global RNG straight into the plan's sample buffer
            y = p.infer(cond, None if noise is None else noise.float(), audio_lens, n_timesteps, clamp_pred,
                        use_graph=not capturing, noise_scale=self.init_noise_scale, out=out)
            if p.f16 and not capturing and not p._range_checked:
                # first call of a plan (it runs eagerly anyway): one 4-byte read-back decides whether this
                # model / input scale fits fp16 operands; if not, redo the call with TF32 operands
                p._range_checked = True
                w_err = self._packed.fp16_weight_error()
                if w_err > 1e-3:
                    p.sat.fill_(4)                       # bit 2: weights do not survive the fp16 conversion
                if int(p.sat.item()) != 0:
                    p = self._fp16_fallback(p, cond, length, audio_lens, "this call (repeated with TF32 operands)")
                    if rng is not None:
                        torch.cuda.set_rng_state(rng, dev)
                    y = p.infer(cond, None if noise is None else noise.float(), audio_lens, n_timesteps, clamp_pred,
                                use_graph=False, noise_scale=self.init_noise_scale, out=out)
            return y

    def _fp16_fallback(self, p: InferencePlan, cond: Tensor, length: int, audio_lens, when: str) -> InferencePlan:
        """fp16 operands share TF32's 11-bit significand but not its exponent range: when a block operand
        (prologue output, hidden activation, conditioning row) left +-65504 the kernels saturated it and
        raised the plan's range flag.  Switch this model to TF32 operands (fp32 containers) for good."""
        import logging
        bits = int(p.sat.item()) | int(p.sat_host[0])
        logging.warning("flow2gan_b200: fp16 block operands out of range (flag %d: %s) in %s -- switching this "
                        "model to TF32 operands", bits,
                        " + ".join(n for b, n in ((1, "GEMM epilogue"), (2, "block prologue"), (4, "weights")) if bits & b), when)
        self._block_operands = "tf32"
        self._plans = {}
        return self.plan(cond.shape[0], cond.shape[2], length, audio_lens is not None)
