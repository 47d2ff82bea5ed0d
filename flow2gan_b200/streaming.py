"""Whole-utterance and chunked ("streaming") synthesis on top of `model.infer`, mirroring
flow2gan/bin/infer_dir.py:99-168 (`infer_audio`, `streaming_infer_audio`): the mel sequence is cut
into `chunk_size`-frame pieces, each piece is synthesised with 24 frames of context on either side
(3 taps x 8 layers of the k=7 depthwise convs) and the context samples are trimmed before
concatenation.

Same call shape and results as the reference (chunks are synthesised in order, one `model.infer`
call each, so the global-RNG noise draws follow the reference's order).  Two extensions:
  * `batch_chunks=True` stacks all equal-length interior chunks along the batch axis and
    synthesises them in ONE call (chunks are independent; at most four distinct shapes exist, so
    at most four cached launch graphs are used) -- the noise draws are then taken per call, not
    per chunk;
  * `noise_fn(chunk_index, (B, samples)) -> Tensor` pins the initial noise (parity tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor, nn

SIDE_CONTEXT_FRAMES = 3 * 8       # infer_dir.py:140 (conv_kernel_size = 7, 8 layers)


def _prepare_cond(model: nn.Module, cond_module: Optional[nn.Module], audio: Optional[Tensor],
                  cond: Optional[Tensor]) -> Tensor:
    assert (audio is not None) or (cond is not None), "Either audio or cond should be provided."
    device = next(model.parameters()).device
    if cond is None:
        return cond_module(audio.to(device))          # (batch, n_mels, frames)
    return cond.to(device)


def infer_audio(params, model: nn.Module, cond_module: Optional[nn.Module], audio: Optional[Tensor] = None,
                cond: Optional[Tensor] = None) -> Tensor:
    """infer_dir.py:99-123: one `model.infer(clamp_pred=True)` over the whole utterance."""
    with torch.inference_mode():
        c = _prepare_cond(model, cond_module, audio, cond)
        pred = model.infer(cond=c, n_timesteps=params.n_timesteps, clamp_pred=True)
    return pred.cpu()


def chunk_plan(num_frames: int, chunk: int, hop: int,
               side_context: int = SIDE_CONTEXT_FRAMES) -> List[Tuple[int, int, int, int]]:
    """(frame_start, frame_end, left_pad_samples, right_pad_samples) per chunk (infer_dir.py:141-149)."""
    assert chunk > 0 and num_frames > 0
    plan = []
    for i in range((num_frames + chunk - 1) // chunk):
        f0 = max(0, i * chunk - side_context)
        f1 = min(num_frames, (i + 1) * chunk + side_context)
        plan.append((f0, f1, (i * chunk - f0) * hop, (f1 - (i + 1) * chunk) * hop))
    return plan


def streaming_infer_audio(params, model: nn.Module, cond_module: Optional[nn.Module],
                          audio: Optional[Tensor] = None, cond: Optional[Tensor] = None, *,
                          batch_chunks: bool = False,
                          noise_fn: Optional[Callable[[int, Tuple[int, int]], Tensor]] = None) -> Tensor:
    """infer_dir.py:126-168.  `params.chunk_size` in mel frames, `params.n_timesteps` ODE steps."""
    hop = model.mel_hop_length
    with torch.inference_mode():
        c = _prepare_cond(model, cond_module, audio, cond)
        B, _, num_frames = c.shape
        plan = chunk_plan(num_frames, params.chunk_size, hop)
        outs: List[Optional[Tensor]] = [None] * len(plan)

        def trim(a: Tensor, lp: int, rp: int) -> Tensor:
            # a negative right pad (last chunk shorter than chunk_size) keeps the whole tail, as
            # the reference's slice `[:, lp : size - rp]` does
            return a[:, lp: a.size(1) - rp]

        def run(idx: List[int]) -> None:
            f0, f1, _, _ = plan[idx[0]]
            pieces = [c[:, :, plan[i][0]:plan[i][1]] for i in idx]
            x = pieces[0] if len(idx) == 1 else torch.cat(pieces, 0)
            kw = {}
            if noise_fn is not None:
                nz = [noise_fn(i, (B, (f1 - f0) * hop)) for i in idx]
                kw["noise"] = (nz[0] if len(idx) == 1 else torch.cat(nz, 0)).to(c.device)
            a = model.infer(cond=x.contiguous(), n_timesteps=params.n_timesteps, clamp_pred=True, **kw)
            for j, i in enumerate(idx):
                outs[i] = trim(a[j * B:(j + 1) * B], plan[i][2], plan[i][3])

        if batch_chunks:
            by_len = {}
            for i, (f0, f1, _, _) in enumerate(plan):
                by_len.setdefault(f1 - f0, []).append(i)
            for idx in by_len.values():
                run(idx)
        else:
            for i in range(len(plan)):
                run([i])
        pred = torch.cat(outs, dim=-1)
    return pred.cpu()
