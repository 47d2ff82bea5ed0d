"""Checkpoint files in the reference's layout (flow2gan/checkpoint.py:40-168): one dict with
"model", "optimizer", "scheduler", "grad_scaler", "sampler", optional "model_avg" / "model_ema" /
"optimizer_disc" / "scheduler_disc", plus the caller's `params` merged at top level.  Files
written here load with the reference's load_checkpoint and vice versa (tests/test_host_cpu.py).

The model-averaging functions of the same reference file live in flow2gan_b200.averaging and are
re-exported here so `from flow2gan.checkpoint import ...` maps one to one."""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional

import torch
from torch import nn

from .averaging import (average_checkpoints, average_checkpoints_with_averaged_model,  # noqa: F401
                        average_state_dict, update_averaged_model, update_ema_model)


def _unwrap(model: nn.Module) -> nn.Module:
    return model.module if isinstance(model, nn.parallel.DistributedDataParallel) else model


def save_checkpoint(filename, model: nn.Module, model_avg: Optional[nn.Module] = None,
                    model_ema: Optional[nn.Module] = None, params: Optional[Dict[str, Any]] = None,
                    optimizer=None, scheduler=None, scaler=None, sampler=None, optimizer_disc=None,
                    scheduler_disc=None, rank: int = 0, downcast_avg_in_place: bool = True) -> None:
    """Rank 0 writes the file.  `downcast_avg_in_place=True` reproduces the reference exactly:
    `model_avg.to(torch.float32).state_dict()` converts the LIVE averaged model to fp32
    (checkpoint.py:94-98), so later running-average updates accumulate in fp32;
    flow2gan_b200.averaging handles both.  False stores the same fp32 file content but leaves the
    accumulators in fp64."""
    if rank != 0:
        return
    logging.info(f"Saving checkpoint to {filename}")

    def state(obj):
        return obj.state_dict() if obj is not None else None

    def fp32_state(m: nn.Module):
        if downcast_avg_in_place:
            return m.to(torch.float32).state_dict()
        return {k: (v.to(torch.float32) if torch.is_floating_point(v) else v.clone())
                for k, v in m.state_dict().items()}

    ckpt = {"model": _unwrap(model).state_dict(), "optimizer": state(optimizer), "scheduler": state(scheduler),
            "grad_scaler": state(scaler), "sampler": state(sampler)}
    if model_avg is not None:
        ckpt["model_avg"] = fp32_state(model_avg)
    if model_ema is not None:
        ckpt["model_ema"] = fp32_state(model_ema)
    if optimizer_disc is not None:
        ckpt["optimizer_disc"] = optimizer_disc.state_dict()
    if scheduler_disc is not None:
        ckpt["scheduler_disc"] = scheduler_disc.state_dict()
    if params:
        for k, v in params.items():
            assert k not in ckpt, k
            ckpt[k] = v
    torch.save(ckpt, filename)


def load_checkpoint(filename, model: nn.Module, model_avg: Optional[nn.Module] = None,
                    model_ema: Optional[nn.Module] = None, optimizer=None, scheduler=None, scaler=None,
                    sampler=None, optimizer_disc=None, scheduler_disc=None, strict: bool = False) -> Dict[str, Any]:
    """Restores whatever is given and present; returns the remaining entries (the saved `params`).
    Checkpoints saved from a DDP-wrapped model ('module.' prefix) are accepted (checkpoint.py:126-138)."""
    logging.info(f"Loading checkpoint from {filename}")
    ckpt = torch.load(filename, map_location="cpu", weights_only=False)
    src = ckpt.pop("model")
    if next(iter(src)).startswith("module."):
        logging.info("Loading checkpoint saved by DDP")
        src = dict(src)
        dst = model.state_dict()
        for key in dst.keys():
            dst[key] = src.pop("module." + key)
        assert len(src) == 0, list(src)[:5]
        src = dst
    model.load_state_dict(src, strict=strict)
    for name, m in (("model_avg", model_avg), ("model_ema", model_ema)):
        if m is not None and name in ckpt:
            m.load_state_dict(ckpt.pop(name), strict=strict)
    for name, obj in (("optimizer", optimizer), ("scheduler", scheduler), ("grad_scaler", scaler),
                      ("sampler", sampler), ("optimizer_disc", optimizer_disc), ("scheduler_disc", scheduler_disc)):
        s = ckpt.get(name)
        if obj and s:
            obj.load_state_dict(s)
            ckpt.pop(name)
    return ckpt
