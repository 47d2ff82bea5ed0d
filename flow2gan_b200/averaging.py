"""fp64 running model average on the device.

Same functions, argument meaning and in-place behaviour as flow2gan/checkpoint.py:
``average_state_dict`` (:504-531), ``update_averaged_model`` (:378-409), ``update_ema_model``
(:411-441) and ``average_checkpoints_with_averaged_model`` (:443-501).  The reference walks the
state_dict with three torch ops per tensor (``v *= w1; v += cur * w2; v *= scale`` -- ~1300
launches for the generator); here all tensors are updated by ONE ``f2g_average_update`` launch that
keeps every rounding step of that sequence (results are bit-identical, tests/test_datapath_cpu.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L

_CHUNK = 4096
_TABLES: Dict[tuple, Tuple[Tensor, Tensor, int, list]] = {}


def unique_float_keys(state_dict: Dict[str, Tensor]) -> List[str]:
    """Keys the reference updates: first name of every distinct storage, floating point only
    (checkpoint.py:516-527)."""
    seen, names = set(), []
    for k, v in state_dict.items():
        p = v.data_ptr()
        if p in seen:
            continue
        seen.add(p)
        names.append(k)
    return [k for k in names if torch.is_floating_point(state_dict[k])]


def build_table(pairs: List[Tuple[Tensor, Tensor]]):
    """host-side (F2GAvgTensor[], chunk list) for (avg, cur) tensor pairs"""
    tab = (L.F2GAvgTensor * len(pairs))()
    chunks: List[int] = []
    for i, (a, c) in enumerate(pairs):
        tab[i].avg, tab[i].cur = a.data_ptr(), c.data_ptr()
        tab[i].numel = a.numel()
        tab[i].cur_is_f64 = int(c.dtype == torch.float64)
        tab[i].avg_is_f32 = int(a.dtype == torch.float32)
        for ci in range((a.numel() + _CHUNK - 1) // _CHUNK):
            chunks += [i, ci]
    return tab, chunks


def average_state_dict(state_dict_1: Dict[str, Tensor], state_dict_2: Dict[str, Tensor], weight_1: float,
                       weight_2: float, scaling_factor: float = 1.0) -> Dict[str, Tensor]:
    """state_dict_1 = (state_dict_1 * weight_1 + state_dict_2 * weight_2) * scaling_factor, in place.
    state_dict_1: fp64 CUDA tensors (the reference's `model_avg`, finetune.py:902) -- or fp32 ones, which
    is what the reference's model_avg becomes after its first save_checkpoint (checkpoint.py:94-95)."""
    keys = unique_float_keys(state_dict_1)
    pairs, keep_alive = [], []
    for k in keys:
        a = state_dict_1[k]
        if a.numel() == 0:
            continue
        L.require_cuda(a, f"average_state_dict: '{k}'")
        if not (a.dtype in (torch.float64, torch.float32) and a.is_contiguous()):
            raise RuntimeError(f"average_state_dict: '{k}' must be a contiguous fp64 / fp32 accumulator "
                               f"(got {a.dtype})")
        c = state_dict_2[k]
        if c.device != a.device or c.dtype not in (torch.float32, torch.float64) or not c.is_contiguous():
            c = c.to(device=a.device)
            c = (c if c.dtype in (torch.float32, torch.float64) else c.to(torch.float64)).contiguous()
            keep_alive.append(c)
        assert c.numel() == a.numel(), k
        pairs.append((a, c))
    if not pairs:
        return state_dict_1
    key = tuple((a.data_ptr(), c.data_ptr(), a.numel(), a.dtype, c.dtype) for a, c in pairs)
    ent = _TABLES.get(key) if not keep_alive else None
    if ent is None:
        tab, chunks = build_table(pairs)
        dev = pairs[0][0].device
        tab_dev = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(dev)
        chunks_dev = torch.tensor(chunks, dtype=torch.int32, device=dev)
        ent = (tab_dev, chunks_dev, len(chunks) // 2, [p[0] for p in pairs])
        if not keep_alive:
            if len(_TABLES) > 8:
                _TABLES.clear()
            _TABLES[key] = ent
    L.average_update(ent[0], ent[1], ent[2], weight_1, weight_2, scaling_factor)
    return state_dict_1


def _unwrap(m: nn.Module) -> nn.Module:
    return m.module if isinstance(m, nn.parallel.DistributedDataParallel) else m


def update_averaged_model(params, model_cur: nn.Module, model_avg: nn.Module) -> None:
    """model_avg = model_cur * (average_period / batch_idx_train) + model_avg * (1 - that)
    (checkpoint.py:378-409; called every `average_period` batches, finetune.py:636-646)."""
    get = params.get if isinstance(params, dict) else (lambda k: getattr(params, k))
    weight_cur = get("average_period") / get("batch_idx_train")
    average_state_dict(model_avg.state_dict(), _unwrap(model_cur).state_dict(), 1 - weight_cur, weight_cur)


def update_ema_model(ema_decay: float, model_cur: nn.Module, model_ema: nn.Module) -> None:
    """model_ema = model_ema * ema_decay + model_cur * (1 - ema_decay) (checkpoint.py:411-441)."""
    average_state_dict(model_ema.state_dict(), _unwrap(model_cur).state_dict(), ema_decay, 1 - ema_decay)


def average_checkpoints_with_averaged_model(filename_start: str, filename_end: str,
                                            device: torch.device = torch.device("cuda")) -> Dict[str, Tensor]:
    """Average over (start, end] from the two checkpoints' `model_avg` (checkpoint.py:443-501):
    avg = (model_end + model_start * (weight_start / weight_end)) * weight_end."""
    sd_start = torch.load(filename_start, map_location=device, weights_only=False)
    sd_end = torch.load(filename_end, map_location=device, weights_only=False)
    start, end = sd_start["batch_idx_train"], sd_end["batch_idx_train"]
    interval = end - start
    assert interval > 0, interval
    weight_end = end / interval
    weight_start = 1 - weight_end
    avg = sd_end["model_avg"]
    average_state_dict(avg, sd_start["model_avg"], 1.0, weight_start / weight_end, weight_end)
    return avg


def average_checkpoints(filenames, device: torch.device = torch.device("cuda")) -> Dict[str, Tensor]:
    """Plain mean of the `model` entries of several checkpoints (checkpoint.py:171-212, used by
    bin/save_averaged_model.py:148-157).  The running sums are one `f2g_average_update` launch per
    checkpoint (v * 1 + cur * 1 is the reference's fp32 `avg[k] += state_dict[k]` bit for bit); the
    final division is the reference's own op on the same device, `avg[k] /= n` per tensor
    (checkpoint.py:207-210) -- on CUDA torch evaluates a tensor / python-scalar division as a
    multiplication by the fp32 reciprocal, which differs from true division by up to 1 ulp, so the
    result is whatever the reference produces on that device, bit for bit."""
    n = len(filenames)
    avg = torch.load(filenames[0], map_location=device, weights_only=False)["model"]
    keys = unique_float_keys(avg)
    seen, int_keys = set(), []
    for k, v in avg.items():
        if v.data_ptr() not in seen:
            seen.add(v.data_ptr())
            if not torch.is_floating_point(v):
                int_keys.append(k)
    for i in range(1, n):
        sd = torch.load(filenames[i], map_location=device, weights_only=False)["model"]
        average_state_dict(avg, sd, 1.0, 1.0)
        for k in int_keys:
            avg[k] += sd[k]
    for k in keys:
        avg[k] /= n
    for k in int_keys:
        avg[k] //= n
    return avg
