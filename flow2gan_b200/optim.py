"""ScaledAdam + Eden2 with the reference's constructor / step / state_dict surface
(flow2gan/optim.py:258-619, :741-951), executed by the fused multi-tensor CUDA step in
csrc/optim.cu: no torch.stack copies, no per-step .item() sync (the host only reads the
100-entry norm ring on the steps where the reference refreshes its clipping threshold)."""
from __future__ import annotations

import ctypes as C
import logging
from collections import defaultdict
from typing import Dict, List, Optional, Tuple, Union

import torch
from torch import Tensor
from torch.optim import Optimizer

from . import _lib as L

_CHUNK = 4096


class _Group:
    """Device tables + per-tensor scalar state for one param group."""

    def __init__(self, params: List[Tensor], names: List[str], period: int, clip_period: int):
        dev = params[0].device
        # same batching rule as BatchedOptimizer.batched_params (optim.py:77-98): group by
        # (str(dtype), *shape); batches ordered by THAT key (optim.py:90-95 sorts
        # `batches_names_keys[i]`, the key tuples) -- the first batch's first parameter holds
        # model_norms / model_norm_threshold / num_clipped in the state_dict, so the order is part
        # of the checkpoint layout
        by_key: Dict[tuple, List[int]] = defaultdict(list)
        for i, p in enumerate(params):
            by_key[(str(p.dtype), *p.shape)].append(i)
        keys = sorted(by_key.keys())
        self.batches = [by_key[k] for k in keys]
        self.params = params
        self.names = names
        self.order = [i for b in self.batches for i in b]          # tensor slot -> param index
        n = len(self.order)
        self.n = n
        self.exp_avg_sq = [torch.zeros(len(b), *params[b[0]].shape, device=dev) for b in self.batches]
        self.delta = [torch.zeros(len(b), *params[b[0]].shape, device=dev) for b in self.batches]
        self.tstate = torch.zeros(n, 8, device=dev)
        self.gstate = torch.tensor([0.0, 1.0, -1.0, 0.0], device=dev)    # tot_norm, clip, threshold, num_clipped
        self.norms = torch.zeros(clip_period, device=dev)
        self.acc = torch.zeros(n, 3, device=dev)
        chunks = []
        for slot, i in enumerate(self.order):
            for c in range((params[i].numel() + _CHUNK - 1) // _CHUNK):
                chunks.append((slot, c))
        self.n_chunks = len(chunks)
        self.chunks = torch.tensor(chunks, dtype=torch.int32, device=dev).contiguous()
        self.tab_host = (L.F2GAdamTensor * n)()
        slot = 0
        for bi, b in enumerate(self.batches):
            for j, i in enumerate(b):
                r = self.tab_host[slot]
                r.v = self.exp_avg_sq[bi][j].data_ptr()
                r.d = self.delta[bi][j].data_ptr()
                r.numel = params[i].numel()
                r.is_scalar = int(params[i].numel() == 1)
                slot += 1
        self.tab_dev = torch.zeros(C.sizeof(L.F2GAdamTensor) * n, dtype=torch.uint8, device=dev)
        self.step = 0
        self.threshold: Optional[float] = None
        self._keep = []

    def upload_table(self):
        self._keep = []
        for slot, i in enumerate(self.order):
            p = self.params[i]
            if not p.is_contiguous():
                raise RuntimeError("ScaledAdam: parameters must be contiguous")
            r = self.tab_host[slot]
            r.p = p.data_ptr()
            g = p.grad
            if g is None:
                r.g = None
            else:
                if g.is_sparse:
                    raise RuntimeError("ScaledAdam optimizer does not support sparse gradients")
                if not g.is_contiguous():
                    g = g.contiguous()
                    self._keep.append(g)
                r.g = g.data_ptr()
        raw = torch.frombuffer(memoryview(self.tab_host).cast("B"), dtype=torch.uint8)
        self.tab_dev.copy_(raw)


class ScaledAdam(Optimizer):
    def __init__(self, params, lr=3e-02, clipping_scale=None, betas=(0.9, 0.98), scalar_lr_scale=0.1,
                 eps=1.0e-08, param_min_rms=1.0e-05, param_max_rms=3.0, scalar_max=10.0,
                 size_update_period=4, clipping_update_period=100):
        defaults = dict(lr=lr, clipping_scale=clipping_scale, betas=betas,
                        scalar_lr_scale=scalar_lr_scale, eps=eps, param_min_rms=param_min_rms,
                        param_max_rms=param_max_rms, scalar_max=scalar_max,
                        size_update_period=size_update_period,
                        clipping_update_period=clipping_update_period)
        self.show_dominant_parameters = True
        if not 1 <= int(size_update_period) <= 4:
            # the fused kernels keep scale_grads in a fixed 4-slot row of the per-tensor state
            # (csrc/optim.cu); the reference's recipes all use 4
            raise ValueError(f"flow2gan_b200.ScaledAdam: size_update_period must be in 1..4, got {size_update_period}")
        groups, names = self._split_names(params)
        super().__init__(groups, defaults)
        assert len(self.param_groups) == len(names)
        self.parameters_names = names
        self._gs: List[Optional[_Group]] = [None] * len(self.param_groups)

    # the four accepted input forms of optim.py:340-445
    def _split_names(self, params_or_named) -> Tuple[List[dict], List[List[str]]]:
        items = list(params_or_named)
        if len(items) == 0:
            raise ValueError("optimizer got an empty parameter list")
        groups, group_names = [], []
        if not isinstance(items[0], dict):
            ps, ns = [], []
            for it in items:
                if isinstance(it, tuple):
                    n, p = it
                else:
                    assert isinstance(it, torch.Tensor)
                    n, p = "foo", it
                    self.show_dominant_parameters = False
                ps.append(p)
                ns.append(n)
            groups.append({"params": ps})
            group_names.append(ns)
        else:
            for g in items:
                if "named_params" in g:
                    named = list(g["named_params"])
                    del g["named_params"]
                    g["params"] = [p for _, p in named]
                    group_names.append([n for n, _ in named])
                else:
                    g["params"] = list(g["params"])
                    group_names.append(["foo" for _ in g["params"]])
                groups.append(g)
        return groups, group_names

    def _group_state(self, gi: int) -> _Group:
        if self._gs[gi] is None:
            g = self.param_groups[gi]
            for p in g["params"]:
                L.require_cuda(p, "ScaledAdam parameter (fused sm_100a kernels)")
                if p.dtype != torch.float32:
                    raise RuntimeError("flow2gan_b200.ScaledAdam runs on fp32 parameters only")
            self._gs[gi] = _Group(list(g["params"]), self.parameters_names[gi],
                                  g["size_update_period"], g["clipping_update_period"])
        return self._gs[gi]

    def _refresh_threshold(self, group: dict, st: _Group) -> None:
        """optim.py:563-603: every clipping_update_period steps (and at 10/20/40) set the
        threshold to clipping_scale x median of the recent grad norms."""
        step, period = st.step, group["clipping_update_period"]
        irregular = [i for i in (10, 20, 40) if i < period]
        if not (step % period == 0 or step in irregular):
            return
        sorted_norms = st.norms.sort()[0].cpu()          # device -> host: ~once per 100 steps
        if step in irregular:
            sorted_norms = sorted_norms[-step:]
        n = sorted_norms.numel()
        quartiles = [sorted_norms[min(n - 1, (n // 4) * k)].item() for k in range(5)]
        median = quartiles[2]
        if median - median != 0:
            raise RuntimeError("Too many grads were not finite")
        thr = group["clipping_scale"] * median
        if step in irregular:
            thr *= 2.0
        st.threshold = thr
        percent_clipped = float(st.gstate[3]) * 100.0 / n       # same read point as the norm ring
        st.gstate[2] = thr
        st.gstate[3] = 0.0                                       # optim.py:585 num_clipped = 0
        logging.warning("Clipping_scale=%s, grad-norm quartiles %s, threshold=%.3e, percent-clipped=%.1f",
                        group["clipping_scale"], " ".join("%.3e" % q for q in quartiles), thr, percent_clipped)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            st = self._group_state(gi)
            st.upload_table()
            h = L.F2GAdamHyper()
            h.lr, h.scalar_lr_scale = group["lr"], group["scalar_lr_scale"]
            h.beta1, h.beta2 = group["betas"]
            h.eps, h.param_min_rms, h.param_max_rms = group["eps"], group["param_min_rms"], group["param_max_rms"]
            h.scalar_max = group["scalar_max"]
            h.size_update_period = group["size_update_period"]
            h.clipping_update_period = group["clipping_update_period"]
            h.use_clipping = int(group["clipping_scale"] is not None)
            args = (st.tab_dev, st.n, st.chunks, st.n_chunks, st.acc, st.tstate, st.gstate, st.norms, st.step)
            L.scaled_adam_step(*args, 0, h)
            if h.use_clipping and st.step > 0:
                self._refresh_threshold(group, st)
            L.scaled_adam_step(*args, 1, h)
            st.step += 1
            # the fused kernel writes the parameters through raw pointers: tell torch (and the
            # generator's packed-weight cache, engine.py::_params_signature) that they changed
            torch.autograd.graph.increment_version([p for p in st.params if p.grad is not None])
        return loss

    # ------------------------------------------------------------------ checkpoint layout
    def state_dict(self):
        """Same nesting as the reference: per same-shape batch, the state sits under the first
        parameter of the batch and holds stacked tensors (optim.py:84-101)."""
        self._export_state()
        return super().state_dict()

    def _export_state(self):
        for gi, group in enumerate(self.param_groups):
            st = self._gs[gi]
            if st is None or st.step == 0:
                continue
            period = group["size_update_period"]
            slot = 0
            for bi, b in enumerate(st.batches):
                p0 = st.params[b[0]]
                nb = len(b)
                ts = st.tstate[slot:slot + nb]
                ones = [1] * (p0.dim())
                d = {"step": st.step, "exp_avg_sq": st.exp_avg_sq[bi], "delta": st.delta[bi]}
                if p0.numel() != 1:
                    d["param_rms"] = ts[:, 0].reshape(nb, *ones).clone()
                    d["scale_exp_avg_sq"] = ts[:, 1].reshape(nb, *ones).clone()
                    d["scale_grads"] = ts[:, 2:2 + period].t().reshape(period, nb, *ones).clone()
                if bi == 0 and group["clipping_scale"] is not None:
                    d["model_norms"] = st.norms.clone()
                    if st.threshold is not None:
                        d["model_norm_threshold"] = st.threshold
                    d["num_clipped"] = int(st.gstate[3])
                self.state[p0] = d
                slot += nb

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for gi, group in enumerate(self.param_groups):
            self._gs[gi] = None
            st = self._group_state(gi)
            period = group["size_update_period"]
            slot = 0
            for bi, b in enumerate(st.batches):
                d = self.state.get(st.params[b[0]], {})
                nb = len(b)
                if "step" in d:
                    st.step = int(d["step"])
                    st.exp_avg_sq[bi].copy_(d["exp_avg_sq"])
                    st.delta[bi].copy_(d["delta"])
                    if "param_rms" in d:
                        st.tstate[slot:slot + nb, 0] = d["param_rms"].reshape(nb)
                        st.tstate[slot:slot + nb, 1] = d["scale_exp_avg_sq"].reshape(nb)
                        st.tstate[slot:slot + nb, 2:2 + period] = d["scale_grads"].reshape(period, nb).t()
                    if "model_norms" in d:
                        st.norms.copy_(d["model_norms"])
                    if "model_norm_threshold" in d:
                        st.threshold = float(d["model_norm_threshold"])
                        st.gstate[2] = st.threshold
                    if "num_clipped" in d:
                        st.gstate[3] = float(d["num_clipped"])
                slot += nb


class LRScheduler(object):
    """Batch/epoch-indexed scheduler base (flow2gan/optim.py:741-849)."""

    def __init__(self, optimizer: Optimizer, verbose: bool = False):
        if not isinstance(optimizer, Optimizer):
            raise TypeError("{} is not an Optimizer".format(type(optimizer).__name__))
        self.optimizer = optimizer
        self.verbose = verbose
        for group in optimizer.param_groups:
            group.setdefault("base_lr", group["lr"])
        self.base_lrs = [group["base_lr"] for group in optimizer.param_groups]
        self.epoch = 0
        self.batch = 0

    def state_dict(self):
        return {"epoch": self.epoch, "batch": self.batch}

    def load_state_dict(self, state_dict):
        base_lrs = self.base_lrs
        self.__dict__.update(state_dict)
        self.base_lrs = base_lrs

    def get_last_lr(self) -> List[float]:
        return self._last_lr

    def get_lr(self):
        raise NotImplementedError

    def step_batch(self, batch: Optional[int] = None) -> None:
        self.batch = batch if batch is not None else self.batch + 1
        self._set_lrs()

    def step_epoch(self, epoch: Optional[int] = None):
        self.epoch = epoch if epoch is not None else self.epoch + 1
        self._set_lrs()

    def _set_lrs(self):
        values = self.get_lr()
        assert len(values) == len(self.optimizer.param_groups)
        for group, lr in zip(self.optimizer.param_groups, values):
            group["lr"] = lr
        self._last_lr = [group["lr"] for group in self.optimizer.param_groups]


class Eden2(LRScheduler):
    """lr = base_lr * ((batch^2 + lr_batches^2) / lr_batches^2)^-0.5 * warmup  (optim.py:904-951)."""

    def __init__(self, optimizer: Optimizer, lr_batches: Union[int, float],
                 warmup_batches: Union[int, float] = 500.0, warmup_start: float = 0.5,
                 verbose: bool = False):
        super().__init__(optimizer, verbose)
        self.lr_batches = lr_batches
        self.warmup_batches = warmup_batches
        assert 0.0 <= warmup_start <= 1.0, warmup_start
        self.warmup_start = warmup_start

    def get_lr(self):
        factor = ((self.batch ** 2 + self.lr_batches ** 2) / self.lr_batches ** 2) ** -0.5
        warmup = (1.0 if self.batch >= self.warmup_batches
                  else self.warmup_start + (1.0 - self.warmup_start) * (self.batch / self.warmup_batches))
        return [x * factor * warmup for x in self.base_lrs]
