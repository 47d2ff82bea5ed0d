"""Process-group plumbing (one process per GPU, NCCL over NVLink) and the data-parallel
gradient exchange of the GAN step.  Mirrors flow2gan/dist.py:25-69 (setup_dist / cleanup_dist /
rank helpers) and replaces the reference's DDP wrapper (bin/finetune.py:913-915), which
all-reduces all 121 M parameters' gradients every iteration, by an all-reduce of the half that
is actually stepped (discriminators: 170 MB fp32, generator: 316 MB), in large flat buckets."""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def setup_dist(rank=None, world_size=None, master_port=None, use_ddp_launch=False, master_addr=None,
               backend: str = "nccl"):
    """Same call shape as the reference's setup_dist; env:// rendezvous on 127.0.0.1 by default."""
    if "MASTER_ADDR" not in os.environ:
        os.environ["MASTER_ADDR"] = "127.0.0.1" if master_addr is None else str(master_addr)
    if "MASTER_PORT" not in os.environ:
        os.environ["MASTER_PORT"] = "12354" if master_port is None else str(master_port)
    if use_ddp_launch is False:
        dist.init_process_group(backend, rank=rank, world_size=world_size)
        if backend == "nccl":
            torch.cuda.set_device(rank)
    else:
        dist.init_process_group(backend)


def cleanup_dist():
    dist.destroy_process_group()


def get_world_size() -> int:
    if "WORLD_SIZE" in os.environ:
        return int(os.environ["WORLD_SIZE"])
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank() -> int:
    if "RANK" in os.environ:
        return int(os.environ["RANK"])
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", 0))          # dist.py:63-69: defaults to 0


def broadcast_module_state(module: torch.nn.Module, src: int = 0, group=None) -> int:
    """What the reference's DDP wrapper does at construction (bin/finetune.py:913-915): every rank
    starts from rank `src`'s parameters AND buffers.  Without it, replicas that were not seeded
    identically (the freshly initialised discriminators) would train apart under averaged
    gradients without any error.  Returns the number of tensors broadcast (0 = single process)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    n = 0
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
            n += 1
    return n


class GradBuckets:
    """Flat fp32 buckets over a fixed parameter list; `allreduce_mean()` averages the gradients of
    those parameters across ranks with one collective per bucket (default 128 MB)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 128 << 20):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in self.params:
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)

    def allreduce_mean(self, group=None) -> int:
        """Returns the number of bytes exchanged per rank (payload)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return 0
        world = dist.get_world_size(group)
        total = 0
        for bi, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            flat = self._flat[bi]
            if flat is None or flat.device != bucket[0].device:
                flat = torch.empty(n, device=bucket[0].device, dtype=torch.float32)
                self._flat[bi] = flat
            # pack / unpack with one multi-tensor copy each (hundreds of small parameters per bucket)
            views, off = [], 0
            for p in bucket:
                m = p.numel()
                views.append(flat[off:off + m].view_as(p))
                off += m
            have = [i for i, p in enumerate(bucket) if p.grad is not None]
            for i, p in enumerate(bucket):
                if p.grad is None:
                    views[i].zero_()
            if have:
                torch._foreach_copy_([views[i] for i in have], [bucket[i].grad for i in have])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.mul_(1.0 / world)
            for i, p in enumerate(bucket):
                if p.grad is None:
                    p.grad = views[i].clone()
            if have:
                torch._foreach_copy_([bucket[i].grad for i in have], [views[i] for i in have])
            total += n * 4
        return total
