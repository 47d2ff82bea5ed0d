"""Data-parallel gradient exchange on real GPUs (needs >= 2): the NCCL all-reduce of
GradBuckets.allreduce_mean (dist.py; replaces the reference's DDP wrapper, bin/finetune.py:913-915) over
two ranks that each ran one GAN phase on HALF of a batch must reproduce the gradient a single rank
computes on the concatenated batch (every loss is a batch mean), for both phases; and differently
seeded replicas must leave GANTrainer construction with rank 0's parameters (DDP's initial broadcast)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from _cases import audio_input, noise_input, rel_rms
from _synth import synth_state_dict

pytestmark = pytest.mark.gpu
B, T = 4, 8192


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gan(seed=4321):
    from flow2gan_b200 import get_gan_config, get_generator_config
    from flow2gan_b200.gan import GAN
    from flow2gan_b200.generator import MelAudioGenerator
    torch.manual_seed(0)
    gen = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    gen.branch_dropout = 0.0
    gan = GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
    gan.load_state_dict(synth_state_dict([(k, tuple(v.shape)) for k, v in gan.state_dict().items()], seed), strict=False)
    return gan


def _phase_grads(gan, audio, noise, train_disc):
    import random
    from flow2gan_b200.modules import LogMelSpectrogram
    dev = audio.device
    mel = LogMelSpectrogram(24000, 1024, 256, 100).to(dev)(audio)
    lens = torch.full((audio.shape[0],), audio.shape[1], device=dev, dtype=torch.int64)
    rr, random.random = random.random, (lambda: 0.99)          # LimitParamValue flips off (host coin, modules.py:267)
    try:
        gan.zero_grad()
        losses = gan(cond=mel, audio=audio, audio_lens=lens, n_timesteps=1, train_disc=train_disc, noise=noise)
        w = (1.0, 0.1) if train_disc else (1.0, 0.1, 1.0, 0.1, 45.0)
        sum(l * wi for l, wi in zip(losses, w)).backward()
    finally:
        random.random = rr
    return gan.discriminator if train_disc else gan.generator


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from flow2gan_b200.dist import GradBuckets, cleanup_dist, setup_dist
    from flow2gan_b200.trainer import GANTrainer
    setup_dist(rank, world, backend="nccl")
    dev = torch.device("cuda", rank)
    # (1) construction-time sync: rank 1 starts from different weights
    gan = _gan(4321 if rank == 0 else 999).to(dev)
    GANTrainer(gan, use_graph=False)
    chk = torch.stack([p.detach().double().sum() for p in gan.parameters()]).sum()
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    same = bool(both[0] == both[1])
    # (2) gradient exchange: each rank runs its half of the batch
    audio = audio_input(B, T, seed=5)[rank * (B // world):(rank + 1) * (B // world)].to(dev)
    noise = noise_input(B, T, seed=6)[rank * (B // world):(rank + 1) * (B // world)].to(dev)
    out = {}
    for train_disc in (True, False):
        sub = _phase_grads(gan, audio, noise, train_disc)
        nbytes = GradBuckets(sub.parameters()).allreduce_mean()
        if rank == 0:
            # numpy arrays travel through the queue by value (torch tensors would be passed as shared-memory
            # handles that die with this process)
            out["d" if train_disc else "g"] = ({k: p.grad.detach().cpu().numpy() for k, p in sub.named_parameters()
                                                if p.grad is not None}, nbytes)
    if rank == 0:
        q.put((same, out))
    dist.barrier()
    cleanup_dist()
    q.close()
    q.join_thread()            # flush the feeder thread before the process exits


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_grad_allreduce_equals_single_rank_on_concatenated_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        same, got = q.get(timeout=600)
    finally:
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    assert same, "ranks left GANTrainer construction with different parameters"
    gan = _gan(4321).cuda()
    audio, noise = audio_input(B, T, seed=5).cuda(), noise_input(B, T, seed=6).cuda()
    for ph, train_disc in (("d", True), ("g", False)):
        sub = _phase_grads(gan, audio, noise, train_disc)
        ref = {k: p.grad.detach().cpu() for k, p in sub.named_parameters() if p.grad is not None}
        grads, nbytes = got[ph]
        grads = {k: torch.from_numpy(v) for k, v in grads.items()}
        assert nbytes == sum(p.numel() for p in sub.parameters() if p.requires_grad) * 4
        assert set(grads) == set(ref)
        num = sum(float((grads[k] - ref[k]).double().pow(2).sum()) for k in ref)
        den = sum(float(ref[k].double().pow(2).sum()) for k in ref)
        errs = sorted(rel_rms(grads[k], ref[k]) for k in ref if float(ref[k].abs().max()) > 0)
        print(ph, "2-rank NCCL mean vs 1-rank concatenated batch: whole-vector %.2e median %.2e max %.2e"
              % ((num / den) ** 0.5, errs[len(errs) // 2], errs[-1]))
        # same kernels, same TF32 products; only batch-dependent tiling / summation order differs
        assert (num / den) ** 0.5 < 5e-3 and errs[len(errs) // 2] < 5e-3
