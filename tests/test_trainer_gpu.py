"""GANTrainer (finetune.py:427-492,590-626 as a class): the whole-phase CUDA-graph path must follow
the eager path step for step, and optimizer updates must reach the generator's packed weights."""
import random

import pytest
import torch

from _cases import audio_input, rel_rms
from _synth import synth_state_dict

pytestmark = pytest.mark.gpu
B, T = 2, 8192


def _gan(seed=4321):
    from flow2gan_b200 import get_gan_config, get_generator_config
    from flow2gan_b200.gan import GAN
    from flow2gan_b200.generator import MelAudioGenerator
    torch.manual_seed(0)
    gen = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    gen.branch_dropout = 0.0                                   # finetune.py:414
    gan = GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
    spec = [(k, tuple(v.shape)) for k, v in gan.state_dict().items()]
    gan.load_state_dict(synth_state_dict(spec, seed), strict=False)
    return gan.cuda()


def _run(use_graph, steps):
    from flow2gan_b200.trainer import GANTrainer
    gan = _gan()
    tr = GANTrainer(gan, use_graph=use_graph, graph_warmup=1)
    audio = audio_input(B, T, seed=5).cuda()
    lens = torch.full((B,), T, device="cuda", dtype=torch.int64)
    torch.manual_seed(123)
    random.seed(321)
    losses = []
    for _ in range(steps):
        info = tr.step(audio, lens)
        losses.append(float(info.get("disc_loss", info.get("gen_loss"))))
    torch.cuda.synchronize()
    return gan, tr, losses


def test_step_graph_follows_eager():
    steps = 8                                   # 4 D + 4 G iterations; phases 3 and 4 replay graphs
    gan_e, _, le = _run(False, steps)
    gan_g, tr, lg = _run(True, steps)
    assert tr.use_graph and sum("graph" in e for e in tr._graphs.values()) == 2, "capture fell back to eager"
    print("eager ", ["%.5f" % v for v in le])
    print("graphs", ["%.5f" % v for v in lg])
    for a, b in zip(le, lg):
        assert abs(a - b) <= 3e-3 * max(1.0, abs(a)), (le, lg)
    pe, pg = dict(gan_e.named_parameters()), dict(gan_g.named_parameters())
    worst = max(rel_rms(pg[k].detach().cpu(), pe[k].detach().cpu()) for k in pe if pe[k].numel() > 64)
    print("worst parameter rel-RMS after %d steps: %.2e" % (steps, worst))
    assert worst < 2e-2


def test_optimizer_update_invalidates_packed_weights():
    """ScaledAdam writes parameters through raw pointers; the generator's TF32-packed copies
    (engine.PackedGenerator) must notice and be rebuilt before the next forward."""
    from flow2gan_b200.trainer import GANTrainer
    gan = _gan()
    gen = gan.generator
    tr = GANTrainer(gan, use_graph=False)
    audio = audio_input(B, T, seed=5).cuda()
    lens = torch.full((B,), T, device="cuda", dtype=torch.int64)
    gen.eval()
    mel = tr.cond_module(audio)
    noise = torch.randn(B, mel.shape[2] * gen.mel_hop_length, device="cuda") * 0.1
    with torch.no_grad():
        y0 = gen.infer(mel, noise=noise).clone()
    tr.step(audio, lens)                        # D iteration: generator untouched
    tr.step(audio, lens)                        # G iteration: generator parameters move
    gen.eval()
    with torch.no_grad():
        y1 = gen.infer(mel, noise=noise).clone()
    fresh = _gan().generator
    fresh.load_state_dict(gen.state_dict())
    fresh = fresh.cuda().eval()
    with torch.no_grad():
        y2 = fresh.infer(mel, noise=noise)
    assert rel_rms(y1.cpu(), y0.cpu()) > 1e-4, "the optimizer step did not change the output"
    assert rel_rms(y1.cpu(), y2.cpu()) < 1e-6, "stale packed weights after an optimizer step"


def test_graphs_of_two_batch_shapes_alternate_like_eager():
    """Batch shapes vary in practice (the collate drops silent items; the last batch of an epoch is
    partial): every (phase, shape) gets its own graph, and graphs replayed in any order -- with eager
    iterations in between -- must hand the optimizer the gradients of THAT iteration (the captured
    `.grad` tensors are re-attached after each replay) computed on current packed weights."""
    from flow2gan_b200.trainer import GANTrainer
    shapes = [(2, 8192), (3, 6144)]
    order = [0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 1, 1, 0]       # (D, G) pairs per shape, then mixed pairs

    def run(use_graph):
        gan = _gan()
        tr = GANTrainer(gan, use_graph=use_graph, graph_warmup=1)
        torch.manual_seed(123)
        random.seed(321)
        losses = []
        for it, si in enumerate(order):
            b, t = shapes[si]
            audio = audio_input(b, t, seed=50 + si).cuda()
            lens = torch.full((b,), t, device="cuda", dtype=torch.int64)
            info = tr.step(audio, lens)
            losses.append(float(info.get("disc_loss", info.get("gen_loss"))))
        torch.cuda.synchronize()
        return gan, tr, losses

    gan_e, _, le = run(False)
    gan_g, tr, lg = run(True)
    n_graphs = sum("graph" in e for e in tr._graphs.values())
    assert tr.use_graph and n_graphs >= 3, n_graphs
    print("eager ", ["%.5f" % v for v in le])
    print("graphs", ["%.5f" % v for v in lg])
    for a, b in zip(le, lg):
        assert abs(a - b) <= 5e-3 * max(1.0, abs(a)), (le, lg)
    pe, pg = dict(gan_e.named_parameters()), dict(gan_g.named_parameters())
    worst = max(rel_rms(pg[k].detach().cpu(), pe[k].detach().cpu()) for k in pe if pe[k].numel() > 64)
    print("worst parameter rel-RMS after %d mixed-shape steps: %.2e" % (len(order), worst))
    assert worst < 3e-2
