"""GEMM census of one GAN D+G iteration pair (eager): every grouped launch timed with CUDA events,
aggregated by shape -> where the tensor time of the training step goes.
python tools/train_gemm_census.py"""
import os, sys, collections, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from _cases import audio_input
from flow2gan_b200 import _lib as L, get_gan_config, get_generator_config
from flow2gan_b200.gan import GAN
from flow2gan_b200.generator import MelAudioGenerator
from flow2gan_b200.trainer import GANTrainer
from _synth import synth_state_dict
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gen = MelAudioGenerator(**get_generator_config(bench.MODEL)); gen.branch_dropout = 0.0
gan = GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
gan.load_state_dict(synth_state_dict([(k, tuple(v.shape)) for k, v in gan.state_dict().items()], 4321), strict=False)
gan = gan.to(dev)
tr = GANTrainer(gan, use_graph=False)
audio = audio_input(bench.B, 24000, seed=2).to(dev)
lens = torch.full((bench.B,), 24000, device=dev, dtype=torch.int64)
for _ in range(2): tr.step(audio, lens)
torch.cuda.synchronize()
rec = []
orig = L.gemm_group


def hooked(descs):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(descs); e1.record()
    rec.append((e0, e1, [(d.M, d.N, d.K, d.a_mn, d.b_mn, d.split_k, int(d.a_seg_len != 0), d.ab_f16) for d in descs], phase[0]))


L.gemm_group = hooked
import flow2gan_b200.convwin, flow2gan_b200.discriminators, flow2gan_b200.train, flow2gan_b200.engine
phase = ["D"]
t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
t0.record(); tr.step(audio, lens); t1.record(); phase[0] = "G"; tr.step(audio, lens); t2.record()
torch.cuda.synchronize()
print("eager D %.1f ms  G %.1f ms" % (t0.elapsed_time(t1), t1.elapsed_time(t2)))
agg = collections.OrderedDict()
tot = 0.0
for e0, e1, shapes, ph in rec:
    us = e0.elapsed_time(e1) * 1e3
    fl = sum(2.0 * m * n * k for m, n, k, *_ in shapes)
    key = (ph, len(shapes)) + shapes[0]
    a = agg.setdefault(key, [0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += fl
    tot += us
print("%d launches, %.1f ms in GEMM launches (eager, includes launch gaps)" % (len(rec), tot / 1e3))
print("phase grp        M      N      K amn bmn spl win f16 | calls   total_us  avg_us   TF/s")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    ph, n, M, N, K, am, bm, sp, win, f16 = k
    print("%3s %3d %9d %6d %6d %3d %3d %3d %3d %3d | %5d %10.0f %7.1f %6.0f" % (ph, n, M, N, K, am, bm, sp, win, f16, v[0], v[1], v[1] / v[0], v[2] / v[1] / 1e6))
