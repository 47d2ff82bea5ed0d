"""One GAN D+G iteration pair at the bench shape between cudaProfilerStart/Stop (for ncu launch lists)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from _cases import audio_input
from flow2gan_b200 import get_gan_config, get_generator_config
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
tr = GANTrainer(gan)
audio = audio_input(bench.B, 24000, seed=2).to(dev)
lens = torch.full((bench.B,), 24000, device=dev, dtype=torch.int64)
for _ in range(6): tr.step(audio, lens)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2): tr.step(audio, lens)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("train pair: %.1f ms" % e0.elapsed_time(e1))
