"""How does the CPU oracle (== the reference's PyTorch CPU path) scale with threads on this host?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from oracle import flow2gan_oracle as O
from _synth import synth_state_dict
from flow2gan_b200 import get_generator_config
from flow2gan_b200.generator import MelAudioGenerator
m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
sd = synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 99)
cfg = O.generator_config("mel_24k_base")
mel, noise = bench.synth_inputs()
print("cpu_count", os.cpu_count(), "default threads", torch.get_num_threads(), flush=True)
for th in (8, 16, 32, 64):
    torch.set_num_threads(th)
    with torch.inference_mode():
        O.generator_infer(sd, cfg, mel[:2], noise[:2], None, 1, False)
        t0 = time.perf_counter(); O.generator_infer(sd, cfg, mel, noise, None, 1, False); dt = time.perf_counter() - t0
    print(f"threads={th} {dt*1e3:.0f} ms/call {bench.SAMPLES_PER_STEP/dt:.0f} samples/s", flush=True)
    if dt > 40: break
