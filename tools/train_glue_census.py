"""Which Python call sites of flow2gan_b200 launch the torch (at::) glue kernels of one eager GAN D+G
iteration pair?  torch.profiler with stacks, kernels attributed to the innermost frame inside the
package, native (C-ABI) launches listed separately.  Guides which glue to fuse next.

    python tools/train_glue_census.py [top_n]        (on a B200; eager, so host-bound: only counts and
                                                       per-site kernel time matter, not the wall time)"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from _cases import audio_input  # noqa: E402
from _synth import synth_state_dict  # noqa: E402
from flow2gan_b200 import get_gan_config, get_generator_config  # noqa: E402
from flow2gan_b200.gan import GAN  # noqa: E402
from flow2gan_b200.generator import MelAudioGenerator  # noqa: E402
from flow2gan_b200.trainer import GANTrainer  # noqa: E402

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gen = MelAudioGenerator(**get_generator_config(bench.MODEL)); gen.branch_dropout = 0.0
gan = GAN(gen, **get_gan_config("gan_multi_scale_mel_recon"))
gan.load_state_dict(synth_state_dict([(k, tuple(v.shape)) for k, v in gan.state_dict().items()], 4321), strict=False)
gan = gan.to(dev)
tr = GANTrainer(gan, use_graph=False)
audio = audio_input(bench.B, 24000, seed=2).to(dev)
lens = torch.full((bench.B,), 24000, device=dev, dtype=torch.int64)
for _ in range(2):
    tr.step(audio, lens)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:     # stacks need verbose
    for _ in range(2):
        tr.step(audio, lens)
    torch.cuda.synchronize()

sites = collections.defaultdict(lambda: [0, 0.0, collections.Counter(), collections.Counter()])
for ev in prof.events():
    kernels = getattr(ev, "kernels", None) or []
    if not kernels:
        continue
    site = "(outside flow2gan_b200: autograd engine / torch internals)"
    for fr in (ev.stack or []):
        if "flow2gan_b200" in fr and "_lib.py" not in fr:
            site = fr.split("flow2gan_b200" + os.sep)[-1].strip()
            break
    else:
        if ev.stack:
            site = "(autograd / torch) " + str(ev.name)[:40]
    for k in kernels:
        s = sites[site]
        s[0] += 1
        s[1] += k.duration
        nm = k.name.replace("void ", "").replace("at::native::", "")
        nm = (nm.split("<")[0] + ("<" + nm.split("<")[1][:34] if "elementwise" in nm and "<" in nm else ""))[:60]
        s[2][nm] += 1
        s[3][nm] += k.duration
tot_n = sum(v[0] for v in sites.values())
tot_t = sum(v[1] for v in sites.values())
print(f"{tot_n} torch-op kernel launches, {tot_t / 1e3:.2f} ms of kernel time in one D+G pair (eager)")
for site, (n, t, names, times) in sorted(sites.items(), key=lambda x: -x[1][1])[:top_n]:
    print(f"{n:5d} {t / 1e3:8.3f} ms  {site[:90]}")
    for k, us in times.most_common(7):
        print(f"            {names[k]:5d} x {us / names[k]:7.1f} us = {us / 1e3:7.3f} ms  {k}")
