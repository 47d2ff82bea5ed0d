"""Pairwise rel-RMS between inference builds (one interpreter per switch setting) and against the fp32 CPU
oracle, at the bench shape (bs 16, 2 ODE steps) -- attributes the fp16-vs-TF32 operand difference seen by
tests/test_zz_experiments_gpu.py (5.4e-4) to a source.    python tools/arm_matrix.py  (on a B200)"""
import itertools, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ARM = r'''
import sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from _cases import mel_input, noise_input
from _synth import synth_state_dict
from flow2gan_b200 import get_generator_config
from flow2gan_b200.generator import MelAudioGenerator
torch.manual_seed(0)
m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
m.load_state_dict(synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 99), strict=False)
m = m.cuda().eval()
outs = []
with torch.no_grad():
    for rep in range(3):      # eager, graph capture, graph replay
        outs.append(m.infer(mel_input(16, 100, 94, seed=0).cuda(), n_timesteps=2, noise=noise_input(16, 24064, seed=1).cuda()).cpu())
torch.save(outs, {out!r})
'''
arms = {"default": {}, "default_again": {}, "tf32": {"F2G_BLOCK_OPERANDS": "tf32"}, "unchained": {"F2G_CHAIN_MLP": "0"},
        "nofork": {"F2G_FORK_COND": "0"}, "nopdl": {"F2G_PDL": "0"}, "nocache": {"F2G_CACHE_TIME": "0"}}
res = {}
for name, env in arms.items():
    out = f"/tmp/arm_{name}.pt"
    r = subprocess.run([sys.executable, "-c", ARM.format(root=ROOT, out=out)], env=dict(os.environ, **env), capture_output=True, text=True)
    if r.returncode:
        print(name, "FAILED", r.stderr[-500:]); continue
    res[name] = torch.load(out)
def rel(a, b):
    return float((a - b).double().pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt())
for name, o in res.items():
    print(f"{name:14s} eager-vs-capture-call {rel(o[0], o[1]):.3e}  capture-call-vs-replay {rel(o[1], o[2]):.3e}")
names = list(res)
print("pairwise rel-RMS (first, eager call):")
for a, b in itertools.combinations(names, 2):
    print(f"  {a:14s} {b:14s} {rel(res[a][0], res[b][0]):.3e}")
if "--oracle" in sys.argv:
    from _cases import mel_input, noise_input
    from _synth import synth_state_dict
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    from oracle import flow2gan_oracle as O
    m = MelAudioGenerator(**get_generator_config("mel_24k_base"))
    sd = synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 99)
    torch.set_num_threads(16)
    with torch.no_grad():
        ref = O.generator_infer(sd, O.generator_config("mel_24k_base"), mel_input(16, 100, 94, seed=0), noise_input(16, 24064, seed=1), None, 2, False)
    for name, o in res.items():
        print(f"  {name:14s} vs fp32 CPU oracle {rel(o[0], ref):.3e}")
