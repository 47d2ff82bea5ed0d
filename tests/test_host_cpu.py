"""CPU-only checks of the host layer: the C-ABI library loads and exports every symbol that
include/flow2gan_b200.h declares, module trees / configs match the reference layout, and the
product path refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

from _cases import GOLDEN, ROOT
from oracle import flow2gan_oracle as O


def test_abi_library_loads_and_exports_all_declared_symbols():
    from flow2gan_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "flow2gan_b200.h")).read()
    declared = set(re.findall(r"\b(f2g_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/flow2gan_b200.h but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.f2g_abi_version() == 5


def test_state_dict_layout_matches_reference_spec():
    from flow2gan_b200 import get_generator_config
    from flow2gan_b200.generator import MelAudioGenerator
    for tag, name in (("24k", "mel_24k_base"), ("44k", "mel_44k_128band_512x_base")):
        g = torch.load(os.path.join(GOLDEN, f"ref_infer_{tag}.pt"), weights_only=False)
        m = MelAudioGenerator(**get_generator_config(name))
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert mine == dict(g["sd_spec"])
        cfg, ocfg = get_generator_config(name), O.generator_config(name)
        for k, v in ocfg.items():
            assert cfg[k] == v, k
        # buffers equal the oracle's own construction
        assert torch.allclose(m.estimators[0].fft.window, O.hann(cfg["n_ffts"][0]), atol=2e-7)
        assert torch.allclose(m.loss_spec.fb, O.linear_fbanks(cfg["loss_n_fft"] // 2 + 1, 256,
                                                              cfg["sampling_rate"]))


def test_get_model_surface(tmp_path):
    import flow2gan_b200 as F
    with pytest.raises(ValueError):
        F.get_generator_config("nope")
    with pytest.raises(AssertionError):
        F.get_model("mel_24k_base", hf_model_name=None, checkpoint=None)
    from flow2gan_b200.generator import MelAudioGenerator
    m = MelAudioGenerator(**F.get_generator_config("mel_24k_base"))
    ck = tmp_path / "g.pt"
    torch.save({"model": {"module." + k: v for k, v in m.state_dict().items()}, "extra": 1}, ck)
    m2, cfg = F.get_model("mel_24k_base", checkpoint=str(ck))
    assert cfg.mel_hop_length == 256
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less box")
def test_no_cpu_fallback():
    import flow2gan_b200 as F
    from flow2gan_b200.generator import MelAudioGenerator
    m = MelAudioGenerator(**F.get_generator_config("mel_24k_base")).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        m.infer(torch.zeros(1, 100, 8))
    from flow2gan_b200.modules import LogMelSpectrogram
    with pytest.raises((RuntimeError, AssertionError)):
        LogMelSpectrogram()(torch.zeros(1, 4000))
