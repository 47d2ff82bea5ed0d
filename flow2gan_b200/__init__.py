"""flow2gan_b200: B200-native (sm_100a) implementation of the Flow2GAN hot path behind the
reference's Python surface (flow2gan/__init__.py:29-47)."""
from __future__ import annotations

from typing import Optional, Tuple

from .config import HF_MODEL_NAMES, HF_REPO, AttributeDict, get_gan_config, get_generator_config

__all__ = ["get_model", "load_checkpoint", "get_generator_config", "get_gan_config",
           "AttributeDict", "MelAudioGenerator", "LogMelSpectrogram"]


def __getattr__(name):
    if name == "MelAudioGenerator":
        from .generator import MelAudioGenerator
        return MelAudioGenerator
    if name == "LogMelSpectrogram":
        from .modules import LogMelSpectrogram
        return LogMelSpectrogram
    raise AttributeError(name)


def load_checkpoint(filename, model, *args, **kwargs) -> dict:
    """flow2gan.checkpoint.load_checkpoint (see flow2gan_b200/checkpoint.py)."""
    from .checkpoint import load_checkpoint as _load
    return _load(filename, model, *args, **kwargs)


def get_model(model_name: str = "mel_24k_base", hf_model_name: Optional[str] = "libritts-mel-4-step",
              checkpoint: Optional[str] = None) -> Tuple["MelAudioGenerator", AttributeDict]:
    assert (checkpoint is not None) or (hf_model_name is not None), \
        "Either checkpoint or hf_model_name must be provided."
    from .generator import MelAudioGenerator
    model_cfg = get_generator_config(model_name)
    model = MelAudioGenerator(**model_cfg)
    if checkpoint is not None:
        print(f"Using local checkpoint: {checkpoint}")
    else:
        print("Using checkpoint from HF hub")
        assert hf_model_name in HF_MODEL_NAMES, "Supported names are " + ", ".join(HF_MODEL_NAMES.keys())
        from huggingface_hub import hf_hub_download
        checkpoint = hf_hub_download(HF_REPO, filename=hf_model_name + ".pt")
    load_checkpoint(checkpoint, model)
    return model, model_cfg
