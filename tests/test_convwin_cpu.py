"""Host-side geometry of the windowed (implicit-im2col) Conv2d (flow2gan_b200/convwin.py) without a
GPU: the C-ABI calls it makes are replaced by a numpy restatement of the operand addressing
documented in include/flow2gan_b200.h (tools/convwin_cpu_emul.py), and forward + all three
gradients are compared with torch.nn.functional.conv2d (discriminators.py:65-76,171-184 shapes in
miniature).  The CUDA kernels themselves are covered by tests/test_gan_gpu.py."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def emul():
    from flow2gan_b200 import _lib as L
    saved = {k: getattr(L, k) for k in ("pad2d", "conv_w_pack", "pack2d", "act_bwd", "act_bwd_win", "conv_w_pack_dgrad", "gemm_group", "ptr")}
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        mod = importlib.import_module("convwin_cpu_emul")
        importlib.reload(mod)
        yield mod
    finally:
        for k, v in saved.items():
            setattr(L, k, v)
        sys.path.remove(os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("case", [
    (2, 4, 9, 32, 32, 3, 9, 2, 1, 4, 0.1),       # DiscriminatorR (3,9) stride (1,2)
    (2, 3, 6, 32, 1, 3, 3, 1, 1, 1, None),       # conv_post (3,3), single output channel
    (2, 1, 14, 32, 40, 1, 5, 3, 0, 2, 0.1),      # DiscriminatorP (5,1) stride 3, period-major
])
def test_windowed_conv_geometry_matches_conv2d(emul, case):
    emul.check(*case)


def test_pick_split_k_fills_waves():
    from flow2gan_b200._lib import pick_split_k
    assert pick_split_k(256, 256, 64) == 1                      # nothing to split
    s = pick_split_k(5120, 1024, 10912)                         # 80 tiles: 2 half-empty waves unsplit
    assert 4 <= s <= 12
    tiles = 20 * 4
    assert (tiles * s) % 74 <= 74 and (tiles * s + 73) // 74 * (341 // s + 8) < 2 * (341 + 8)
    assert pick_split_k(864, 32, 218880) >= 16                  # skinny wgrad: split over all pairs
    for args in ((384, 1152, 6016), (768, 2304, 1504), (32, 54, 385024)):
        s = pick_split_k(*args)
        assert 1 <= s <= 64 and (args[2] + 31) // 32 // s >= 8  # never fewer than 8 k-blocks per tile
