"""CPU checks of the fused multi-tensor loss reductions: csrc/losses.cu compiled for the host from the
same source (tests/_emul.py) behind the product's autograd wrappers (flow2gan_b200/losses.py,
gan.py), against torch autograd of the reference's expressions (flow2gan/models/gan.py:57-99)."""
import os
import subprocess
import sys

import pytest
import torch

import _emul
import _losses_cases as LC

pytestmark = pytest.mark.skipif(not _emul.available(), reason="g++ not available")


@pytest.fixture
def emulated_losses(monkeypatch):
    return _emul.native_fixture(monkeypatch)


def test_l1_terms(emulated_losses):
    LC.case_l1_terms("cpu")
    LC.case_l1_terms_strided_grad_flow("cpu")


def test_hinge_terms(emulated_losses):
    LC.case_hinge_terms("cpu")


def test_gan_loss_methods(emulated_losses, monkeypatch):
    LC.case_gan_loss_methods("cpu", monkeypatch)


def test_malformed_terms_are_refused(emulated_losses):
    L = emulated_losses
    x = torch.zeros(4, 4)
    t = L.loss_term(L.LOSS_L1, x, x, None)
    t.dims[0] = 3                                   # dims no longer multiply to numel
    with pytest.raises(RuntimeError, match="dims do not multiply"):
        L.loss_terms([t], False, torch.zeros(1), None)
    t = L.loss_term(L.LOSS_HINGE, x, None, None, 1.0)
    with pytest.raises(RuntimeError, match="malformed"):
        L.loss_terms([t], True, None, torch.ones(1))         # backward without a grad buffer


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less box")
def test_no_cpu_fallback():
    from flow2gan_b200.losses import hinge_terms
    with pytest.raises((RuntimeError, AssertionError)):
        hinge_terms([torch.zeros(3)], [1.0])


@pytest.mark.skipif(os.environ.get("F2G_EMUL_REVERSE") == "1", reason="already the reversed run")
def test_emulated_kernels_do_not_depend_on_thread_order():
    """The host emulation runs blocks and threads sequentially; a kernel that silently relied on that
    order would still be wrong on the GPU.  Re-run both emulated suites with blocks / threads in
    DESCENDING order (-DF2G_EMUL_REVERSE)."""
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, F2G_EMUL_REVERSE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
                        os.path.join(here, "test_datapath_cpu.py"), os.path.join(here, "test_losses_cpu.py"),
                        os.path.join(here, "test_blocks_cpu.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
