"""f3: the loss-side S x S / Gram products on the tcgen05 GEMM (fmi_gemm_nt through ops.bmm_nt, 3xTF32 operands)
against the oracle (oracle/ref_ops.py: gram_matrix / style_loss / contextual_loss, restating
modules/pluralistic_model/external_function.py:180-192, 231-274) and its golden recorded from the reference's own functions.
fp32 contract: max|a-b|/max|b| <= 1e-3 (measured ~1e-6: the split operands give fp32-class products)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(GOLD / name).items()}


@pytest.mark.parametrize("shape", [(2, 64, 40, 128), (4, 1024, 1024, 512), (1, 8, 8, 8), (3, 136, 72, 264), (2, 64, 64, 16384)])
def test_bmm_nt_forward_backward(shape):
    from face_mask_inpaint_b200 import _lib, ops
    bs, m, n, k = shape
    g = torch.Generator().manual_seed(m + n)
    a = torch.randn(bs, m, k, generator=g).to(DEV)
    b = torch.randn(bs, n, k, generator=g).to(DEV)
    gc = torch.randn(bs, m, n, generator=g).to(DEV)
    assert ops.bmm_nt_supported(a, b)

    def run(fn, dt):
        a_, b_ = a.detach().clone().to(dt).requires_grad_(True), b.detach().clone().to(dt).requires_grad_(True)
        c = fn(a_, b_)
        c.backward(gc.to(dt))
        return c.detach(), a_.grad, b_.grad

    want = run(lambda x, y: torch.bmm(x, y.transpose(1, 2)), torch.float64)
    n0 = _lib.load().fmi_kernel_launch_count()
    got = run(ops.bmm_nt, torch.float32)
    torch.cuda.synchronize()
    assert _lib.load().fmi_kernel_launch_count() - n0 >= 9        # 3 GEMMs, 2 operand splits each
    assert not ops.bmm_nt_supported(a[:, :, :4], b[:, :, :4])      # K must be a multiple of 8 (it is the N of the backward GEMMs)
    for name, x, r in zip(("c", "da", "db"), got, want):
        assert rel_err(x, r) <= 2e-5, (name, rel_err(x, r))


def test_gram_and_contextual_loss_match_the_reference_golden():
    from face_mask_inpaint_b200 import ops
    g = load("loss_side.npz")
    x, y = g["x"].to(DEV), g["y"].to(DEV)
    assert rel_err(ops.gram_matrix(x), g["gram"]) <= 1e-5
    assert rel_err(ops.contextual_loss(x, y), g["cx"]) <= 1e-4
    assert rel_err(ops.contextual_loss(y, x, h=1.0), g["cx_h1"]) <= 1e-4


@pytest.mark.parametrize("shape", [(4, 512, 32, 32), (2, 256, 28, 28)])
def test_contextual_and_style_loss_gradients(shape):
    """VGG-scale features (train_reference_fill: relu4_1 of a 256^2 image is [N,512,32,32]): value and gradient vs the oracle in fp64."""
    from face_mask_inpaint_b200 import ops
    g = torch.Generator().manual_seed(shape[1])
    x = torch.relu(torch.randn(*shape, generator=g)).to(DEV)
    y = torch.relu(x.cpu() + 0.7 * torch.randn(*shape, generator=g)).to(DEV)

    def run(fn, dt):
        x_ = x.detach().clone().to(dt).requires_grad_(True)
        loss = fn(x_, y.to(dt))
        loss.backward()
        return loss.detach(), x_.grad

    for ours, oracle in ((ops.contextual_loss, O.contextual_loss),
                         (lambda a, b: torch.nn.functional.l1_loss(ops.gram_matrix(a), ops.gram_matrix(b).detach()), O.style_loss)):
        want = run(oracle, torch.float64)
        got = run(ours, torch.float32)
        assert rel_err(got[0], want[0]) <= 1e-4, rel_err(got[0], want[0])
        assert rel_err(got[1], want[1]) <= 1e-3, rel_err(got[1], want[1])
