"""attention on shapes the tiles do not take directly (H*W not a multiple of 128, value channels not a
multiple of 32). ops._pad_attention_args embeds them in a shape the kernels take; here its algebra is checked on CPU: the
attention of the padded / augmented problem, restated densely as the kernel computes it (q' = Wq' x', P = softmax(q'^T q'),
O = V' P^T), equals the reference's attention on the original tensors (oracle/ref_ops.py). The kernel run itself is
tests/test_attention_gpu.py::test_padded_shapes."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from face_mask_inpaint_b200 import ops  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


def _dense_attention(xa, wa, va):
    n, c, sp, _ = xa.shape
    q = torch.einsum("dc,ncs->nds", wa.double(), xa.reshape(n, c, sp).double())
    p = torch.softmax(q.transpose(1, 2) @ q, dim=-1)                       # [n, S', S']
    return torch.einsum("ncj,nij->nci", va.reshape(n, -1, sp).double(), p)  # O[:, i] = sum_j P[i, j] V[:, j]


@pytest.mark.parametrize("h,w,c,bias", [(6, 5, 20, True), (24, 20, 48, False), (12, 11, 32, True)])
def test_padding_is_exact(h, w, c, bias):
    g = torch.Generator().manual_seed(h * w)
    n, d = 2, max(c // 4, 1)
    x = torch.randn(n, c, h, w, generator=g)
    wq = torch.randn(d, c, 1, 1, generator=g) * 0.5
    bq = torch.randn(d, generator=g) if bias else None
    assert ops._needs_padding(h * w, c, 0)
    xa, wa, va, _, _, (s, sp, c0, c1, c0p, c1p) = ops._pad_attention_args(x, wq, bq, x, None, None)
    assert sp % 128 == 0 and c0p % 32 == 0 and xa.shape == (n, c + 2, sp, 1) and va.shape == (n, c0p, sp, 1)
    got = _dense_attention(xa, wa, va)[:, :c, :s].reshape(n, c, h, w)
    want = O.auto_attn(x.double(), wq.double(), (bq if bias else torch.zeros(d)).double(), torch.ones(1).double())[0] - x.double()
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-9


def test_channel_totals_fit_the_value_tile():
    for c0, c1 in [(512, 512), (130, 130), (300, 0), (40, 24)]:
        x = torch.zeros(1, c0, 3, 3)
        v1 = torch.zeros(1, c1, 3, 3) if c1 else None
        _, _, v0a, v1a, _, meta = ops._pad_attention_args(x, torch.zeros(8, c0, 1, 1), None, x, v1, None)
        cv = meta[4] + meta[5]
        assert meta[4] % 32 == 0 and meta[5] % 32 == 0 and (cv <= 256 or cv % 256 == 0)
