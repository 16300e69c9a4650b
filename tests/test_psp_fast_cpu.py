"""Host-side arithmetic of the pSp-encoder kernel path (modules/psp_fast.py) checked on CPU against the IR-SE unit itself
(encoders/helpers.py:97-119 as mirrored in modules/psp.py): eval-mode BatchNorm folded into weights, the border-class bias of
the BatchNorm that precedes a zero-padded conv, the tap layout [ky*3+kx][O][I], the parity-plane formulation of the stride-2
conv and the strided shortcut. The kernels themselves are checked on the GPU (tests/test_ir_encoder_gpu.py)."""
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from face_mask_inpaint_b200.modules import psp as P  # noqa: E402
from face_mask_inpaint_b200.modules import psp_fast as PF  # noqa: E402


def _randomize_bn(m, g):
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.5)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
            mod.weight.data.copy_(1 + 0.3 * torch.randn(mod.weight.shape, generator=g))
            mod.bias.data.copy_(0.3 * torch.randn(mod.bias.shape, generator=g))


def _conv_from_taps(x, wp, bias_cls=None, stride=1):
    """What fmi_conv_nhwc computes, restated with F.conv2d: wp [9 or 1][O][I] -> Conv2d weight; bias [O] or border classes [9][O]."""
    t, o, i = wp.shape
    k = int(t ** 0.5)
    w = wp.reshape(k, k, o, i).permute(2, 3, 0, 1).float()
    y = F.conv2d(x, w, None, stride, k // 2)
    if bias_cls is None:
        return y
    if bias_cls.dim() == 1:
        return y + bias_cls.view(1, -1, 1, 1)
    h, wd = y.shape[-2:]
    vy = torch.ones(h, dtype=torch.long); vy[0] = 0; vy[-1] = 2
    vx = torch.ones(wd, dtype=torch.long); vx[0] = 0; vx[-1] = 2
    cls = vy.view(-1, 1) * 3 + vx.view(1, -1)                    # [h, w]
    return y + bias_cls[cls].permute(2, 0, 1).unsqueeze(0)


def _planes_conv(a, wp, bias):
    """The stride-2 3x3 conv as fmi_conv_nhwc(planes=1) runs it: parity planes, tap (dy,dx) reads plane (dy&1, dx&1) at shift
    (dy<0 ? -1 : 0, dx<0 ? -1 : 0) with zero fill."""
    n, c, h, w = a.shape
    planes = [[a[:, :, py::2, px::2] for px in range(2)] for py in range(2)]
    out = 0
    for t in range(9):
        dy, dx = t // 3 - 1, t % 3 - 1
        pl = planes[dy & 1][dx & 1]
        sy, sx = (-1 if dy < 0 else 0), (-1 if dx < 0 else 0)
        sh = F.pad(pl, (1, 0, 1, 0))[:, :, 1 + sy:1 + sy + h // 2, 1 + sx:1 + sx + w // 2]
        out = out + torch.einsum("oi,nihw->nohw", wp[t].float(), sh)
    return out + bias.view(1, -1, 1, 1)


@pytest.mark.parametrize("in_c,depth,stride,se", [(16, 16, 1, True), (16, 32, 2, True), (16, 16, 2, True), (32, 32, 1, False)])
def test_folded_unit_equals_ir_unit(monkeypatch, in_c, depth, stride, se):
    monkeypatch.setattr(PF, "_operand", lambda w, mma: w.contiguous())       # exact fp32 operands: this test is about the algebra
    g = torch.Generator().manual_seed(3)
    torch.manual_seed(4)
    unit = P._IRUnit(in_c, depth, stride, se).eval()
    _randomize_bn(unit, g)
    x = torch.randn(2, in_c, 12, 10, generator=g)
    with torch.no_grad():
        want = unit(x)
        u = PF._prep_unit(unit, 0)
        a1 = _conv_from_taps(x, u.w1, u.b1)
        a1 = torch.where(a1 > 0, a1, a1 * u.slope.view(1, -1, 1, 1))
        r = _planes_conv(a1, u.w2, u.b2) if stride == 2 else _conv_from_taps(a1, u.w2, u.b2)
        xs = x[:, :, ::stride, ::stride]
        sc = xs if u.ws is None else _conv_from_taps(xs, u.ws, u.bs)
        if u.se1 is not None:
            m = r.mean(dim=(2, 3))
            gate = torch.sigmoid(F.linear(torch.relu(F.linear(m, u.se1)), u.se2))
            r = r * gate.view(*gate.shape, 1, 1)
        got = r + sc
    assert got.shape == want.shape
    assert ((got - want).abs().max() / want.abs().max()).item() < 2e-5


def test_tf32_operand_rounding_matches_cvt_rna():
    w = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -1.0 - 2 ** -11, 3.1415927, 1e-30])
    r = PF._operand(w, 0)          # _lib.MMA_TF32 == 0
    assert (r.view(torch.int32) & 0x1FFF).abs().max().item() == 0                 # 13 low mantissa bits cleared
    assert r[1].item() == 1.0 + 2 ** -10 and r[3].item() == -1.0 - 2 ** -10        # ties away from zero
    assert ((r - w).abs() <= w.abs() * 2 ** -11 + 1e-38).all()
