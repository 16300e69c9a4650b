"""f2: the pSp-encoder kernels (csrc/ir_encoder.cu, modules/psp_fast.py) against PyTorch fp32 on the same
inputs — through the C ABI (fmi_conv_nhwc, fmi_space_to_planes_nhwc, fmi_se_gate_nhwc, fmi_se_scale_add_nhwc,
fmi_upsample_add_nhwc) and as a whole encoder against the cuDNN formulation of the same module. Tolerances: max|a-b|/max|b| <=
1e-3 with TF32 operands, <= 2e-2 with bf16 operands (north_star)."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from conftest import rel_err  # noqa: E402

TOL = {"tf32": 1e-3, "bf16": 2e-2}


@pytest.fixture(params=["tf32", "bf16"])
def mode(request, monkeypatch):
    monkeypatch.setenv("FMI_PRECISION", request.param)
    return request.param


def _ctx():
    from face_mask_inpaint_b200.modules import psp_fast as PF
    return PF, PF._Ctx(torch.device("cuda", 0))


def _nhwc(PF, k, x):
    """NCHW fp32 -> NHWC in the operand type (rounded like a producing kernel would)."""
    return PF._operand(x.permute(0, 2, 3, 1).contiguous(), k.mma)


def _nchw(y):
    return y.float().permute(0, 3, 1, 2).contiguous()


@pytest.fixture(autouse=True)
def _restore_tf32_switches():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _strict():
    """The PyTorch reference of these tests runs in strict fp32 (restored after every test: other test files assert the default)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("b,i,o,h,w", [(3, 64, 64, 20, 24), (2, 128, 256, 16, 16), (1, 64, 64, 6, 128), (2, 32, 512, 9, 7)])
def test_conv3x3_border_bias_prelu(mode, b, i, o, h, w):
    PF, k = _ctx()
    _strict()
    g = torch.Generator().manual_seed(b * 1000 + i + o)
    dev = k.dev
    x = torch.randn(b, i, h, w, generator=g).to(dev)
    wt = (torch.randn(o, i, 3, 3, generator=g) / (3 * i ** 0.5)).to(dev)
    shift = torch.randn(i, generator=g).to(dev)           # BatchNorm shift in front of the zero-padded conv
    slope = (0.25 + 0.1 * torch.randn(o, generator=g)).to(dev)
    xs = _nhwc(PF, k, x)
    want = F.conv2d(_nchw(xs) + shift.view(1, -1, 1, 1), PF._taps(wt, k.mma).float().reshape(3, 3, o, i).permute(2, 3, 0, 1), padding=1)
    want = torch.where(want > 0, want, want * slope.view(1, -1, 1, 1))
    y = k.empty(b, h, w, o)
    k.conv(xs, i, PF._taps(wt, k.mma), PF._border_bias(wt, shift), y, b, i, o, h, w, act=4, slope_c=slope, classes=9, round_y=0)
    assert rel_err(_nchw(y), want) <= TOL[mode]


@pytest.mark.parametrize("b,i,o,h,w", [(2, 64, 128, 16, 24), (3, 128, 128, 8, 8), (4, 256, 512, 4, 4), (5, 64, 64, 2, 2)])
def test_conv3x3_stride2_parity_planes(mode, b, i, o, h, w):
    PF, k = _ctx()
    _strict()
    g = torch.Generator().manual_seed(7 + b)
    dev = k.dev
    x = torch.randn(b, i, h, w, generator=g).to(dev)
    wt = (torch.randn(o, i, 3, 3, generator=g) / (3 * i ** 0.5)).to(dev)
    bias = torch.randn(o, generator=g).to(dev)
    xs = _nhwc(PF, k, x)
    want = F.leaky_relu(F.conv2d(_nchw(xs), PF._taps(wt, k.mma).float().reshape(3, 3, o, i).permute(2, 3, 0, 1), bias, stride=2, padding=1), 0.01)
    y = k.empty(b, h // 2, w // 2, o)
    k.conv(k.planes(xs, b, i, h, w), i, PF._taps(wt, k.mma), bias, y, b, i, o, h // 2, w // 2, planes=1, act=1, slope=0.01, round_y=0)
    assert rel_err(_nchw(y), want) <= TOL[mode]


def test_conv1x1_strided_shortcut(mode):
    PF, k = _ctx()
    _strict()
    g = torch.Generator().manual_seed(11)
    b, i, o, h, w = 2, 64, 128, 16, 24
    x = torch.randn(b, i, h, w, generator=g).to(k.dev)
    wt = (torch.randn(o, i, 1, 1, generator=g) / i ** 0.5).to(k.dev)
    bias = torch.randn(o, generator=g).to(k.dev)
    xs = _nhwc(PF, k, x)
    want = F.conv2d(_nchw(xs), PF._taps(wt, k.mma).float().reshape(1, 1, o, i).permute(2, 3, 0, 1), bias, stride=2)
    y = k.empty(b, h // 2, w // 2, o)
    k.conv(xs, i, PF._taps(wt, k.mma), bias, y, b, i, o, h // 2, w // 2, ksize=1, x_strides=(2 * i, 2 * w * i, h * w * i), round_y=0)
    assert rel_err(_nchw(y), want) <= TOL[mode]


@pytest.mark.parametrize("nh,n,hw", [(3, 4, 8), (2, 8, 4), (3, 2, 2), (11, 8, 2), (2, 3, 4)])
def test_heads_as_batch_entries(mode, nh, n, hw):
    """Per-head weight sets (w_group), per-head bias, several images per tile; input = the level-0 layout [n, hw, hw, nh*C]."""
    PF, k = _ctx()
    _strict()
    g = torch.Generator().manual_seed(nh * 100 + n * 10 + hw)
    c = 512
    x = torch.randn(n, nh * c, hw, hw, generator=g).to(k.dev)
    wts = [(torch.randn(c, c, 3, 3, generator=g) / (3 * c ** 0.5)).to(k.dev) for _ in range(nh)]
    biases = [torch.randn(c, generator=g).to(k.dev) for _ in range(nh)]
    xs = _nhwc(PF, k, x)
    want = torch.stack([F.leaky_relu(F.conv2d(_nchw(xs)[:, hd * c:(hd + 1) * c],
                                              PF._taps(wts[hd], k.mma).float().reshape(3, 3, c, c).permute(2, 3, 0, 1),
                                              biases[hd], stride=2, padding=1), 0.01) for hd in range(nh)])     # [nh, n, c, hw/2, hw/2]
    wp = torch.cat([PF._taps(wt, k.mma) for wt in wts], dim=0).contiguous()
    y = k.empty(nh * n, hw // 2, hw // 2, c)
    k.conv(k.planes(xs, n, c, hw, hw, heads=nh), c, wp, torch.cat(biases).contiguous(), y, nh * n, c, c, hw // 2, hw // 2, planes=1,
           w_group=n, bias_per_set=1, act=1, slope=0.01, round_y=0)
    assert rel_err(_nchw(y).reshape(nh, n, c, hw // 2, hw // 2), want) <= TOL[mode]


def test_se_gate_scale_add_and_upsample_add(mode):
    from face_mask_inpaint_b200 import _lib
    PF, k = _ctx()
    g = torch.Generator().manual_seed(5)
    b, c, h, w, red = 3, 128, 12, 10, 8
    dev = k.dev
    r = torch.randn(b, c, h, w, generator=g).to(dev)
    xfull = torch.randn(b, c, 2 * h, 2 * w, generator=g).to(dev)
    w1, w2 = (torch.randn(red, c, generator=g) / c ** 0.5).to(dev), torch.randn(c, red, generator=g).to(dev)
    rs, xs = _nhwc(PF, k, r), _nhwc(PF, k, xfull)
    mean = torch.empty(b, c, device=dev)
    gate = torch.empty(b, c, device=dev)
    scratch = torch.empty(b, 32, c, device=dev)
    _lib.check(k.lib.fmi_se_gate_nhwc(rs.data_ptr(), w1.data_ptr(), w2.data_ptr(), scratch.data_ptr(), mean.data_ptr(), gate.data_ptr(),
                                      b, c, red, h * w, k.mma, k.st), "fmi_se_gate_nhwc")
    m_want = _nchw(rs).mean(dim=(2, 3))
    g_want = torch.sigmoid(F.linear(torch.relu(F.linear(m_want, w1)), w2))
    assert rel_err(mean, m_want) <= 1e-5 and rel_err(gate, g_want) <= 1e-5
    y = k.empty(b, h, w, c)
    _lib.check(k.lib.fmi_se_scale_add_nhwc(rs.data_ptr(), gate.data_ptr(), xs.data_ptr(), 2 * c, 2 * 2 * w * c, 4 * h * w * c,
                                           y.data_ptr(), b, c, h, w, k.mma, k.st), "fmi_se_scale_add_nhwc")
    want = _nchw(rs) * g_want.view(b, c, 1, 1) + _nchw(xs)[:, :, ::2, ::2]
    assert rel_err(_nchw(y), want) <= (1e-3 if mode == "tf32" else 8e-3)
    # _upsample_add: bilinear, align_corners=True
    add = torch.randn(b, c, 2 * h, 2 * w, generator=g).to(dev)
    adds = _nhwc(PF, k, add)
    up = k.empty(b, 2 * h, 2 * w, c)
    _lib.check(k.lib.fmi_upsample_add_nhwc(rs.data_ptr(), adds.data_ptr(), up.data_ptr(), b, c, h, w, 2 * h, 2 * w, k.mma, k.st),
               "fmi_upsample_add_nhwc")
    want = F.interpolate(_nchw(rs), size=(2 * h, 2 * w), mode="bilinear", align_corners=True) + _nchw(adds)
    assert rel_err(_nchw(up), want) <= (1e-3 if mode == "tf32" else 8e-3)


def _randomize(enc, g):
    with torch.no_grad():
        for mod in enc.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
                mod.weight.copy_(1 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))


@pytest.mark.parametrize("use_ref", [True, False])
def test_whole_encoder_matches_cudnn_formulation(mode, monkeypatch, use_ref):
    """GradualStyleEncoder.forward on the kernels == the same module on cuDNN (strict fp32), codes [N, 18, 512]; and the kernel
    path launches no cuDNN convolution (launch count of this package grows by the trunk's GEMMs)."""
    from face_mask_inpaint_b200 import _lib
    from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    net = pSp(refpsp_opts(output_size=1024)).eval().to(dev)
    _randomize(net.encoder, torch.Generator().manual_seed(4))
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    ref = (torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).to(dev) if use_ref else None
    mask = torch.zeros(2, 256, 256, device=dev)
    mask[:, 128:230, 50:206] = 1
    lib = _lib.load()
    with torch.no_grad():
        n0 = lib.fmi_kernel_launch_count()
        got = net.encoder(x, ref=ref, mask=mask if use_ref else None)
        n1 = lib.fmi_kernel_launch_count()
        monkeypatch.setenv("FMI_PSP_CUDNN", "1")
        _strict()
        want = net.encoder(x, ref=ref, mask=mask if use_ref else None)
        n2 = lib.fmi_kernel_launch_count()
    assert got.shape == want.shape == (2, 18, 512)
    assert n1 - n0 > 100 and n1 - n0 > (n2 - n1) + 80, (n1 - n0, n2 - n1)
    assert rel_err(got, want) <= (2e-3 if mode == "tf32" else 2e-2), rel_err(got, want)


@pytest.mark.parametrize("tag", ["se_s1", "se_s2", "se_pool"])     # (ir_s1 has 16 channels: below the 32-channel tile granularity)
def test_ir_unit_kernels_match_the_reference_golden(mode, tag):
    """One bottleneck_IR(_SE) unit on the kernels (psp_fast._unit_forward) against the output the REFERENCE's own class produced
    for the same weights and input (tests/golden/psp_encoder.npz, recorded by make_golden.py from /root/reference)."""
    import numpy as np
    from face_mask_inpaint_b200.modules import psp as P
    PF, k = _ctx()
    g = {kk: torch.from_numpy(v) for kk, v in np.load(ROOT / "tests" / "golden" / "psp_encoder.npz").items()}
    sd = {kk[len(tag) + 4:]: v for kk, v in g.items() if kk.startswith(f"{tag}.sd.")}
    stride = int(g[f"{tag}.stride"])
    cin, depth = sd["res_layer.1.weight"].shape[1], sd["res_layer.1.weight"].shape[0]
    unit = P._IRUnit(cin, depth, stride, "res_layer.5.fc1.weight" in sd).eval()
    unit.load_state_dict(sd, strict=False)
    unit = unit.to(k.dev)
    u = PF._prep_unit(unit, k.mma)
    x = g[f"{tag}.x"].to(k.dev)
    b, _, h, w = x.shape
    y, oh, ow = PF._unit_forward(k, u, _nhwc(PF, k, x), b, h, w)
    want = g[f"{tag}.y"]
    assert (oh, ow) == tuple(want.shape[-2:])
    assert rel_err(_nchw(y), want) <= (2e-3 if mode == "tf32" else 2e-2)
