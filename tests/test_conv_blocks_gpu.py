"""f1: the PICNet decoder conv-block kernels (csrc/conv_blocks.cu) through the C ABI against the oracle
(oracle/ref_ops.py: F.conv2d / F.conv_transpose2d / F.instance_norm restatements of base_function.py:308-398) and against
the golden recorded from the reference's own ResBlockDecoder / Output (tests/golden/picnet_blocks.npz).
Tolerances: north_star's max|a-b|/max|b| <= 1e-3 for the fp32 contract (TF32 operands), <= 2e-2 for bf16 operands."""
import functools
import os
import types
from pathlib import Path

import numpy as np
import pytest
import torch
from torch import nn

from conftest import rel_err
from oracle import ref_ops as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "picnet_blocks.npz"
TOL = {0: 1e-3, 1: 2e-2}   # _lib.MMA_TF32, _lib.MMA_BF16


def _ctx(mma):
    from face_mask_inpaint_b200 import _lib
    from face_mask_inpaint_b200.modules import picnet_fast as PF
    os.environ["FMI_PRECISION"] = "bf16" if mma == _lib.MMA_BF16 else "fp32"
    try:
        return PF._Ctx(torch.device("cuda"))
    finally:
        os.environ.pop("FMI_PRECISION", None)


def _to_nhwc(k, x, stride=None, offset=0):
    """fp32 NCHW -> operand-type NHWC (optionally a channel slice [offset, offset + C) of a `stride`-channel buffer)."""
    from face_mask_inpaint_b200 import _lib
    b, c, h, w = x.shape
    stride = stride or c
    buf = torch.full((b, h, w, stride), float("nan"), dtype=k.dt, device="cuda")
    esz = buf.element_size()
    _lib.check(k.lib.fmi_nchw_to_nhwc_slice(x.contiguous().data_ptr(), buf.data_ptr() + offset * esz, b, c, h, w, stride,
                                            _lib.F32, 1, k.mma, k.st), "fmi_nchw_to_nhwc_slice")
    return buf


@pytest.mark.parametrize("mma", [0, 1])
@pytest.mark.parametrize("shape", [(2, 64, 32, 12, 20), (1, 96, 64, 33, 17), (3, 32, 256, 8, 8), (1, 512, 512, 4, 4),
                                   (1, 64, 32, 6, 260), (2, 32, 128, 3, 128)])   # the last two: halo mode of the plain conv (W >= 128)
@pytest.mark.parametrize("mode", [0, 2, 3])
def test_conv3x3_matches_oracle(mma, shape, mode):
    b, i, o, h, w = shape
    if mode == 3 and o > 64:
        pytest.skip("merged parity classes need O <= 64")
    g = torch.Generator().manual_seed(b * 1000 + i + o + h)
    x = torch.randn(b, i, h, w, generator=g)
    wt = torch.randn((i, o, 3, 3) if mode >= 2 else (o, i, 3, 3), generator=g) / (i * 9) ** 0.5
    bias = 0.1 * torch.randn(o, generator=g)
    if mode >= 2:
        want = torch.nn.functional.conv_transpose2d(x, wt, bias, stride=2, padding=1, output_padding=1)
    else:
        want = torch.nn.functional.conv2d(x, wt, bias, padding=1)
    k = _ctx(mma)
    xin = _to_nhwc(k, x.cuda(), stride=i + 32, offset=32)                # input is a channel slice of a wider buffer
    wp = k.weights([(wt.cuda().contiguous(), mode >= 2)], o, merged=mode == 3)
    oh, ow = want.shape[-2:]
    y = torch.full((b, oh, ow, o + 64), float("nan"), dtype=k.dt, device="cuda")   # output slice [64, 64 + o)
    nchw = torch.empty((b, o, oh, ow), dtype=torch.float32, device="cuda")
    esz = y.element_size()
    k.conv(xin.data_ptr() + 32 * esz, i + 32, wp, bias.cuda(), y.data_ptr() + 64 * esz, o + 64, 0, None if mode == 3 else nchw,
           0 if mode == 3 else o, b, i, o, h, w, mode, 2)
    got = y[..., 64:].float().permute(0, 3, 1, 2).cpu()
    assert rel_err(got, want) <= TOL[mma], rel_err(got, want)
    if mode != 3:
        assert rel_err(nchw.cpu(), want) <= TOL[mma]
    assert torch.isnan(y[..., :64].float()).all()                        # nothing written outside the slice


@pytest.mark.parametrize("shape", [(2, 64, 32, 12, 20), (1, 96, 64, 33, 17), (1, 512, 256, 4, 4), (1, 64, 32, 6, 260)])
@pytest.mark.parametrize("mode", [0, 2, 3, 4])
def test_conv3x3_split_operands_are_fp32_class(shape, mode, monkeypatch):
    """Strict-fp32 contract (FMI_PRECISION=tf32x3, or torch.backends.cudnn.allow_tf32 = False): the same GEMM kernel over the
    [hi | hi | lo] x [hi | lo | hi] operands of fmi_tf32_split3 — held to 2e-5 of a float64 convolution (single TF32: ~3e-4;
    measured 6e-6 at K = 576: what is left is the tensor core's truncating fp32 accumulation over K/8 instructions, not the split)."""
    from face_mask_inpaint_b200 import _lib
    from face_mask_inpaint_b200.modules import picnet_fast as PF
    b, i, o, h, w = shape
    if mode == 3 and o > 64:
        pytest.skip("merged parity classes need O <= 64")
    g = torch.Generator().manual_seed(b * 1000 + i + o + h + mode)
    x = torch.randn(b, i, h, w, generator=g)
    ks = 1 if mode == 4 else 3
    wt = torch.randn((i, o, 3, 3) if mode in (2, 3) else (o, i, ks, ks), generator=g) / (i * ks * ks) ** 0.5
    bias = 0.1 * torch.randn(o, generator=g)
    if mode in (2, 3):
        want = torch.nn.functional.conv_transpose2d(x.double(), wt.double(), bias.double(), stride=2, padding=1, output_padding=1)
    else:
        want = torch.nn.functional.conv2d(x.double(), wt.double(), bias.double(), padding=ks // 2)
    monkeypatch.setenv("FMI_PRECISION", "tf32x3")
    k = PF._Ctx(torch.device("cuda"))
    assert k.x3 and k.lib.fmi_get_tf32_exact() == 1
    try:
        xin = torch.full((b, h, w, i + 32), float("nan"), dtype=torch.float32, device="cuda")
        _lib.check(k.lib.fmi_nchw_to_nhwc_slice(x.cuda().data_ptr(), xin.data_ptr() + 32 * 4, b, i, h, w, i + 32, _lib.F32, 0,
                                                k.mma, k.st), "fmi_nchw_to_nhwc_slice")
        wp = k.weights([(wt.cuda().contiguous(), mode in (2, 3))], o, merged=mode == 3)
        assert wp.shape[-1] == 3 * i
        oh, ow = want.shape[-2:]
        y = torch.full((b, oh, ow, o + 64), float("nan"), dtype=torch.float32, device="cuda")
        k.conv(xin.data_ptr() + 32 * 4, i + 32, wp, bias.cuda(), y.data_ptr() + 64 * 4, o + 64, 0, None, 0, b, i, o, h, w, mode, 2)
    finally:
        k.finish()
    assert k.lib.fmi_get_tf32_exact() == 0
    got = y[..., 64:].permute(0, 3, 1, 2).cpu().double()
    e = float((got - want).abs().max() / want.abs().max())
    assert e <= (2e-5 if i * (1 if mode == 4 else 9) <= 1024 else 1e-4), e    # grows with K: accumulator truncation
    assert torch.isnan(y[..., :64]).all()


@pytest.mark.parametrize("mma", [0, 1])
def test_valid_conv_tanh_on_reflect_padded_input(mma):
    g = torch.Generator().manual_seed(5)
    b, c, h, w = 2, 32, 10, 24
    x = torch.randn(b, c, h, w, generator=g)
    wt = torch.randn(3, c, 3, 3, generator=g) / (c * 9) ** 0.5
    bias = 0.1 * torch.randn(3, generator=g)
    want = O.output_block(x, wt, bias, slope=0.1)
    k = _ctx(mma)
    # lrelu into the interior of the padded buffer, border by the reflect kernel, valid conv + tanh -> NCHW fp32
    from face_mask_inpaint_b200 import _lib
    xin = _to_nhwc(k, x.cuda())
    padded = torch.full((b, h + 2, w + 2, c), float("nan"), dtype=k.dt, device="cuda")
    interior = padded[:, 1:-1, 1:-1, :]
    act = torch.empty_like(xin)
    k.norm_act(xin.data_ptr(), c, act.data_ptr(), c, None, b, c, h * w, 0.1)
    interior.copy_(act)
    _lib.check(k.lib.fmi_reflect_border_nhwc(padded.data_ptr(), b, c, h, w, k.mma, k.st), "fmi_reflect_border_nhwc")
    ref_pad = torch.nn.functional.pad(torch.nn.functional.leaky_relu(x, 0.1), (1, 1, 1, 1), mode="reflect")
    assert rel_err(padded.float().permute(0, 3, 1, 2).cpu(), ref_pad) <= (1e-3 if mma == 0 else 8e-3)
    bo = torch.zeros(32, device="cuda")
    bo[:3] = bias.cuda()
    img = torch.empty((b, 3, h, w), dtype=torch.float32, device="cuda")
    k.conv(padded.data_ptr(), c, k.weights([(wt.cuda().contiguous(), False)], 32), bo, None, 32, 0, img, 3, b, c, 32, h, w, 1, 3)
    assert rel_err(img.cpu(), want) <= TOL[mma], rel_err(img.cpu(), want)


@pytest.mark.parametrize("mma", [0, 1])
@pytest.mark.parametrize("shape", [(2, 32, 3, 40, 72), (1, 64, 3, 16, 36), (2, 16, 1, 20, 20), (1, 32, 2, 5, 9)])
def test_output_conv_tanh_kernel_matches_oracle(mma, shape):
    """fmi_output_conv_tanh (SIMT, fp32 accumulation) on a reflection-padded input, full image and fused 4x4 pooling."""
    from face_mask_inpaint_b200 import _lib
    b, c, o, h, w = shape
    g = torch.Generator().manual_seed(c + h + o)
    x = torch.randn(b, c, h, w, generator=g)
    wt = torch.randn(o, c, 3, 3, generator=g) / (c * 9) ** 0.5
    bias = 0.1 * torch.randn(o, generator=g)
    k = _ctx(mma)
    xp = torch.nn.functional.pad(torch.nn.functional.leaky_relu(x, 0.1), (1, 1, 1, 1), mode="reflect")
    xin = _to_nhwc(k, xp.cuda())                                      # [B, H+2, W+2, C] operand type
    xr = xin.float().permute(0, 3, 1, 2).cpu()                        # what the kernel reads (bf16 / tf32 rounded)
    want = torch.tanh(torch.nn.functional.conv2d(xr, wt, bias))
    pool_ok = h % 4 == 0 and w % 4 == 0
    img = torch.empty((b, o, h, w), dtype=torch.float32, device="cuda")
    pooled = torch.empty((b, o, h // 4, w // 4), dtype=torch.float32, device="cuda") if pool_ok else None
    scratch = torch.empty(27 * c + 4, dtype=torch.float32, device="cuda")
    wt_d, bias_d = wt.cuda(), bias.cuda()       # keep the device copies alive: the launches are asynchronous
    _lib.check(k.lib.fmi_output_conv_tanh(xin.data_ptr(), wt_d.data_ptr(), bias_d.data_ptr(), img.data_ptr(),
                                          None if pooled is None else pooled.data_ptr(), scratch.data_ptr(), b, c, o, h, w,
                                          k.mma, k.st), "fmi_output_conv_tanh")
    assert rel_err(img.cpu(), want) <= 2e-5, rel_err(img.cpu(), want)   # fp32 arithmetic on identical inputs
    if pool_ok:
        assert rel_err(pooled.cpu(), torch.nn.functional.avg_pool2d(want, 4)) <= 2e-5
        only = torch.empty_like(pooled)                                # pooled output alone (no full-size image written)
        _lib.check(k.lib.fmi_output_conv_tanh(xin.data_ptr(), wt_d.data_ptr(), bias_d.data_ptr(), None, only.data_ptr(),
                                              scratch.data_ptr(), b, c, o, h, w, k.mma, k.st), "fmi_output_conv_tanh")
        assert torch.equal(only, pooled)


@pytest.mark.parametrize("mma", [0, 1])
@pytest.mark.parametrize("shape", [(2, 64, 12, 20), (1, 256, 64, 64), (3, 32, 7, 5)])
def test_instance_norm_act_matches_oracle(mma, shape):
    b, c, h, w = shape
    g = torch.Generator().manual_seed(c + h)
    x = torch.randn(b, c, h, w, generator=g) * 2 + 0.7
    gamma, beta = 1 + 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    k = _ctx(mma)
    xin = _to_nhwc(k, x.cuda(), stride=c + 32, offset=32)
    xr = xin[..., 32:].float().permute(0, 3, 1, 2).cpu()     # the operand-type values the kernel normalises
    want = torch.nn.functional.leaky_relu(torch.nn.functional.instance_norm(xr, weight=gamma, bias=beta, eps=1e-5), 0.1)
    norm = nn.InstanceNorm2d(c, affine=True).cuda()
    with torch.no_grad():
        norm.weight.copy_(gamma)
        norm.bias.copy_(beta)
    y = torch.empty((b, h, w, c), dtype=k.dt, device="cuda")
    esz = y.element_size()
    k.norm_act(xin.data_ptr() + 32 * esz, c + 32, y.data_ptr(), c, norm, b, c, h * w, 0.1)
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert rel_err(got, want) <= (5e-4 if mma == 0 else 8e-3), rel_err(got, want)


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("shape", [(32, 64), (256, 256), (3, 32)])
def test_spectral_norm_weight_prep_matches_oracle(transposed, shape):
    """fmi_conv_weight_prep_sn = SpectralNorm._update_u_v (one power iteration, u / v updated in place) + w_bar / sigma in the
    operand layout, against oracle.spectral_norm_weight (external_function.py:44-57)."""
    from face_mask_inpaint_b200 import _lib
    o, i = shape
    g = torch.Generator().manual_seed(o + i)
    wshape = (i, o, 3, 3) if transposed else (o, i, 3, 3)
    w_bar = torch.randn(wshape, generator=g) / (i * 9) ** 0.5
    hh = wshape[0]
    u = torch.randn(hh, generator=g)
    u /= u.norm()
    v = torch.randn(w_bar.numel() // hh, generator=g)
    v /= v.norm()
    w_eff, u_new, v_new = O.spectral_norm_weight(w_bar, u, v)
    k = _ctx(0)
    o_rows = max(32, o)
    wp = torch.zeros((9, o_rows, i), dtype=torch.float32, device="cuda")
    wb_d, u_d, v_d = w_bar.cuda(), u.cuda(), v.cuda()
    scratch = torch.empty(u.numel() + v.numel(), dtype=torch.float32, device="cuda")
    _lib.check(k.lib.fmi_conv_weight_prep_sn(wb_d.data_ptr(), u_d.data_ptr(), v_d.data_ptr(), scratch.data_ptr(), wp.data_ptr(),
                                             o, i, int(transposed), o_rows, i, 0, 0, 3, k.mma, k.st), "fmi_conv_weight_prep_sn")
    want = (w_eff.permute(2, 3, 1, 0) if transposed else w_eff.permute(2, 3, 0, 1)).reshape(9, o, i)
    assert rel_err(wp[:, :o].cpu(), want) <= 5e-4            # tf32 rounding of the stored weights (2^-11)
    assert rel_err(u_d.cpu(), u_new) <= 1e-5 and rel_err(v_d.cpu(), v_new) <= 1e-5
    assert float(wp[:, o:].abs().max()) == 0.0 if o_rows > o else True


@pytest.mark.parametrize("mma", [0, 1])
def test_conv1x1_and_residual_sum_epilogue(mma):
    """mode 4 (1x1 conv) writes the shortcut, a 3x3 conv with act + 10 adds the main path onto it (ResBlock sum,
    base_function.py:262-268), then AvgPool2d(2) of the sum."""
    from face_mask_inpaint_b200 import _lib
    g = torch.Generator().manual_seed(9)
    b, cin, ch, co, h, w = 2, 64, 32, 96, 12, 20
    x, a2 = torch.randn(b, cin, h, w, generator=g), torch.randn(b, ch, h, w, generator=g)
    wb = torch.randn(co, cin, 1, 1, generator=g) / cin ** 0.5
    w2 = torch.randn(co, ch, 3, 3, generator=g) / (ch * 9) ** 0.5
    bb, b2 = 0.1 * torch.randn(co, generator=g), 0.1 * torch.randn(co, generator=g)
    F = torch.nn.functional
    want = F.conv2d(x, wb, bb) + F.conv2d(a2, w2, b2, padding=1)
    k = _ctx(mma)
    xin, ain = _to_nhwc(k, x.cuda()), _to_nhwc(k, a2.cuda())
    y = torch.full((b, h, w, co), float("nan"), dtype=k.dt, device="cuda")
    k.conv(xin.data_ptr(), cin, k.weights([(wb.cuda().contiguous(), False)], co), bb.cuda(), y.data_ptr(), co, 0, None, 0, b, cin,
           co, h, w, 4, 2, round_y=0)
    short = y.float().permute(0, 3, 1, 2).cpu()
    assert rel_err(short, F.conv2d(x, wb, bb)) <= TOL[mma]
    k.conv(ain.data_ptr(), ch, k.weights([(w2.cuda().contiguous(), False)], co), b2.cuda(), y.data_ptr(), co, 0, None, 0, b, ch, co,
           h, w, 0, 12, round_y=0)
    assert rel_err(y.float().permute(0, 3, 1, 2).cpu(), want) <= TOL[mma]
    yp = torch.empty((b, h // 2, w // 2, co), dtype=k.dt, device="cuda")
    _lib.check(k.lib.fmi_avgpool2_nhwc(y.data_ptr(), co, yp.data_ptr(), co, b, co, h, w, 0, k.mma, k.st), "fmi_avgpool2_nhwc")
    assert rel_err(yp.float().permute(0, 3, 1, 2).cpu(), F.avg_pool2d(y.float().permute(0, 3, 1, 2).cpu(), 2)) <= (1e-6 if mma == 0 else 4e-3)


@pytest.mark.parametrize("kind", ["src_encoder", "ref_encoder"])
def test_res_encoder_kernel_path_matches_cudnn_fp32(kind):
    """ResEncoder.forward (network.py:133-172; 5 trunk blocks + 7 / 1 distribution-head blocks) on the kernels vs the same
    module on cuDNN in strict fp32, same weights, same fresh SpectralNorm state."""
    import copy
    from face_mask_inpaint_b200 import _lib
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    from golden_util import fill_by_name
    base = getattr(fill_by_name(build_picnet_ref()).eval(), kind)
    img = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(3)).cuda()
    old = torch.backends.cudnn.allow_tf32

    def run(force_cudnn, tf32):
        m = copy.deepcopy(base).cuda()
        os.environ["FMI_PICNET_CUDNN"] = "1" if force_cudnn else "0"
        torch.backends.cudnn.allow_tf32 = tf32
        n0 = _lib.load().fmi_kernel_launch_count()
        with torch.no_grad():
            (mu, std), feats = m(img)
        return mu, std, feats, _lib.load().fmi_kernel_launch_count() - n0, m

    try:
        t_mu, t_std, t_f, n_c, m_c = run(True, False)
        r_mu, r_std, r_f, _, _ = run(True, True)
        o_mu, o_std, o_f, n_o, m_o = run(False, True)
        x_mu, x_std, x_f, n_x, _ = run(False, False)      # TF32 off: the kernels with split (3xTF32) operands
    finally:
        os.environ.pop("FMI_PICNET_CUDNN", None)
        torch.backends.cudnn.allow_tf32 = old
    assert n_x > n_o
    for got, want, name in ((x_f, t_f, "features"), (x_mu, t_mu, "mu"), (x_std, t_std, "std")):
        assert rel_err(got, want) <= 1e-4, (name, rel_err(got, want))     # measured 2e-5; single TF32: ~1e-3
    # the forced-cuDNN path still runs its SpectralNorm power iterations on this package's 3 kernels per wrapped convolution
    from face_mask_inpaint_b200.modules.picnet_blocks import SpectralNorm
    n_sn = sum(isinstance(mm, SpectralNorm) for mm in base.modules())
    assert n_c == 3 * n_sn and n_o > 40
    assert o_f.shape == t_f.shape == (2, 128, 8, 12) and o_mu.shape == t_mu.shape == (2, 128, 8, 12)
    for got, ref_gpu, want, name in ((o_f, r_f, t_f, "features"), (o_mu, r_mu, t_mu, "mu"), (o_std, r_std, t_std, "std")):
        e, e_ref = rel_err(got, want), rel_err(ref_gpu, want)
        assert e <= 1.5 * e_ref + 1e-3, (name, e, e_ref)
    # the power iteration advanced u / v exactly as the module's own forward does
    u_c, u_o = m_c.block0.conv1.module.weight_u, m_o.block0.conv1.module.weight_u
    assert rel_err(u_o, u_c) <= 1e-5


def test_batched_weight_plan_matches_per_conv_path(monkeypatch):
    """fmi_conv_weight_prep_sn_batch (3 launches for all convolutions, from the second forward on) gives the same images and
    the same SpectralNorm state as one fmi_conv_weight_prep_sn per convolution."""
    import copy
    from face_mask_inpaint_b200 import _lib
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    from golden_util import fill_by_name, mean_z, picnet_inputs
    base = fill_by_name(build_picnet_ref()).eval()
    src, ref, mask = (t.cuda() for t in picnet_inputs(1))
    outs, states, launches = [], [], []
    for batch in ("1", "0"):
        monkeypatch.setenv("FMI_SN_BATCH", batch)
        m = copy.deepcopy(base).cuda()
        m.decoder.get_z = types.MethodType(mean_z, m.decoder)
        with torch.no_grad():
            m(src, ref, mask)                       # records the plan (or not)
            n0 = _lib.load().fmi_kernel_launch_count()
            outs.append(m(src, ref, mask))
            launches.append(_lib.load().fmi_kernel_launch_count() - n0)
            outs.append(m(src, ref, mask))
        states.append((m.decoder.decoder3.conv2.module.weight_u.clone(), m.src_encoder.prior.bypass.module.weight_v.clone()))
    assert launches[0] < launches[1] - 150, launches           # 69 convolutions x 3 launches -> 3 x 3 launches
    for a, b in zip(states[0], states[1]):
        assert rel_err(a, b) <= 1e-5                            # measured 3e-7: another summation order of the same products
    # This random-weight generator amplifies a 3e-7 change of the weights to ~3e-3 of the image (tools/debug/dbg_plan.py: two per-conv
    # runs are bit-identical, batched vs per-conv differ by 3.6e-3), the same sensitivity that turns TF32 operand rounding into
    # 1e-2 (test_whole_generator_kernel_path_vs_cudnn_paths); the images are therefore only held to that scale here.
    assert rel_err(outs[0], outs[2]) <= 2e-2 and rel_err(outs[1], outs[3]) <= 2e-2


def test_weight_plan_is_keyed_by_call_pattern():
    """The same ResGenerator called with and without z (the z -> f block's convolutions are only used with z) keeps one
    weight plan per call pattern."""
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    from golden_util import fill_by_name
    dec = fill_by_name(build_picnet_ref()).eval().cuda().decoder
    x = torch.randn(1, 256, 8, 8, device="cuda")
    z = torch.randn(1, 256, 8, 8, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            a = dec(x, z=z)
            b = dec(x)
            c = dec(x, z=z, pool_to=(64, 64))
    assert a.shape == b.shape == (1, 3, 256, 256) and c.shape == (1, 3, 64, 64)
    assert torch.isfinite(a).all() and torch.isfinite(b).all() and rel_err(a, b) > 1e-3
    import copy
    clone = copy.deepcopy(dec)                      # a module that has run (plans, streams cached) stays deep-copyable
    with torch.no_grad():
        d = clone(x, z=z)
        e = dec(x, z=z)
    assert rel_err(d, e) <= 2e-2                    # same state, own weight plan (see the sensitivity note below)


def _mirror_from_golden():
    from face_mask_inpaint_b200.modules import picnet as P
    g = {k: torch.from_numpy(v) for k, v in np.load(GOLD).items()}
    norm = functools.partial(nn.InstanceNorm2d, affine=True)
    blk = P.ResBlockDecoder(64, 32, 32, norm, nn.LeakyReLU(0.1), True, False).eval()
    out = P.Output(32, 3, 3, None, nn.LeakyReLU(0.1), True, False).eval()
    for tag, m in (("blk", blk), ("out", out)):
        sd = {k[len(tag) + 1:]: v for k, v in g.items() if k.startswith(tag + ".")}
        missing, unexpected = m.load_state_dict(sd, strict=False)   # the golden keeps each shared conv under its first name only
        assert not unexpected and all(".module." in k for k in missing)
    return g, blk.cuda(), out.cuda()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_decoder_block_and_output_match_reference_golden(prec):
    """One ResBlockDecoder + Output through picnet_fast.decoder_forward (SpectralNorm power iteration included) against the
    reference's own classes."""
    from face_mask_inpaint_b200.modules import picnet_fast as PF
    g, blk, out = _mirror_from_golden()
    gen = types.SimpleNamespace(layers=1, use_attn=False, decoder0=blk, out0=out)
    os.environ["FMI_PRECISION"] = prec
    try:
        with torch.no_grad():
            assert PF.supported(gen, g["x"].cuda())
            img = PF.decoder_forward(gen, g["x"].cuda())
    finally:
        os.environ.pop("FMI_PRECISION", None)
    e = rel_err(img.cpu(), g["img"])
    assert e <= (1e-3 if prec == "fp32" else 2e-2), e


def test_whole_generator_kernel_path_vs_cudnn_paths():
    """ReferenceFill forward (README configuration, 1024^2 decoder output) three ways on the same weights and the same fresh
    SpectralNorm state: cuDNN strict fp32 (the truth), cuDNN with TF32 operands (what the reference executes on a GPU under
    PyTorch's defaults), and the decoder blocks on this package's kernels (TF32 operands). The kernel path must be as close
    to the truth as the reference's own GPU path (measured: 1.0e-2 vs 1.4e-2 on the image, 3.4e-3 vs 4.6e-3 before the
    Output block), must launch this package's kernels, and with TF32 switched off must run the strict-fp32 contract on the kernels too (split operands, <= 1e-3 of the truth)."""
    import copy
    from face_mask_inpaint_b200 import _lib
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    from golden_util import fill_by_name, mean_z, picnet_inputs
    base = fill_by_name(build_picnet_ref()).eval()
    src, ref, mask = (t.cuda() for t in picnet_inputs(2))
    old = torch.backends.cudnn.allow_tf32

    def run(force_cudnn, tf32):
        m = copy.deepcopy(base).cuda()
        m.decoder.get_z = types.MethodType(mean_z, m.decoder)
        os.environ["FMI_PICNET_CUDNN"] = "1" if force_cudnn else "0"
        torch.backends.cudnn.allow_tf32 = tf32
        n0 = _lib.load().fmi_kernel_launch_count()
        with torch.no_grad():
            out = m(src, ref, mask, resize=False)
        return out, _lib.load().fmi_kernel_launch_count() - n0

    try:
        truth, n_cudnn = run(True, False)
        ref_gpu, _ = run(True, True)
        ours, n_ours = run(False, True)
        strict, n_strict = run(False, False)
        os.environ["FMI_PRECISION"] = "fp32"             # single-pass TF32 pinned + TF32 switched off: cuDNN strict fp32
        pinned, n_pinned = run(False, False)
    finally:
        os.environ.pop("FMI_PICNET_CUDNN", None)
        os.environ.pop("FMI_PRECISION", None)
        torch.backends.cudnn.allow_tf32 = old
    assert truth.shape == ours.shape == (2, 3, 1024, 1024)
    assert n_ours > n_cudnn + 40, (n_ours, n_cudnn)      # 5 blocks x (2 stats + 2 norm_act + 5 GEMMs + 3 weight preps) + Output
    # TF32 off: strict fp32 convolutions on the same kernels with split (3xTF32) operands — one split pass per GEMM on top
    assert n_strict > n_ours, (n_strict, n_ours)
    assert n_pinned == n_cudnn
    e_strict = rel_err(strict, truth)
    print(f"strict-fp32 contract on the kernels: {e_strict:.3e}; cuDNN strict fp32 twice: {rel_err(pinned, truth):.3e}")
    # two strict-fp32 cuDNN runs differ by 6e-4 .. 8e-4 themselves (cuDNN algorithm choice; these N(0, 1/fan_in) weights amplify
    # rounding ~1000x), so `truth` is only known to that: measured 1.0e-3 for the split-operand kernels. The <= 1e-3 check against
    # the reference's own CPU fp32 arithmetic is tests/test_picnet_gpu.py (golden; measured 5e-4).
    assert rel_err(pinned, truth) <= 2e-3
    assert e_strict <= 2e-3 and e_strict <= 2 * rel_err(pinned, truth) + 5e-4, (e_strict, rel_err(pinned, truth))
    e_ours, e_ref = rel_err(ours, truth), rel_err(ref_gpu, truth)
    assert e_ours <= 1.5 * e_ref + 1e-3, (e_ours, e_ref)     # measured 1.56e-2 vs 1.37e-2 (stable over the round's runs)
    assert e_ours <= 2e-2, e_ours
    # resize=True: the 4x4 average pooling is fused into the Output kernel
    m = copy.deepcopy(base).cuda()
    m.decoder.get_z = types.MethodType(mean_z, m.decoder)
    with torch.no_grad():
        small = m(src, ref, mask)
    assert small.shape == (2, 3, 256, 256)
    assert rel_err(small, torch.nn.functional.avg_pool2d(ours, 4)) <= 1e-5


@pytest.mark.parametrize("shape", [(64, 32, 3, 3), (32, 96, 3, 3), (256, 128, 1, 1), (3, 32, 3, 3)])
def test_fused_spectral_norm_forward_backward(shape):
    """ops.spectral_norm_weight (fmi_spectral_norm_fwd / _bwd) == SpectralNorm._update_u_v of the reference
    (external_function.py:30-42) restated in fp64 with autograd: the weight, the in-place u / v update and the gradient."""
    from face_mask_inpaint_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    w = torch.randn(*shape, generator=g)
    hh = shape[0]
    u = torch.nn.functional.normalize(torch.randn(hh, generator=g), dim=0)
    v = torch.nn.functional.normalize(torch.randn(w.numel() // hh, generator=g), dim=0)
    go = torch.randn(*shape, generator=g)
    # reference formulation, fp64
    wr = w.double().requires_grad_(True)
    l2 = lambda t: t / (t.norm() + 1e-12)
    vn = l2(torch.mv(wr.view(hh, -1).data.t(), u.double()))
    un = l2(torch.mv(wr.view(hh, -1).data, vn))
    sigma = un.dot(wr.view(hh, -1).mv(vn))
    want = wr / sigma
    want.backward(go.double())
    # kernels
    wd = w.cuda().requires_grad_(True)
    ud, vd = u.cuda(), v.cuda()
    got = ops.spectral_norm_weight(wd, ud, vd)
    got.backward(go.cuda())
    assert rel_err(got, want.detach().float()) <= 1e-5
    assert rel_err(ud, un.float()) <= 1e-5 and rel_err(vd, vn.float()) <= 1e-5
    assert rel_err(wd.grad, wr.grad.float()) <= 1e-5
