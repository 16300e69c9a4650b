"""Inference path of the pSp encoder — IR-SE50 trunk, FPN adds, the 18 map2style heads — on the sm_100a kernels
(csrc/ir_encoder.cu).

`GradualStyleEncoder.forward` (modules/psp/encoders/psp_encoders.py:100-152) runs the IR-SE50 body on the source and on the
reference image (:101-125; bottleneck_IR_SE units, encoders/helpers.py:97-119), resizes the mask, calls the two
ExampleGuidedAttention modules / the masked blend (:127-138), then the map2style heads over the three pyramid levels
(:140-151). Here source and reference go through the trunk as ONE 2N batch, activations stay NHWC in the tensor-core operand
type (bf16 with FMI_PRECISION=bf16, else tf32-rounded fp32), every convolution is the tcgen05 implicit GEMM
(fmi_conv_nhwc) and eval-mode BatchNorm is folded into weights / biases once per set of parameters:
  conv1 of a unit  W1[o,i,t] * s1[i]  and the BORDER-CLASS bias  b1[cls][o] = sum_{t inside for cls} sum_i W1[o,i,t] * t1[i]
                   (BN1 precedes a zero-padded conv: its shift reaches a pixel only through the taps inside the image);
  conv2 / shortcut W[o,i,t] * s[o],  bias t[o];   the input layer likewise;
with s = gamma / sqrt(var + eps), t = beta - mean * s. The folded weights are [tap][O][I] in the operand type.
Used when autograd is off, the module is in eval mode (BatchNorm uses running statistics) and the tensors are CUDA fp32;
training keeps the cuDNN formulation (BatchNorm batch statistics, autograd).
Works on this package's mirror (modules/psp.py) and on the reference's own class (patch.py): same attribute layout.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn

from .. import _lib, ops
from ..graphs import module_cache

_TAPS = (6, 20, 23)   # body indices whose outputs are c1, c2, c3 (psp_encoders.py:103-108)


def _p(t):
    return None if t is None else t.data_ptr()


def _bn_fold(bn):
    s = (bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps))
    return s, bn.bias.detach().float() - bn.running_mean.detach().float() * s


def _operand(w, mma):
    """fp32 tensor -> the tensor-core operand type: bf16 (round to nearest even) or tf32 (10-bit mantissa, round to nearest,
    ties away — what cvt.rna.tf32.f32 does) in an fp32 container."""
    w = w.contiguous()
    if mma == _lib.MMA_BF16:
        return w.to(torch.bfloat16)
    return ((w.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _taps(w, mma, i_row=None):
    """Conv2d weight [O, I, k, k] -> [k*k][O][I_row] in the operand type (input channels zero-padded to I_row)."""
    o, i, kh, kw = w.shape
    t = w.permute(2, 3, 0, 1).reshape(kh * kw, o, i)
    if i_row is not None and i_row > i:
        t = torch.nn.functional.pad(t, (0, i_row - i))
    return _operand(t, mma)


def _border_bias(w, shift):
    """[9][O]: class 3*vy + vx (0 first row / column, 2 last, 1 inside) -> sum over the taps inside the image of
    sum_i w[o,i,ky,kx] * shift[i]   (3x3 conv, padding 1, plane of at least 2x2)."""
    per_tap = torch.einsum("oikl,i->okl", w, shift)            # [O, 3, 3]
    out = []
    for vy in range(3):
        for vx in range(3):
            m = torch.ones(3, 3, device=w.device)
            if vy == 0:
                m[0, :] = 0          # the row above the image is padding
            if vy == 2:
                m[2, :] = 0
            if vx == 0:
                m[:, 0] = 0
            if vx == 2:
                m[:, 2] = 0
            out.append((per_tap * m).sum(dim=(1, 2)))
    return torch.stack(out).contiguous()


class _Unit:
    __slots__ = ("in_c", "depth", "stride", "w1", "b1", "slope", "w2", "b2", "ws", "bs", "se1", "se2", "red")


def _prep_unit(unit, mma):
    res = list(unit.res_layer)
    bn1, c1, pr, c2, bn2 = res[:5]
    se = res[5] if len(res) > 5 else None
    u = _Unit()
    u.in_c, u.depth, u.stride = c1.in_channels, c1.out_channels, c2.stride[0]
    s1, t1 = _bn_fold(bn1)
    w1 = c1.weight.detach().float()
    u.w1 = _taps(w1 * s1.view(1, -1, 1, 1), mma)
    u.b1 = _border_bias(w1, t1)
    u.slope = pr.weight.detach().float().contiguous()
    s2, t2 = _bn_fold(bn2)
    u.w2 = _taps(c2.weight.detach().float() * s2.view(-1, 1, 1, 1), mma)
    u.b2 = t2.contiguous()
    if isinstance(unit.shortcut_layer, nn.MaxPool2d):
        u.ws = u.bs = None
    else:
        cs, bns = unit.shortcut_layer[0], unit.shortcut_layer[1]
        ss, ts = _bn_fold(bns)
        u.ws = _taps(cs.weight.detach().float() * ss.view(-1, 1, 1, 1), mma)
        u.bs = ts.contiguous()
    if se is not None:
        u.se1 = se.fc1.weight.detach().float().reshape(se.fc1.out_channels, -1).contiguous()
        u.se2 = se.fc2.weight.detach().float().reshape(se.fc2.out_channels, -1).contiguous()
        u.red = se.fc1.out_channels
    else:
        u.se1 = u.se2 = None
        u.red = 0
    return u


class _Plan:
    """Folded weights of one encoder for one operand type."""

    def __init__(self, enc, mma):
        conv, bn, pr = enc.input_layer[0], enc.input_layer[1], enc.input_layer[2]
        s, t = _bn_fold(bn)
        self.w_in = _taps(conv.weight.detach().float() * s.view(-1, 1, 1, 1), mma, i_row=32)
        self.b_in = t.contiguous()
        self.slope_in = pr.weight.detach().float().contiguous()
        self.units = [_prep_unit(u, mma) for u in enc.body]
        self.lat1 = (_taps(enc.latlayer1.weight.detach().float(), mma), enc.latlayer1.bias.detach().float().contiguous())
        self.lat2 = (_taps(enc.latlayer2.weight.detach().float(), mma), enc.latlayer2.bias.detach().float().contiguous())
        # out_conv of the two attention modules (example_guided_att.py:13,38-39): a 1x1 GEMM on the NHWC copy of the concat
        self.att_out = {}
        for name in ("attention1", "attention2"):
            att = getattr(enc, name, None)
            if att is not None and getattr(att, "out_channels", None) is not None:
                self.att_out[name] = (_taps(att.out_conv.weight.detach().float(), mma), att.out_conv.bias.detach().float().contiguous())
        # map2style heads grouped by the pyramid level they read
        groups = [(0, enc.coarse_ind), (enc.coarse_ind, enc.middle_ind), (enc.middle_ind, enc.style_count)]
        self.heads = []
        for lo, hi in groups:
            blocks = [enc.styles[j] for j in range(lo, hi)]
            if not blocks:
                self.heads.append(None)
                continue
            depth = len([m for m in blocks[0].convs if isinstance(m, nn.Conv2d)])
            levels = []
            for d in range(depth):
                convs = [[m for m in b.convs if isinstance(m, nn.Conv2d)][d] for b in blocks]
                # level 0: ONE weight set [9][heads*O][I]; deeper: one set per head [heads][9][O][I] (and one bias per head)
                w = torch.cat([_taps(c.weight.detach().float(), mma) for c in convs], dim=1 if d == 0 else 0)
                bias = torch.cat([c.bias.detach().float() for c in convs]).contiguous()
                levels.append((w.contiguous(), bias))
            slope = float([m for m in blocks[0].convs if isinstance(m, nn.LeakyReLU)][0].negative_slope)
            lin_w = torch.stack([b.linear.weight.detach().float() * b.linear.scale for b in blocks])     # [heads, out, in] fp32
            lin_b = torch.cat([(b.linear.bias.detach().float() * b.linear.lr_mul) for b in blocks]).contiguous()
            self.heads.append((len(blocks), levels, slope, lin_w.contiguous(), lin_b, blocks[0].out_c))


def _versions(enc):
    return sum(p._version for p in enc.parameters()) + sum(b._version for b in enc.buffers())


def supported(enc, x, ref) -> bool:
    if os.environ.get("FMI_PSP_CUDNN") == "1" or torch.is_grad_enabled() or enc.training:
        return False
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
        return False
    if x.shape[2] % 32 or x.shape[3] % 32 or x.shape[2] < 64 or x.shape[3] < 64:
        return False
    ok = getattr(enc, "_fmi_fast_ok", None)
    if ok is None:
        try:
            units = list(enc.body)
            ok = len(units) == 24 and all(len(list(u.res_layer)) in (5, 6) for u in units)
            ok = ok and all(isinstance(u.res_layer[2], nn.PReLU) and u.res_layer[1].kernel_size == (3, 3) for u in units)
            ok = ok and all(b.out_c == 512 for b in enc.styles) and enc.style_count == len(enc.styles)
            # a head reduces its level to 1x1 with stride-2 convs: spatial must be the level's size at a 256x256 input
            ok = ok and isinstance(enc.input_layer[2], nn.PReLU)
        except Exception:
            ok = False
        enc._fmi_fast_ok = ok
    if not ok:
        return False
    # map2style heads reduce their pyramid level to 1x1: level sizes (H/16, H/8, H/4) must equal the heads' `spatial`
    h, w = x.shape[2], x.shape[3]
    if h != w:
        return False
    for j, blk in enumerate(enc.styles):
        level = h // 16 if j < enc.coarse_ind else h // 8 if j < enc.middle_ind else h // 4
        if blk.spatial != level:
            return False
    return True


class _Ctx:
    def __init__(self, dev):
        self.lib = _lib.load()
        self.mma = ops.mma_mode(torch.float32)
        self.dt = torch.float32 if self.mma == _lib.MMA_TF32 else torch.bfloat16
        self.dev = dev
        self.st = ops._stream()

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.dt, device=self.dev)

    def conv(self, x, xs, wp, bias, y, b, i, o, h, w, ksize=3, planes=0, w_group=0, bias_per_set=0, act=2, slope=0.0, slope_c=None,
             classes=1, x_strides=None, round_y=1):
        """x: tensor or pointer; xs: channels per pixel of the input buffer; x_strides: (pixel, row, image) element strides of a
        strided view, default dense [.., h, w, xs]."""
        ps, rs, ims = x_strides if x_strides is not None else (xs, w * xs, h * w * xs)
        _lib.check(self.lib.fmi_conv_nhwc(x if isinstance(x, int) else x.data_ptr(), ps, rs, ims, _p(wp), _p(bias), classes,
                                          _p(slope_c), float(slope), y.data_ptr(), y.shape[-1], b, i, o, h, w, ksize, planes,
                                          w_group, bias_per_set, act, 0, round_y, self.mma, self.st), "fmi_conv_nhwc")

    def planes(self, x, b, c, h, w, heads=1):
        y = self.empty(4 * heads * b, h // 2, w // 2, c)
        _lib.check(self.lib.fmi_space_to_planes_nhwc(x.data_ptr(), x.shape[-1], y.data_ptr(), b, c, h, w, heads, self.mma, self.st),
                   "fmi_space_to_planes_nhwc")
        return y


def _unit_forward(k: _Ctx, u: _Unit, x, b, h, w):
    """One bottleneck_IR(_SE) unit on x [b, h, w, in_c]; returns (y [b, h/s, w/s, depth], h/s, w/s)."""
    a1 = k.empty(b, h, w, u.depth)
    k.conv(x, u.in_c, u.w1, u.b1, a1, b, u.in_c, u.depth, h, w, act=4, slope_c=u.slope, classes=9)
    oh, ow = h // u.stride, w // u.stride
    r = k.empty(b, oh, ow, u.depth)
    if u.stride == 2:
        k.conv(k.planes(a1, b, u.depth, h, w), u.depth, u.w2, u.b2, r, b, u.depth, u.depth, oh, ow, planes=1)
    else:
        k.conv(a1, u.depth, u.w2, u.b2, r, b, u.depth, u.depth, oh, ow)
    del a1
    sub = (u.stride * u.in_c, u.stride * w * u.in_c, h * w * u.in_c)      # x[:, ::s, ::s, :]
    if u.ws is None:
        sc, sc_str = x, sub
    else:
        sc = k.empty(b, oh, ow, u.depth)
        k.conv(x, u.in_c, u.ws, u.bs, sc, b, u.in_c, u.depth, oh, ow, ksize=1, x_strides=sub)
        sc_str = (u.depth, ow * u.depth, oh * ow * u.depth)
    if u.se1 is None:
        gate = torch.ones((b, u.depth), dtype=torch.float32, device=k.dev)
    else:
        scratch = torch.empty((b, 32, u.depth), dtype=torch.float32, device=k.dev)     # [b][32 slabs][C] partial sums
        mean, gate = torch.empty((2, b, u.depth), dtype=torch.float32, device=k.dev)
        _lib.check(k.lib.fmi_se_gate_nhwc(r.data_ptr(), _p(u.se1), _p(u.se2), scratch.data_ptr(), _p(mean), _p(gate), b, u.depth,
                                          u.red, oh * ow, k.mma, k.st), "fmi_se_gate_nhwc")
    y = k.empty(b, oh, ow, u.depth)
    _lib.check(k.lib.fmi_se_scale_add_nhwc(r.data_ptr(), _p(gate), sc.data_ptr(), sc_str[0], sc_str[1], sc_str[2], y.data_ptr(), b,
                                           u.depth, oh, ow, k.mma, k.st), "fmi_se_scale_add_nhwc")
    return y, oh, ow


def _to_nchw(k, x, b, c, h, w):
    y = torch.empty((b, c, h, w), dtype=torch.float32, device=k.dev)
    _lib.check(k.lib.fmi_nhwc_to_nchw(x.data_ptr(), _p(y), b, c, h, w, k.mma, _lib.F32, k.st), "fmi_nhwc_to_nchw")
    return y


def _to_nhwc(k, x):
    b, c, h, w = x.shape
    x = x.contiguous()
    y = k.empty(b, h, w, c)
    _lib.check(k.lib.fmi_nchw_to_nhwc_slice(_p(x), y.data_ptr(), b, c, h, w, c, _lib.F32, 1, k.mma, k.st), "fmi_nchw_to_nhwc_slice")
    return y


def _attention_nhwc(k: _Ctx, plan, name, att, mask_s, src, ref):
    """ExampleGuidedAttention.forward (example_guided_att.py:21-41) -> NHWC operand-type features: the fused attention kernel on
    the module's NCHW interface, then — when the module has an out_conv — that 1x1 convolution as an implicit GEMM on the NHWC
    copy of the concat (the module's own out_conv is a SIMT fp32 kernel: 0.26 ms of a batch-8 forward for 2 x 2 GFLOP)."""
    from .attention import _EGAFunction
    ent = plan.att_out.get(name)
    if ent is None:
        return _to_nhwc(k, att(mask_s, src, ref))
    cat = _to_nhwc(k, _EGAFunction.apply(mask_s, src, ref, att.conv.weight))
    n, h, w, c2 = cat.shape
    wp, bias = ent
    y = k.empty(n, h, w, wp.shape[1])
    k.conv(cat, c2, wp, bias, y, n, c2, wp.shape[1], h, w, ksize=1)
    return y


def _heads(k: _Ctx, group, feat, n, hw):
    """All map2style heads (psp_encoders.py:13-37) that read one pyramid level: feat [n, hw, hw, C] -> codes [n, heads, 512]
    fp32. Level 0 is ONE GEMM with the heads' weights concatenated along O; deeper levels run the heads as extra batch entries
    (entry head * n + b) with one weight set and one bias per head."""
    nh, levels, slope, lin_w, lin_b, oc = group
    c_in = feat.shape[-1]
    w, bias = levels[0]
    xp = k.planes(feat, n, c_in, hw, hw)
    hw //= 2
    x = k.empty(n, hw, hw, nh * oc)
    k.conv(xp, c_in, w, bias, x, n, c_in, nh * oc, hw, hw, planes=1, act=1, slope=slope)
    for d in range(1, len(levels)):
        w, bias = levels[d]
        xp = k.planes(x, n, oc, hw, hw, heads=nh) if d == 1 else k.planes(x, nh * n, oc, hw, hw)
        hw //= 2
        x = k.empty(nh * n, hw, hw, oc)
        k.conv(xp, oc, w, bias, x, nh * n, oc, oc, hw, hw, planes=1, w_group=n, bias_per_set=1, act=1, slope=slope,
               round_y=0 if d == len(levels) - 1 else 1)
    if hw != 1 or len(levels) < 2:
        raise RuntimeError("fmi_b200: map2style heads must reduce their level to 1x1 in at least two convolutions")
    # EqualLinear (stylegan2/model.py:135-171) of every head: a plain batched library GEMM on [heads, n, 512] in fp32
    v = x.reshape(nh, n, oc).float()
    return torch.baddbmm(lin_b.view(nh, 1, oc), v, lin_w.transpose(1, 2)).transpose(0, 1)      # [n, heads, 512]


def encoder_forward(enc, x, ref=None, mask=None):
    """GradualStyleEncoder.forward (psp_encoders.py:100-152) -> codes [N, n_styles, 512] fp32."""
    k = _Ctx(x.device)
    cache = module_cache(enc).setdefault("psp_fast", {})
    key = (k.mma, x.device.index)
    ver = _versions(enc)
    ent = cache.get(key)
    if ent is None or ent[0] != ver:
        with torch.no_grad():
            ent = cache[key] = (ver, _Plan(enc, k.mma))
    plan = ent[1]
    n = x.shape[0]
    if ref is not None and mask is None:
        raise AssertionError("ref and mask should both be provided")
    img = x if ref is None else torch.cat([x, ref], dim=0)
    b, _, h, w = img.shape
    img = img.contiguous()
    a = torch.zeros((b, h, w, 32), dtype=k.dt, device=x.device)        # 3 channels zero-padded to one 16-byte-multiple row
    _lib.check(k.lib.fmi_nchw_to_nhwc_slice(_p(img), a.data_ptr(), b, 3, h, w, 32, _lib.F32, 1, k.mma, k.st), "fmi_nchw_to_nhwc_slice")
    t = k.empty(b, h, w, 64)
    k.conv(a, 32, plan.w_in, plan.b_in, t, b, 32, 64, h, w, act=4, slope_c=plan.slope_in)      # input_layer: conv + BN + PReLU
    del a
    taps = {}
    for i, u in enumerate(plan.units):
        t, h, w = _unit_forward(k, u, t, b, h, w)
        if i in _TAPS:
            taps[i] = (t, h, w)
    (f1, h1, w1), (f2, h2, w2), (f3, h3, w3) = taps[6], taps[20], taps[23]
    if ref is None:
        c1, c2, c3 = f1, f2, f3
    else:
        # attention / blend on the reference's NCHW fp32 interface (psp_encoders.py:127-138); the features are small
        mask_full = mask.unsqueeze(1)
        n1, n2, n3 = (_to_nchw(k, f, b, f.shape[-1], hh, ww) for f, hh, ww in ((f1, h1, w1), (f2, h2, w2), (f3, h3, w3)))
        if enc.use_attention:
            c3 = _attention_nhwc(k, plan, "attention1", enc.attention1, ops.scale_img(mask_full, (h3, w3)), n3[:n], n3[n:])
            c2 = _attention_nhwc(k, plan, "attention2", enc.attention2, ops.scale_img(mask_full, (h2, w2)), n2[:n], n2[n:])
        else:
            c3 = _to_nhwc(k, ops.composite(n3[:n].contiguous(), n3[n:].contiguous(), mask_full))
            c2 = _to_nhwc(k, ops.composite(n2[:n].contiguous(), n2[n:].contiguous(), mask_full))
        c1 = _to_nhwc(k, ops.composite(n1[:n].contiguous(), n1[n:].contiguous(), mask_full))
    codes = []
    if plan.heads[0] is not None:
        codes.append(_heads(k, plan.heads[0], c3, n, h3))
    # p2 = upsample(c3) + latlayer1(c2); p1 = upsample(p2) + latlayer2(c1)   (psp_encoders.py:144-149)
    l1 = k.empty(n, h2, w2, 512)
    k.conv(c2, c2.shape[-1], plan.lat1[0], plan.lat1[1], l1, n, c2.shape[-1], 512, h2, w2, ksize=1)
    p2 = k.empty(n, h2, w2, 512)
    _lib.check(k.lib.fmi_upsample_add_nhwc(c3.data_ptr(), l1.data_ptr(), p2.data_ptr(), n, 512, h3, w3, h2, w2, k.mma, k.st),
               "fmi_upsample_add_nhwc")
    if plan.heads[1] is not None:
        codes.append(_heads(k, plan.heads[1], p2, n, h2))
    l2 = k.empty(n, h1, w1, 512)
    k.conv(c1, c1.shape[-1], plan.lat2[0], plan.lat2[1], l2, n, c1.shape[-1], 512, h1, w1, ksize=1)
    p1 = k.empty(n, h1, w1, 512)
    _lib.check(k.lib.fmi_upsample_add_nhwc(p2.data_ptr(), l2.data_ptr(), p1.data_ptr(), n, 512, h2, w2, h1, w1, k.mma, k.st),
               "fmi_upsample_add_nhwc")
    if plan.heads[2] is not None:
        codes.append(_heads(k, plan.heads[2], p1, n, h1))
    return torch.cat(codes, dim=1)
