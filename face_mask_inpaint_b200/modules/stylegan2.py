"""Drop-in StyleGAN2 decoder modules (modules/psp/stylegan2/model.py) on the sm_100a kernels.

Same class names, constructor arguments, forward signatures and parameter/buffer names as the reference
(`conv.weight [1,O,I,k,k]`, `conv.modulation.{weight,bias}`, `conv.blur.kernel`, `noise.weight`, `activate.bias`,
`bias [1,3,1,1]`, `upsample.kernel`, `noises.noise_i`, `input.input`, `style.{1..n}.{weight,bias}`), so a reference
`state_dict` loads with strict=True (psp.py:55).

`Generator.forward` runs the whole synthesis network in the kernels' native layout (NHWC, tensor-core operand
type) and only returns to NCHW for the RGB image; the individual modules (ModulatedConv2d / StyledConv / ToRGB)
keep their NCHW in/out contract by converting at their boundary.
"""
from __future__ import annotations

import math
import random

import torch
from torch import nn
from torch.nn import functional as F

from .. import _lib, ops
from ..ops import FusedLeakyReLU, fused_leaky_relu, upfirdn2d  # noqa: F401  (re-exported like the reference's op package)


def _p(t):
    return None if t is None else t.data_ptr()


def _TORGB_FUSE_MIN_O():
    """Smallest output width whose ToRGB rides in the conv epilogue (A/B switch, FMI_TORGB_FUSE_MIN_O; default: always)."""
    import os
    return int(os.environ.get("FMI_TORGB_FUSE_MIN_O", "0"))


def _wants_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class PixelNorm(nn.Module):
    """model.py:10-16 (mapping network only; not on the hot path when input_is_latent=True)."""

    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    """model.py:19-27."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k /= k.sum()
    return k


class Upsample(nn.Module):
    """model.py:30-49."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel) * (factor ** 2)
        self.register_buffer('kernel', kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.factor, down=1, pad=self.pad)


class Downsample(nn.Module):
    """model.py:52-71."""

    def __init__(self, kernel, factor=2):
        super().__init__()
        self.factor = factor
        kernel = make_kernel(kernel)
        self.register_buffer('kernel', kernel)
        p = kernel.shape[0] - factor
        self.pad = ((p + 1) // 2, p // 2)

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=1, down=self.factor, pad=self.pad)


class Blur(nn.Module):
    """model.py:74-91."""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        kernel = make_kernel(kernel)
        if upsample_factor > 1:
            kernel = kernel * (upsample_factor ** 2)
        self.register_buffer('kernel', kernel)
        self.pad = pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, pad=self.pad)


class EqualLinear(nn.Module):
    """model.py:135-171. The dense product is a plain library GEMM (cuBLAS via F.linear); the fused bias +
    leaky-ReLU is the sm_100a kernel. (The per-layer `modulation` EqualLinear is evaluated by
    fmi_style_modulation inside ModulatedConv2d instead.)"""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init))
        else:
            self.bias = None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            out = F.linear(input, self.weight * self.scale)
            out = fused_leaky_relu(out, self.bias * self.lr_mul)
        else:
            out = F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)
        return out

    def __repr__(self):
        return f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})'


# ----------------------------------------------------------------------------------------------------
# native-layout helpers (NHWC, operand type)
# ----------------------------------------------------------------------------------------------------
def _op_dtype(mma):
    return torch.float32 if mma == _lib.MMA_TF32 else torch.bfloat16


def _to_nhwc_raw(x, mma):
    b, c, h, w = x.shape
    xc = x.contiguous()
    y = torch.empty((b, h, w, c), dtype=_op_dtype(mma), device=x.device)
    _lib.check(_lib.load().fmi_nchw_to_nhwc(_p(xc), _p(y), b, c, h, w, ops._dt(xc), mma, ops._stream()),
               "fmi_nchw_to_nhwc")
    return y


def _to_nchw_raw(y, mma, dtype):
    b, h, w, c = y.shape
    yc = y.contiguous()
    x = torch.empty((b, c, h, w), dtype=dtype, device=y.device)
    _lib.check(_lib.load().fmi_nhwc_to_nchw(_p(yc), _p(x), b, c, h, w, mma, ops._DT[dtype], ops._stream()),
               "fmi_nhwc_to_nchw")
    return x


class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mma):
        ctx.mma, ctx.dtype = mma, x.dtype
        return _to_nhwc_raw(x, mma)

    @staticmethod
    def backward(ctx, g):
        return _to_nchw_raw(g, ctx.mma, ctx.dtype), None


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, mma, dtype):
        ctx.mma = mma
        return _to_nchw_raw(y, mma, dtype)

    @staticmethod
    def backward(ctx, g):
        return _to_nhwc_raw(g, ctx.mma), None, None


def to_nhwc(x, mma):
    """NCHW (any supported dtype) -> NHWC in the tensor-core operand type; differentiable."""
    return _ToNHWC.apply(x, mma)


def to_nchw(y, mma, dtype):
    return _ToNCHW.apply(y, mma, dtype)


class _StyleModFn(torch.autograd.Function):
    """s = EqualLinear(style) (model.py:244, :159-167 with lr_mul = 1). The backward is three tiny dense products
    ([B,I] x [I,K]) done as library GEMMs, like EqualLinear's own F.linear."""

    @staticmethod
    def forward(ctx, style, mod_weight, mod_bias):
        st = style.float()
        if st.stride(-1) != 1:
            st = st.contiguous()
        b, k = st.shape
        i = mod_weight.shape[0]
        s = torch.empty((b, i), dtype=torch.float32, device=st.device)
        _lib.check(_lib.load().fmi_style_modulation(_p(st), st.stride(0), _p(mod_weight), _p(mod_bias), _p(s), b, k, i,
                                                    ops._stream()), "fmi_style_modulation")
        ctx.save_for_backward(st, mod_weight)
        ctx.style_dtype = style.dtype
        return s

    @staticmethod
    def backward(ctx, ds):
        st, mw = ctx.saved_tensors
        scale = 1.0 / math.sqrt(st.shape[1])
        dstyle = (ds @ mw) * scale if ctx.needs_input_grad[0] else None
        dmw = (ds.t() @ st) * scale if ctx.needs_input_grad[1] else None
        dmb = ds.sum(0) if ctx.needs_input_grad[2] else None
        return (dstyle.to(ctx.style_dtype) if dstyle is not None else None), dmw, dmb


def style_modulation(style, mod_weight, mod_bias):
    """s = EqualLinear(style) (model.py:244); style may be a strided row view latent[:, i]."""
    return _StyleModFn.apply(style, mod_weight, mod_bias)


def prep_weights(weight, s, demodulate, mma):
    """Per-sample modulated/demodulated weights in operand layout [B][tap][O][I] (model.py:245-252)."""
    _, o, i, k, _ = weight.shape
    b = s.shape[0]
    lib = _lib.load()
    nbytes = lib.fmi_modconv_weight_bytes(b, i, o, k, mma)
    wp = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    _lib.check(lib.fmi_modconv_weight_prep(_p(weight), _p(s), _p(wp), b, i, o, k, int(demodulate), mma, ops._stream()),
               "fmi_modconv_weight_prep")
    return wp


def styled_conv_nhwc(x, wp, o, upsample, act, mma, noise=None, noise_w=None, act_bias=None, blur_k=None):
    """x NHWC operand-type [B,H,W,I] -> [B,OH,OW,O]."""
    b, h, w, i = x.shape
    oh, ow = (2 * h, 2 * w) if upsample else (h, w)
    y = torch.empty((b, oh, ow, o), dtype=x.dtype, device=x.device)
    lib = _lib.load()
    ws_bytes = lib.fmi_styled_conv_workspace_bytes(b, o, h, w, int(upsample), mma)
    ws = ops._workspace(x.device, ws_bytes) if ws_bytes else None
    nb = 0
    if noise is not None:
        noise = noise.float().contiguous()
        nb = int(noise.shape[0] == b and b > 1) if noise.shape[0] != 1 else 0
        if noise.shape[0] not in (1, b):
            raise RuntimeError("noise batch must be 1 or B")
    _lib.check(lib.fmi_styled_conv_nhwc(_p(x), _p(wp), _p(y), _p(noise), nb, _p(noise_w), _p(act_bias), _p(blur_k), b, i,
                                        o, h, w, int(upsample), int(act), mma, _p(ws), ws.numel() if ws is not None else 0,
                                        ops._stream()), "fmi_styled_conv_nhwc")
    return y


def styled_conv_torgb_nhwc(x, wp, o, mma, noise, noise_w, act_bias, rgb_weight, rgb_s, rgb_bias, skip, up_kernel, want_y=True):
    """Plain StyledConv + the ToRGB that follows it, one kernel (fmi_styled_conv_torgb_nhwc; O <= 256).
    Returns (activations NHWC, rgb [B,3,H,W] fp32); with want_y=False the activations are not stored (None): the last layer
    of an inference forward has no other reader (at 1024^2, batch 8 that is a 512 MiB write)."""
    b, h, w, i = x.shape
    lib = _lib.load()
    y = torch.empty((b, h, w, o), dtype=x.dtype, device=x.device) if want_y else None
    rgb = torch.empty((b, 3, h, w), dtype=torch.float32, device=x.device)
    rgb_w = torch.empty((b, 3, o), dtype=torch.float32, device=x.device)
    _lib.check(lib.fmi_torgb_weights(_p(rgb_weight), _p(rgb_s), _p(rgb_w), b, o, ops._stream()), "fmi_torgb_weights")
    noise = noise.float().contiguous()
    nb = 1 if (noise.shape[0] == b and b > 1) else 0
    if skip is not None:
        skip = skip.float().contiguous()
    _lib.check(lib.fmi_styled_conv_torgb_nhwc(_p(x), _p(wp), _p(y), _p(noise), nb, _p(noise_w), _p(act_bias), _p(rgb_w),
                                              _p(rgb_bias), _p(skip), _p(up_kernel), _p(rgb), b, i, o, h, w, mma,
                                              ops._stream()), "fmi_styled_conv_torgb_nhwc")
    return y, rgb


def torgb_nhwc(x, weight, s, bias, skip, blur_k, mma):
    b, h, w, i = x.shape
    rgb = torch.empty((b, 3, h, w), dtype=torch.float32, device=x.device)
    if skip is not None:
        skip = skip.float().contiguous()
    _lib.check(_lib.load().fmi_torgb_nhwc(_p(x), _p(weight), _p(s), _p(bias), _p(skip), _p(blur_k), _p(rgb), b, i, h, w,
                                          mma, ops._stream()), "fmi_torgb_nhwc")
    return rgb


class _StyledConvFn(torch.autograd.Function):
    """fmi_styled_conv_nhwc / fmi_styled_conv_bwd_nhwc: the whole StyledConv (model.py:340-346) or, with act=False,
    ModulatedConv2d alone (:241-279) on NHWC operands, differentiable w.r.t. x, weight, s, noise_w and act_bias."""

    @staticmethod
    def forward(ctx, x, weight, s, noise, noise_w, act_bias, blur_k, demodulate, upsample, act, mma):
        o = weight.shape[1]
        x = x.contiguous()
        wp = prep_weights(weight, s, demodulate, mma)
        y = styled_conv_nhwc(x, wp, o, upsample, act, mma, noise, noise_w, act_bias, blur_k)
        if noise is not None:
            noise = noise.float().contiguous()
        ctx.save_for_backward(x, weight, s, noise, blur_k, y if act else None)
        ctx.cfg = (bool(demodulate), bool(upsample), bool(act), mma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, s, noise, blur_k, y = ctx.saved_tensors
        demodulate, upsample, act, mma = ctx.cfg
        b, h, w, i = x.shape
        o = weight.shape[1]
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dev = x.device
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty(weight.shape, dtype=torch.float32, device=dev)
        ds = torch.empty((b, i), dtype=torch.float32, device=dev)
        dnw = torch.empty(1, dtype=torch.float32, device=dev) if act else None
        dbias = torch.empty(o, dtype=torch.float32, device=dev) if act else None
        lib = _lib.load()
        nbytes = lib.fmi_styled_conv_bwd_workspace_bytes(b, i, o, h, w, int(upsample), int(act), mma)
        ws = ops._workspace(dev, nbytes)
        nb = int(noise is not None and noise.shape[0] == b and b > 1)
        _lib.check(lib.fmi_styled_conv_bwd_nhwc(_p(x), _p(y), _p(dy), _p(weight), _p(s), _p(noise), nb, _p(blur_k), _p(dx),
                                                _p(dw), _p(ds), _p(dnw), _p(dbias), b, i, o, h, w, int(upsample), int(act),
                                                int(demodulate), mma, ws.data_ptr(), ws.numel(), ops._stream()),
                   "fmi_styled_conv_bwd_nhwc")
        return dx, dw, ds, None, dnw, dbias, None, None, None, None, None


class _ToRGBFn(torch.autograd.Function):
    """fmi_torgb_nhwc without the skip branch / fmi_torgb_bwd_nhwc (model.py:360-364)."""

    @staticmethod
    def forward(ctx, x, weight, s, bias, mma):
        x = x.contiguous()
        ctx.save_for_backward(x, weight, s)
        ctx.mma = mma
        return torgb_nhwc(x, weight, s, bias, None, None, mma)

    @staticmethod
    def backward(ctx, drgb):
        x, weight, s = ctx.saved_tensors
        mma = ctx.mma
        b, h, w, c = x.shape
        dev = x.device
        lib = _lib.load()
        drgb = drgb.float().contiguous()
        rgb_w = torch.empty((b, 3, c), dtype=torch.float32, device=dev)
        _lib.check(lib.fmi_torgb_weights(_p(weight), _p(s), _p(rgb_w), b, c, ops._stream()), "fmi_torgb_weights")
        dx = torch.empty_like(x)
        d_rgbw = torch.empty((b, 3, c), dtype=torch.float32, device=dev)
        dbias = torch.empty(3, dtype=torch.float32, device=dev)
        _lib.check(lib.fmi_torgb_bwd_nhwc(_p(x), _p(drgb), _p(rgb_w), _p(dx), _p(d_rgbw), _p(dbias), b, c, h, w, mma,
                                          ops._stream()), "fmi_torgb_bwd_nhwc")
        # rgb_w[b,o,c] = W[o,c] s[b,c] / sqrt(C): [B,3,C]-sized products, host-side plumbing
        scale = 1.0 / math.sqrt(c)
        w2 = weight.reshape(3, c)
        dw = (torch.einsum('boc,bc->oc', d_rgbw, s) * scale).reshape(weight.shape)
        ds = torch.einsum('boc,oc->bc', d_rgbw, w2) * scale
        return dx, dw, ds, dbias.reshape(1, 3, 1, 1), None


# ----------------------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------------------
class ModulatedConv2d(nn.Module):
    """model.py:187-279."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        if downsample:
            factor = 2
            p = (len(blur_kernel) - factor) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        fan_in = in_channel * kernel_size ** 2
        self.scale = 1 / math.sqrt(fan_in)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, '
                f'upsample={self.upsample}, downsample={self.downsample})')

    def _check(self):
        if self.upsample and (self.blur.kernel.shape != (4, 4) or self.blur.pad != (1, 1)):
            raise NotImplementedError("fmi_b200: upsampling modulated conv supports the 4-tap blur with pad (1,1) only")

    def styles(self, style):
        return style_modulation(style, self.modulation.weight, self.modulation.bias)

    def forward_nhwc(self, x, s, mma, act=False, noise=None, noise_w=None, act_bias=None, wp=None):
        """x NHWC operand type; s the modulation [B,I] (from `styles`). `wp`: the per-sample weights already prepared
        (Generator.forward prepares them ahead of the layers on a side stream in inference); then `s` is not needed."""
        self._check()
        if wp is not None:
            return styled_conv_nhwc(x.contiguous(), wp, self.out_channel, self.upsample, act, mma, noise, noise_w, act_bias,
                                    self.blur.kernel if self.upsample else None)
        return _StyledConvFn.apply(x, self.weight, s, noise, noise_w, act_bias, self.blur.kernel if self.upsample else None,
                                   self.demodulate, self.upsample, act, mma)

    def forward(self, input, style):
        ops._need_cuda(input, style)
        mma = ops.mma_mode(input.dtype)
        s = self.styles(style)
        if self.kernel_size == 1:
            if self.out_channel != 3 or self.demodulate:
                raise NotImplementedError("fmi_b200: 1x1 modulated conv is implemented for ToRGB (3 channels, no demod)")
            zero = torch.zeros(3, device=input.device)
            return _ToRGBFn.apply(to_nhwc(input, mma), self.weight, s, zero, mma).to(input.dtype)
        if self.downsample:
            # model.py:265-273: blur with pad (2,2) -> [B,I,H+1,W+1], then conv2d(padding 0, stride 2, groups = batch).
            # Not reachable from the reference's scripts, so no dedicated kernel: the stride-2 valid conv is read out of the
            # stride-1 'same' implicit GEMM on the blurred input (valid output j = same output j + 1; 4x the FLOPs).
            if self.kernel_size != 3:
                raise NotImplementedError("fmi_b200: downsampling modulated conv supports 3x3 kernels")
            xb = self.blur(input)
            up, self.upsample = self.upsample, False
            try:
                y = self.forward_nhwc(to_nhwc(xb, mma), s, mma)
            finally:
                self.upsample = up
            return to_nchw(y, mma, input.dtype)[:, :, 1::2, 1::2].contiguous()
        y = self.forward_nhwc(to_nhwc(input, mma), s, mma)
        return to_nchw(y, mma, input.dtype)


class NoiseInjection(nn.Module):
    """model.py:282-294 (parameter holder; the addition is fused into the conv epilogue)."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        raise RuntimeError("fmi_b200: NoiseInjection is fused into StyledConv; it is not called on its own")


class ConstantInput(nn.Module):
    """model.py:297-308."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """model.py:311-346: modconv -> + noise_w * noise -> fused bias + leaky relu, as ONE epilogue."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def _noise(self, noise, b, h, w, device):
        if noise is None:  # model.py:290-292: drawn with PyTorch's generator, one plane per sample
            noise = torch.empty(b, 1, h, w, device=device).normal_()
        return noise

    def forward_nhwc(self, x, style, mma, noise=None, wp=None):
        b, h, w, _ = x.shape
        oh, ow = (2 * h, 2 * w) if self.conv.upsample else (h, w)
        noise = self._noise(noise, b, oh, ow, x.device)
        s = self.conv.styles(style) if wp is None else None
        return self.conv.forward_nhwc(x, s, mma, act=True, noise=noise, noise_w=self.noise.weight,
                                      act_bias=self.activate.bias, wp=wp)

    def forward(self, input, style, noise=None):
        ops._need_cuda(input, style)
        mma = ops.mma_mode(input.dtype)
        y = self.forward_nhwc(to_nhwc(input, mma), style, mma, noise)
        return to_nchw(y, mma, input.dtype)


class ToRGB(nn.Module):
    """model.py:349-369: 1x1 modconv (no demod) + bias + upsampled skip in one kernel."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward_nhwc(self, x, style, mma, skip=None):
        s = self.conv.styles(style)
        if _wants_grad(x, s, skip, self.conv.weight, self.bias):
            # training: the 1x1 modulated conv through its own backward kernel, the skip through upfirdn2d's
            rgb = _ToRGBFn.apply(x, self.conv.weight, s, self.bias, mma)
            if skip is not None:
                rgb = rgb + self.upsample(skip.float())
            return rgb
        blur_k = None
        if skip is not None:
            if self.upsample.kernel.shape != (4, 4) or self.upsample.pad != (2, 1):
                raise NotImplementedError("fmi_b200: ToRGB skip upsampling supports the 4-tap kernel only")
            blur_k = self.upsample.kernel
        return torgb_nhwc(x, self.conv.weight, s, self.bias, skip, blur_k, mma)

    def forward(self, input, style, skip=None):
        ops._need_cuda(input, style)
        mma = ops.mma_mode(input.dtype)
        return self.forward_nhwc(to_nhwc(input, mma), style, mma, skip).to(input.dtype)


class Generator(nn.Module):
    """model.py:372-550."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        self.size = size
        self.style_dim = style_dim
        layers = [PixelNorm()]
        for i in range(n_mlp):
            layers.append(EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation='fused_lrelu'))
        self.style = nn.Sequential(*layers)
        self.channels = {
            4: 512, 8: 512, 16: 512, 32: 512,
            64: 256 * channel_multiplier, 128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
            512: 32 * channel_multiplier, 1024: 16 * channel_multiplier,
        }
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[4]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f'noise_{layer_idx}', torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True, blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2

    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 2 ** 2, 2 ** 2, device=device)]
        for i in range(3, self.log_size + 1):
            for _ in range(2):
                noises.append(torch.randn(1, 1, 2 ** i, 2 ** i, device=device))
        return noises

    def mean_latent(self, n_latent):
        latent_in = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(latent_in).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def _prepare_ahead(self, lat, mma):
        """[(wp, event)] for conv1 and every entry of self.convs, computed on a side stream that first waits for the current
        one (so its allocator blocks are only reused after every earlier consumer was enqueued). StyledConv k uses lat[:, k]
        for k = 0 (conv1) and lat[:, k] for convs[k-1] (model.py:529-538). FMI_SG2_PREP_AHEAD=0 switches it off."""
        import os
        if os.environ.get("FMI_SG2_PREP_AHEAD") == "0" or not lat.is_cuda:
            return None
        from ..graphs import module_cache
        cache = module_cache(self)
        side = cache.get("side_stream")
        if side is None or side.device != lat.device:
            side = cache["side_stream"] = torch.cuda.Stream(device=lat.device)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        out = []
        with torch.cuda.stream(side):
            for k, layer in enumerate([self.conv1] + list(self.convs)):
                conv = layer.conv
                conv._check()
                wp = prep_weights(conv.weight, conv.styles(lat[:, k]), conv.demodulate, mma)
                ev = torch.cuda.Event()
                ev.record(side)
                out.append((wp, ev))
        return out

    def forward(self, styles, return_latents=False, return_features=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if noise is None:
            if randomize_noise:
                noise = [None] * self.num_layers
            else:
                noise = [getattr(self.noises, f'noise_{i}') for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent + truncation * (style - truncation_latent) for style in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            if styles[0].ndim < 3:
                latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
            else:
                latent = styles[0]
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1)
            latent2 = styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)
            latent = torch.cat([latent, latent2], 1)

        ops._need_cuda(latent)
        in_dtype = latent.dtype
        mma = ops.mma_mode(in_dtype)
        train = _wants_grad(latent, *self.parameters())
        lat = latent.float()
        # synthesis (model.py:528-541) in NHWC operand layout
        # Inference: the per-sample weights of every StyledConv depend on the latent only, so they are prepared ahead of
        # the layers on a side stream (style modulation + weight_prep: 0.75 ms of a 5.3 ms batch-8 forward, and the early
        # 512-channel layers are nothing but weight preparation); each layer waits for its own event. Under CUDA-graph
        # capture this is a parallel branch of the graph.
        ahead = self._prepare_ahead(lat, mma) if not train else None

        def pre(k):   # prepared weights of StyledConv k (0 = conv1, 1.. = convs[k-1]) or None
            if ahead is None:
                return None
            torch.cuda.current_stream().wait_event(ahead[k][1])
            return ahead[k][0]

        out = to_nhwc(self.input(lat), mma)
        out = self.conv1.forward_nhwc(out, lat[:, 0], mma, noise=noise[0], wp=pre(0))
        skip = self.to_rgb1.forward_nhwc(out, lat[:, 1], mma)
        i = 1
        for conv1, conv2, noise1, noise2, to_rgb in zip(self.convs[::2], self.convs[1::2], noise[1::2], noise[2::2],
                                                        self.to_rgbs):
            out = conv1.forward_nhwc(out, lat[:, i], mma, noise=noise1, wp=pre(i))
            if _TORGB_FUSE_MIN_O() <= conv2.conv.out_channel <= 256 and not conv2.conv.upsample and not train:
                # conv2 + ToRGB in one kernel (the RGB projection rides in conv2's epilogue)
                b, h, w, _ = out.shape
                conv2.conv._check()
                wp = pre(i + 1)
                if wp is None:
                    wp = prep_weights(conv2.conv.weight, conv2.conv.styles(lat[:, i + 1]), conv2.conv.demodulate, mma)
                last = to_rgb is self.to_rgbs[-1] and not return_features
                out, skip = styled_conv_torgb_nhwc(
                    out, wp, conv2.conv.out_channel, mma, conv2._noise(noise2, b, h, w, out.device), conv2.noise.weight,
                    conv2.activate.bias, to_rgb.conv.weight, to_rgb.conv.styles(lat[:, i + 2]), to_rgb.bias, skip,
                    to_rgb.upsample.kernel, want_y=not last)
            else:
                out = conv2.forward_nhwc(out, lat[:, i + 1], mma, noise=noise2, wp=pre(i + 1))
                skip = to_rgb.forward_nhwc(out, lat[:, i + 2], mma, skip)
            i += 2
        image = skip.to(in_dtype) if in_dtype != torch.float32 else skip
        if return_latents:
            return image, latent
        elif return_features:
            return image, to_nchw(out, mma, in_dtype)
        else:
            return image, None
