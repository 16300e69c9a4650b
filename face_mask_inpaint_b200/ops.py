"""Autograd wrappers over the C ABI (include/fmi_b200.h) that mirror the reference's op-level API:

  fused_leaky_relu / FusedLeakyReLU      modules/psp/stylegan2/op/fused_act.py:72-85
  upfirdn2d                              modules/psp/stylegan2/op/upfirdn2d.py:142-147
  scale_img / composite                  modules/model.py:10-12, :99 ; psp_encoders.py:135-138

Same names, argument meaning and error behaviour (RuntimeError for non-CUDA tensors, as the reference's
CHECK_CUDA, fused_bias_act.cpp:7-16). PyTorch is used for device memory and streams only; all arithmetic
happens in libfmi_b200.so. There is no CPU path.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"fmi_b200: unsupported dtype {t.dtype} (fp32 / bf16 / fp16 only)") from None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fmi_b200: tensor must be a CUDA tensor (there is no CPU fallback)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# fused bias + activation
# ------------------------------------------------------------------------------------------------
def fused_bias_act(input, bias, refer, act, grad, alpha, scale):
    """Drop-in for the pybind op `fused.fused_bias_act` (op/fused_bias_act.cpp:11-21): empty tensors mean
    'absent', a fresh output tensor is returned, nothing is mutated."""
    _need_cuda(input, bias, refer)
    x = input.contiguous()
    b = bias.contiguous() if bias is not None and bias.numel() else None
    r = refer.contiguous() if refer is not None and refer.numel() else None
    if r is not None and r.dtype != x.dtype:
        r = r.to(x.dtype)
    if b is not None and b.dtype != x.dtype and b.dtype != torch.float32:
        b = b.float()
    y = torch.empty_like(x)
    step_b = 1
    for i in range(2, x.dim()):
        step_b *= x.size(i)
    size_b = b.numel() if b is not None else 1
    lib = _lib.load()
    _lib.check(lib.fmi_fused_bias_act(_ptr(x), _ptr(b), _ptr(r), _ptr(y), int(act), int(grad), float(alpha),
                                      float(scale), x.numel(), step_b, size_b, _dt(x),
                                      _dt(b) if b is not None else _dt(x), _stream()), "fmi_fused_bias_act")
    return y


class FusedLeakyReLUFunctionBackward(Function):
    """op/fused_act.py:18-47 with the bias-gradient reduction fused into the kernel."""

    @staticmethod
    def forward(ctx, grad_output, out, negative_slope, scale, bias_numel):
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        g = grad_output.contiguous()
        if g.dtype != out.dtype:
            g = g.to(out.dtype)
        grad_input = torch.empty_like(g)
        grad_bias = torch.zeros(bias_numel, dtype=torch.float32, device=g.device)
        step_b = 1
        for i in range(2, g.dim()):
            step_b *= g.size(i)
        lib = _lib.load()
        _lib.check(lib.fmi_bias_act_bwd(_ptr(g), _ptr(out), _ptr(grad_input), _ptr(grad_bias), float(negative_slope),
                                        float(scale), g.numel(), step_b, bias_numel, _dt(g), _stream()),
                   "fmi_bias_act_bwd")
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        gradgrad_out = fused_bias_act(gradgrad_input, gradgrad_bias.to(gradgrad_input.dtype), out, 3, 1,
                                      ctx.negative_slope, ctx.scale)
        return gradgrad_out, None, None, None, None


class FusedLeakyReLUFunction(Function):
    """op/fused_act.py:50-69."""

    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        out = fused_bias_act(input, bias, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.negative_slope = negative_slope
        ctx.scale = scale
        ctx.bias_numel = bias.numel()
        ctx.bias_dtype = bias.dtype
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale,
                                                                     ctx.bias_numel)
        return grad_input, grad_bias.to(ctx.bias_dtype), None, None


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:84-85."""
    return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    """op/fused_act.py:72-81 — same parameter name (`bias`) and shape."""

    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


# ------------------------------------------------------------------------------------------------
# upfirdn2d
# ------------------------------------------------------------------------------------------------
def upfirdn2d_op(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """Drop-in for the pybind op `upfirdn2d_op.upfirdn2d` (op/upfirdn2d.cpp:12-23): input is
    [major, in_h, in_w, minor]; returns [major, out_h, out_w, minor]."""
    _need_cuda(input, kernel)
    x = input.contiguous()
    k = kernel.contiguous().float()
    major, in_h, in_w, minor = x.shape
    kh, kw = k.shape
    lib = _lib.load()
    out_h = lib.fmi_upfirdn2d_out_size(in_h, up_y, down_y, pad_y0, pad_y1, kh)
    out_w = lib.fmi_upfirdn2d_out_size(in_w, up_x, down_x, pad_x0, pad_x1, kw)
    if out_h < 1 or out_w < 1:
        raise RuntimeError(f"upfirdn2d: empty output extent ({out_h} x {out_w})")
    y = torch.empty((major, out_h, out_w, minor), dtype=x.dtype, device=x.device)
    _lib.check(lib.fmi_upfirdn2d(_ptr(x), _ptr(k), _ptr(y), major, in_h, in_w, minor, kh, kw, up_x, up_y, down_x,
                                 down_y, pad_x0, pad_x1, pad_y0, pad_y1, _dt(x), _stream()), "fmi_upfirdn2d")
    return y


class UpFirDn2dBackward(Function):
    """op/upfirdn2d.py:17-82."""

    @staticmethod
    def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
        up_x, up_y = up
        down_x, down_y = down
        g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1 = g_pad
        grad_output = grad_output.reshape(-1, out_size[0], out_size[1], 1)
        grad_input = upfirdn2d_op(grad_output, grad_kernel, down_x, down_y, up_x, up_y, g_pad_x0, g_pad_x1, g_pad_y0,
                                  g_pad_y1)
        grad_input = grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])
        ctx.save_for_backward(kernel)
        ctx.up, ctx.down, ctx.pad = up, down, pad
        ctx.in_size, ctx.out_size = in_size, out_size
        return grad_input

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        gradgrad_input = gradgrad_input.reshape(-1, ctx.in_size[2], ctx.in_size[3], 1)
        gradgrad_out = upfirdn2d_op(gradgrad_input, kernel, ctx.up[0], ctx.up[1], ctx.down[0], ctx.down[1], *ctx.pad)
        gradgrad_out = gradgrad_out.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1])
        return gradgrad_out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):
    """op/upfirdn2d.py:85-139."""

    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kernel_h, kernel_w = kernel.shape
        batch, channel, in_h, in_w = input.shape
        ctx.in_size = input.shape
        input = input.reshape(-1, in_h, in_w, 1)
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
        out_h = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) // down_y + 1
        out_w = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) // down_x + 1
        ctx.out_size = (out_h, out_w)
        ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
        g_pad_x0 = kernel_w - pad_x0 - 1
        g_pad_y0 = kernel_h - pad_y0 - 1
        g_pad_x1 = in_w * up_x - out_w * down_x + pad_x0 - up_x + 1
        g_pad_y1 = in_h * up_y - out_h * down_y + pad_y0 - up_y + 1
        ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
        out = upfirdn2d_op(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)
        return out.view(-1, channel, out_h, out_w)

    @staticmethod
    def backward(ctx, grad_output):
        kernel, grad_kernel = ctx.saved_tensors
        grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad,
                                             ctx.in_size, ctx.out_size)
        return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """op/upfirdn2d.py:142-147."""
    return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))


# ------------------------------------------------------------------------------------------------
# compositing
# ------------------------------------------------------------------------------------------------
def scale_img(img, size):
    """modules/model.py:10-12 for a single-channel mask [N,1,Hm,Wm] (bilinear, align_corners=True)."""
    _need_cuda(img)
    if img.dim() != 4 or img.size(1) != 1:
        raise RuntimeError("fmi_b200.scale_img: expects a [N,1,H,W] mask")
    m = img.contiguous().float()
    n, _, hm, wm = m.shape
    h, w = int(size[0]), int(size[1])
    out = torch.empty((n, 1, h, w), dtype=torch.float32, device=m.device)
    _lib.check(_lib.load().fmi_scale_mask(_ptr(m), _ptr(out), n, hm, wm, h, w, _stream()), "fmi_scale_mask")
    return out.to(img.dtype)


class _SpectralNormWeight(Function):
    """w = w_bar / sigma(w_bar) with one in-place power iteration on (u, v): SpectralNorm._update_u_v
    (modules/pluralistic_model/external_function.py:30-42) as 3 kernels forward, 2 backward."""

    @staticmethod
    def forward(ctx, w_bar, u, v):
        hh = w_bar.shape[0]
        wd = w_bar.numel() // hh
        w_out = torch.empty_like(w_bar)
        snap = torch.empty(hh + wd + 1, dtype=torch.float32, device=w_bar.device)
        scratch = torch.empty(hh + wd, dtype=torch.float32, device=w_bar.device)
        _lib.check(_lib.load().fmi_spectral_norm_fwd(_ptr(w_bar), _ptr(u), _ptr(v), _ptr(scratch), _ptr(w_out), _ptr(snap), hh, wd,
                                                     _stream()), "fmi_spectral_norm_fwd")
        ctx.save_for_backward(w_out, snap)
        ctx.mark_non_differentiable(u, v)
        return w_out

    @staticmethod
    def backward(ctx, g):
        w_out, snap = ctx.saved_tensors
        hh = w_out.shape[0]
        wd = w_out.numel() // hh
        g = g.contiguous().float()
        grad = torch.empty_like(w_out)
        dot = torch.empty(1, dtype=torch.float32, device=g.device)
        _lib.check(_lib.load().fmi_spectral_norm_bwd(_ptr(g), _ptr(w_out), _ptr(snap), _ptr(dot), _ptr(grad), hh, wd, _stream()),
                   "fmi_spectral_norm_bwd")
        return grad, None, None


def spectral_norm_weight(w_bar, u, v):
    """The normalised weight of a SpectralNorm-wrapped layer; `u`, `v` (their storage) advance by one power iteration. CUDA
    fp32 contiguous tensors only (no CPU path)."""
    _need_cuda(w_bar, u, v)
    if w_bar.dtype != torch.float32 or not (w_bar.is_contiguous() and u.is_contiguous() and v.is_contiguous()):
        raise RuntimeError("fmi_b200.spectral_norm_weight: contiguous fp32 tensors only")
    return _SpectralNormWeight.apply(w_bar, u.data, v.data)


def adaptive_avg_pool(img, size):
    """nn.AdaptiveAvgPool2d(size) of an image batch (psp.py:33,113-114 `face_pool`; model.py:79,111). The exact k x k case
    (k = 2 or 4, fp32, inference) is fmi_avgpool_planes; anything else — other ratios, autograd — is ATen's own pooling (metric-
    side resizing, not the hot path)."""
    h, w = int(size[0]), int(size[1])
    n, c, ih, iw = img.shape
    k = ih // h if h else 0
    if (img.is_cuda and img.dtype == torch.float32 and not (torch.is_grad_enabled() and img.requires_grad) and k in (2, 4)
            and ih == k * h and iw == k * w and iw % 4 == 0):
        x = img.contiguous()
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=img.device)
        _lib.check(_lib.load().fmi_avgpool_planes(_ptr(x), _ptr(out), n * c, ih, iw, k, _stream()), "fmi_avgpool_planes")
        return out
    return torch.nn.functional.adaptive_avg_pool2d(img, (h, w))


class _Composite(Function):
    @staticmethod
    def forward(ctx, src, ref, mask_full):
        _need_cuda(src, ref, mask_full)
        s = src.contiguous()
        r = ref.contiguous().to(s.dtype)
        m = mask_full.contiguous().float()
        if m.dim() == 3:
            m = m.unsqueeze(1)
        n, c, h, w = s.shape
        out = torch.empty_like(s)
        _lib.check(_lib.load().fmi_composite(_ptr(s), _ptr(r), _ptr(m), _ptr(out), n, c, h, w, m.size(-2), m.size(-1),
                                             _dt(s), _stream()), "fmi_composite")
        ctx.save_for_backward(m)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        m, = ctx.saved_tensors
        g = grad_out.contiguous()
        n, c, h, w = g.shape
        gs = torch.empty_like(g) if ctx.needs_input_grad[0] else None
        gr = torch.empty_like(g) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.load().fmi_composite_bwd(_ptr(g), _ptr(m), _ptr(gs), _ptr(gr), n, c, h, w, m.size(-2),
                                                 m.size(-1), _dt(g), _stream()), "fmi_composite_bwd")
        return gs, gr, None


def composite(src, ref, mask_full):
    """(1 - m) * src + m * ref with m = scale_img(mask_full, src.shape[-2:]) fused in one pass
    (modules/model.py:98-99; psp_encoders.py:127-138). mask_full is the [N,1,Hm,Wm] (or [N,Hm,Wm]) mask."""
    return _Composite.apply(src, ref, mask_full)


# ------------------------------------------------------------------------------------------------
# attention (a1 ExampleGuidedAttention, a2 Auto_Attn) and the 1x1 convolutions around it
# ------------------------------------------------------------------------------------------------
import os

_WS = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only per-device scratch (operand staging for the tensor-core kernels); stream-ordered reuse. The returned view
    starts on a 1024-byte boundary (TMA / SWIZZLE_128B staging; the caching allocator only guarantees 512)."""
    key = (device.type, device.index, torch.cuda.current_stream().cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes + 1024:
        buf = torch.empty(max(nbytes, 1 << 20) + 1024, dtype=torch.uint8, device=device)
        _WS[key] = buf
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + max(nbytes, buf.numel() - 1024)]


def mma_mode(dtype: torch.dtype) -> int:
    """Tensor-core operand precision: TF32 for the fp32 contract (<=1e-3), bf16 for the bf16 contract (<=2e-2).
    FMI_PRECISION=bf16 forces bf16 operands for fp32 I/O (SURVEY.md §5 'Config / flags')."""
    env = os.environ.get("FMI_PRECISION", "").lower()
    if env == "bf16":
        return _lib.MMA_BF16
    if env in ("fp32", "tf32", "tf32x3"):
        return _lib.MMA_TF32
    return _lib.MMA_TF32 if dtype == torch.float32 else _lib.MMA_BF16


def tf32_split() -> bool:
    """True when fp32 convolutions run with error-compensated 3xTF32 operands (fmi_tf32_split3) instead of single TF32: the
    caller switched TF32 convolutions off (torch.backends.cudnn.allow_tf32 = False — the switch that makes the reference's own
    GPU run strict fp32), or FMI_PRECISION=tf32x3. FMI_PRECISION=bf16 / tf32 / fp32 pin the single-pass modes."""
    env = os.environ.get("FMI_PRECISION", "").lower()
    if env == "tf32x3":
        return True
    if env in ("bf16", "tf32", "fp32"):
        return False
    return not torch.backends.cudnn.allow_tf32


def conv1x1(x, weight, bias=None):
    """1x1 nn.Conv2d forward (example_guided_att.py:9,13) on [N,C,H,W] through fmi_conv1x1."""
    _need_cuda(x, weight, bias)
    xc = x.contiguous()
    n, cin = xc.shape[0], xc.shape[1]
    cout = weight.shape[0]
    s = xc[0, 0].numel()
    w = weight.reshape(cout, cin).contiguous().float()
    b = bias.contiguous().float() if bias is not None else None
    y = torch.empty((n, cout) + tuple(xc.shape[2:]), dtype=xc.dtype, device=xc.device)
    _lib.check(_lib.load().fmi_conv1x1(_ptr(xc), _ptr(w), _ptr(b), _ptr(y), n, cin, cout, s, _dt(xc), _stream()),
               "fmi_conv1x1")
    return y


# ------------------------------------------------------------------------------------------------
# batch-shared convolutions of the PICNet conv blocks in TRAINING (SURVEY 8f rank 1 + row g)
# ------------------------------------------------------------------------------------------------
def _as_nhwc(t, fresh=False):
    """[N,C,H,W] fp32 -> dense [N,H,W,C]: a view when `t` already is channels_last in memory, else one copy. A block input or a
    gradient in NCHW layout is read by two Functions (main path and shortcut), which would otherwise transpose it twice: the copy
    is remembered on the tensor (keyed by its version counter) — but only for tensors that live for ONE iteration (`fresh`: a
    gradient handed to backward; or an intermediate with a grad_fn). A leaf that persists across iterations (a model input, the
    static input buffer of a captured CUDA graph, rewritten in place before every replay) is converted every time."""
    v = t.permute(0, 2, 3, 1)
    if v.is_contiguous() and t.dtype == torch.float32:
        return v
    if not (fresh or t.grad_fn is not None):
        return v.contiguous().float()
    hit = getattr(t, "_fmi_nhwc", None)
    if hit is not None and hit[0] == t._version:
        return hit[1]
    out = v.contiguous().float()
    try:
        t._fmi_nhwc = (t._version, out)
    except (AttributeError, RuntimeError):
        pass
    return out


def _conv_wp(weight, o, i, k, transposed, st):
    """[k*k][o][i] operand-layout weights (tf32-rounded) of `weight` [o,i,k,k] (transposed = 0) or [i,o,k,k] (= 1)."""
    wp = torch.empty((k * k, o, i), dtype=torch.float32, device=weight.device)
    _lib.check(_lib.load().fmi_conv_weight_prep(_ptr(weight), _ptr(wp), o, i, transposed, o, i, 0, 0, k, _lib.MMA_TF32, st),
               "fmi_conv_weight_prep")
    return wp


def _conv_nhwc_call(x, i, wp, bias, o, b, h, w, k, st):
    y = torch.empty((b, h, w, o), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().fmi_conv_nhwc(_ptr(x), i, w * i, h * w * i, _ptr(wp), _ptr(bias), 1, None, 0.0, _ptr(y), o, b, i, o, h, w,
                                         k, 0, 0, 0, 2, 0, 0, _lib.MMA_TF32, st), "fmi_conv_nhwc")
    return y


class _ConvShared(Function):
    """nn.Conv2d(k in {1, 3}, stride 1, padding k // 2) with batch-shared weights, forward AND backward on the tcgen05 kernels
    (the reference: F.conv2d + cudnn_convolution_backward behind SpectralNorm.forward, external_function.py:70-72). Tensors stay
    channels_last in memory: the [N,H,W,C] buffers of the kernels are returned as [N,C,H,W] views, so the LeakyReLU / pooling /
    residual adds between two convolutions run on that layout and nothing is transposed. TF32 operands (the reference's own
    GPU numerics under torch.backends.cudnn.allow_tf32), fp32 accumulation and storage.
      forward   fmi_conv_nhwc (implicit GEMM, bias in the epilogue)
      dx        the same kernel on the flipped, transposed weights
      dweight   fmi_conv_wgrad_nhwc (pixel-contraction GEMM, split-K, fp32 red.add)
      dbias     one reduction of dy"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        st = _stream()
        o, i, k, _ = weight.shape
        b, _, h, w = x.shape
        xn = _as_nhwc(x)
        wc = weight.detach().contiguous().float()
        bc = None if bias is None else bias.detach().contiguous().float()
        y = _conv_nhwc_call(xn, i, _conv_wp(wc, o, i, k, 0, st), bc, o, b, h, w, k, st)
        ctx.save_for_backward(xn, wc)
        ctx.has_bias = bias is not None
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, wc = ctx.saved_tensors
        st = _stream()
        o, i, k, _ = wc.shape
        b, h, w, _ = xn.shape
        g = _as_nhwc(gy, fresh=True)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wf = wc.flip(2, 3).contiguous() if k == 3 else wc
            dx = _conv_nhwc_call(g, o, _conv_wp(wf, i, o, k, 1, st), None, i, b, h, w, k, st).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dwp = torch.zeros((k * k, o, i), dtype=torch.float32, device=g.device)
            _lib.check(_lib.load().fmi_conv_wgrad_nhwc(_ptr(xn), _ptr(g), _ptr(dwp), b, i, o, h, w, k, 0, _lib.MMA_TF32, st),
                       "fmi_conv_wgrad_nhwc")
            dw = dwp.permute(1, 2, 0).reshape(o, i, k, k)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _channel_sum(g)
        return dx, dw, db


def _channel_sum(g):
    """sum over (batch, rows, columns) of a dense NHWC fp32 tensor — a bias gradient — with the streaming statistics kernel of
    InstanceNorm (fmi_instnorm_stats_nhwc: per-(image, channel) sums in double, 0.89 of the HBM rate) instead of ATen's
    column reduction."""
    b, h, w, c = g.shape
    if c % 4 or c > 1024 or os.environ.get("FMI_CONV_TRAIN_DBIAS") == "0":
        return g.sum((0, 1, 2))
    ss = torch.empty((b, c, 2), dtype=torch.float32, device=g.device)
    sums = torch.empty((b, c, 2), dtype=torch.float64, device=g.device)
    _lib.check(_lib.load().fmi_instnorm_stats_nhwc(_ptr(g), c, None, None, _ptr(ss), _ptr(sums), b, c, h * w, 1e-5, _lib.MMA_TF32,
                                                   _stream()), "fmi_instnorm_stats_nhwc")
    return sums[..., 0].sum(0).float()


class _ConvTShared(Function):
    """nn.ConvTranspose2d(3, stride 2, padding 1, output_padding 1) with batch-shared weights [I,O,3,3] (ResBlockDecoder's conv2 /
    bypass, base_function.py:330-336), forward and backward on the kernels, channels_last in memory like `_ConvShared`:
      forward   fmi_conv3x3_nhwc mode 3 (O <= 32: one GEMM over the 4 output-parity classes) or mode 2 (one GEMM per class)
      dx        the gradient's 4 pixel-parity planes (fmi_space_to_planes_nhwc) through the stride-2 implicit GEMM
                fmi_conv_nhwc(planes = 1) — dx[m] = sum_ky dy[2m - 1 + ky] W[ky] is Conv2d(O -> I, 3, stride 2, padding 1) with W as it lies
      dweight   fmi_conv_wgrad_nhwc(transposed = 1) on the same planes
      dbias     one reduction of dy"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        st = _stream()
        lib = _lib.load()
        i, o = weight.shape[0], weight.shape[1]
        b, _, h, w = x.shape
        xn = _as_nhwc(x)
        wc = weight.detach().contiguous().float()
        bc = None if bias is None else bias.detach().contiguous().float()
        merged = o <= 32
        wp = (torch.zeros if merged else torch.empty)((4, 4 * o, i) if merged else (9, o, i), dtype=torch.float32, device=x.device)
        _lib.check(lib.fmi_conv_weight_prep(_ptr(wc), _ptr(wp), o, i, 1, 4 * o if merged else o, i, 0, int(merged), 3,
                                            _lib.MMA_TF32, st), "fmi_conv_weight_prep")
        y = torch.empty((b, 2 * h, 2 * w, o), dtype=torch.float32, device=x.device)
        _lib.check(lib.fmi_conv3x3_nhwc(_ptr(xn), i, _ptr(wp), _ptr(bc), _ptr(y), o, 0, None, 0, b, i, o, h, w, 3 if merged else 2, 2,
                                        0.0, 0, _lib.MMA_TF32, st), "fmi_conv3x3_nhwc")
        ctx.save_for_backward(xn, wc)
        ctx.has_bias = bias is not None
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, wc = ctx.saved_tensors
        st = _stream()
        lib = _lib.load()
        i, o = wc.shape[0], wc.shape[1]
        b, h, w, _ = xn.shape
        g = _as_nhwc(gy, fresh=True)
        dx = dw = db = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            planes = torch.empty((4 * b, h, w, o), dtype=torch.float32, device=g.device)
            _lib.check(lib.fmi_space_to_planes_nhwc(_ptr(g), o, _ptr(planes), b, o, 2 * h, 2 * w, 1, _lib.MMA_TF32, st),
                       "fmi_space_to_planes_nhwc")
        if ctx.needs_input_grad[0]:
            wp = _conv_wp(wc, i, o, 3, 0, st)      # W [I,O,3,3] read as the Conv2d weight [out = I][in = O]
            dxn = torch.empty((b, h, w, i), dtype=torch.float32, device=g.device)
            _lib.check(lib.fmi_conv_nhwc(_ptr(planes), o, w * o, h * w * o, _ptr(wp), None, 1, None, 0.0, _ptr(dxn), i, b, o, i, h, w,
                                         3, 1, 0, 0, 2, 0, 0, _lib.MMA_TF32, st), "fmi_conv_nhwc")
            dx = dxn.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dwp = torch.zeros((9, o, i), dtype=torch.float32, device=g.device)
            _lib.check(lib.fmi_conv_wgrad_nhwc(_ptr(xn), _ptr(planes), _ptr(dwp), b, i, o, h, w, 3, 1, _lib.MMA_TF32, st),
                       "fmi_conv_wgrad_nhwc")
            dw = dwp.permute(2, 1, 0).reshape(i, o, 3, 3)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _channel_sum(g)
        return dx, dw, db


class _NormAct(Function):
    """leaky_relu(InstanceNorm2d(x)) — the norm + activation pairs of ResBlockDecoder (base_function.py:338-344) — forward
    (fmi_instnorm_stats_nhwc + fmi_norm_act_nhwc: one statistics pass, one normalise + activate pass) and backward
    (fmi_instnorm_act_bwd_nhwc: two passes over (dy, x)) on channels_last tensors; the output is tf32-rounded, which is what
    the convolution that reads it would make of it anyway."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, slope):
        st = _stream()
        lib = _lib.load()
        b, c, h, w = x.shape
        hw = h * w
        xn = _as_nhwc(x)
        y = torch.empty((b, h, w, c), dtype=torch.float32, device=x.device)
        ctx.slope, ctx.affine, ctx.normed = float(slope), (gamma is not None, beta is not None), eps is not None
        if eps is None:     # activation only (blocks without normalisation): y = tf32(lrelu(x))
            _lib.check(lib.fmi_norm_act_nhwc(_ptr(xn), c, _ptr(y), c, None, b, c, hw, float(slope), _lib.MMA_TF32, st),
                       "fmi_norm_act_nhwc")
            ctx.save_for_backward(xn)
            return y.permute(0, 3, 1, 2)
        ga = None if gamma is None else gamma.detach().contiguous().float()
        be = None if beta is None else beta.detach().contiguous().float()
        ss = torch.empty((b, c, 2), dtype=torch.float32, device=x.device)
        sums = torch.empty((b, c, 2), dtype=torch.float64, device=x.device)
        _lib.check(lib.fmi_instnorm_stats_nhwc(_ptr(xn), c, _ptr(ga), _ptr(be), _ptr(ss), _ptr(sums), b, c, hw, float(eps),
                                               _lib.MMA_TF32, st), "fmi_instnorm_stats_nhwc")
        _lib.check(lib.fmi_norm_act_nhwc(_ptr(xn), c, _ptr(y), c, _ptr(ss), b, c, hw, float(slope), _lib.MMA_TF32, st),
                   "fmi_norm_act_nhwc")
        mean = sums[..., 0] / hw
        rstd = ((sums[..., 1] / hw - mean * mean).clamp_min_(0.0) + eps).rsqrt_()
        ctx.save_for_backward(xn, ss, torch.stack((mean, rstd), -1).float().contiguous())
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        g = _as_nhwc(gy, fresh=True)
        if not ctx.normed:
            xn, = ctx.saved_tensors
            return torch.ops.aten.leaky_relu_backward(g, xn, ctx.slope, False).permute(0, 3, 1, 2), None, None, None, None
        xn, ss, mr = ctx.saved_tensors
        b, h, w, c = xn.shape
        dx = torch.empty_like(xn)
        s2 = torch.empty((b, c, 2), dtype=torch.float64, device=g.device)
        _lib.check(_lib.load().fmi_instnorm_act_bwd_nhwc(_ptr(g), _ptr(xn), _ptr(ss), _ptr(mr), _ptr(dx), _ptr(s2), b, c, h * w,
                                                         ctx.slope, _stream()), "fmi_instnorm_act_bwd_nhwc")
        tot = s2.sum(0).float()
        return (dx.permute(0, 3, 1, 2), tot[:, 1].contiguous() if ctx.affine[0] else None,
                tot[:, 0].contiguous() if ctx.affine[1] else None, None, None)


class _RoundTF32(Function):
    """x rounded to TF32 (round to nearest), gradient passed through: for a convolution input that no kernel of this package
    produced (a block input read by the shortcut convolution): see `act_round`."""

    @staticmethod
    def forward(ctx, x):
        b, c, h, w = x.shape
        xn = _as_nhwc(x)
        y = torch.empty((b, h, w, c), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().fmi_norm_act_nhwc(_ptr(xn), c, _ptr(y), c, None, b, c, h * w, 1.0, _lib.MMA_TF32, _stream()),
                   "fmi_norm_act_nhwc")
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        return gy


def _mark_tf32(t):
    t._fmi_tf32 = True      # values are TF32-representable: a convolution reading it needs no rounding pass
    return t


def _train_kernels_on(x) -> bool:
    """The conv blocks' training path (forward + backward on the kernels) applies: CUDA fp32 batch under autograd, TF32
    convolutions allowed (the reference's own GPU numerics), single-pass TF32 operand mode. FMI_CONV_TRAIN=0 switches it off."""
    if not (torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and torch.is_grad_enabled()):
        return False
    if os.environ.get("FMI_CONV_TRAIN") == "0" or os.environ.get("FMI_PICNET_CUDNN") == "1":
        return False
    return torch.backends.cudnn.allow_tf32 and mma_mode(torch.float32) == _lib.MMA_TF32 and not tf32_split()


def norm_act_supported(norm, act, x) -> bool:
    if not (isinstance(norm, nn.InstanceNorm2d) and not norm.track_running_stats and isinstance(act, (nn.LeakyReLU, nn.ReLU))):
        return False
    return _train_kernels_on(x) and x.shape[1] % 4 == 0 and x.shape[1] <= 1024 and x.shape[1] == norm.num_features


def norm_act(norm, act, x):
    """act(norm(x)) for an (InstanceNorm2d, LeakyReLU / ReLU) pair through `_NormAct`."""
    slope = float(act.negative_slope) if isinstance(act, nn.LeakyReLU) else 0.0
    return _mark_tf32(_NormAct.apply(x, norm.weight, norm.bias, norm.eps, slope))


def act_round_supported(act, x) -> bool:
    return isinstance(act, (nn.LeakyReLU, nn.ReLU)) and not act.inplace and _train_kernels_on(x) and x.shape[1] % 4 == 0


def act_round(act, x):
    """act(x) with the result rounded to TF32 (round to nearest): the convolution that reads it would otherwise TRUNCATE the fp32
    values (kind::tf32 ignores the low 13 mantissa bits), a bias of -3e-4 per operand that does not average out over the
    pixel sum of a weight gradient."""
    slope = float(act.negative_slope) if isinstance(act, nn.LeakyReLU) else 0.0
    return _mark_tf32(_NormAct.apply(x, None, None, None, slope))


class _AvgPool2(Function):
    """nn.AvgPool2d(2, 2) with the gradient as a nearest-neighbour up-sampling (ATen's avg_pool2d_backward kernel measured 55 us
    per call on the encoder / discriminator maps, 2 ms per GAN step)."""

    @staticmethod
    def forward(ctx, x):
        return torch.nn.functional.avg_pool2d(x, 2, 2)

    @staticmethod
    def backward(ctx, g):
        return torch.nn.functional.interpolate(g, scale_factor=2, mode="nearest") * 0.25


def _is_pool2(m) -> bool:
    def two(v):
        return v in (2, (2, 2))
    return (isinstance(m, nn.AvgPool2d) and two(m.kernel_size) and two(m.stride) and m.padding in (0, (0, 0)) and not m.ceil_mode
            and m.divisor_override is None)


def avg_pool2(pool, x):
    if _is_pool2(pool) and _train_kernels_on(x) and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
        return _AvgPool2.apply(x)
    return pool(x)


def pool_sum(pool, a, b):
    """pool(a) + pool(b) (ResBlock with sample_type 'down', base_function.py:262-264); average pooling is linear, so under the
    training kernels the sum is pooled once."""
    if _is_pool2(pool) and _train_kernels_on(a) and a.shape == b.shape and a.shape[2] % 2 == 0 and a.shape[3] % 2 == 0:
        return _AvgPool2.apply(a + b)
    return pool(a) + pool(b)


class _ActReflectPad(Function):
    """ReflectionPad2d(1)(leaky_relu(x)) — the head of the Output block (base_function.py:387-393) — on channels_last tensors:
    the activation is written straight into the interior of the padded [N,H+2,W+2,C] buffer and fmi_reflect_border_nhwc fills the
    border; the gradient folds the border back with four thin slice-adds. ATen's reflection_pad2d wants NCHW: on the 32-channel
    1024^2 activation of the PICNet decoder its `.contiguous()` alone was a 2.4 ms strided copy per GAN step."""

    @staticmethod
    def forward(ctx, x, slope):
        b, c, h, w = x.shape
        xn = _as_nhwc(x)
        pad = torch.empty((b, h + 2, w + 2, c), dtype=torch.float32, device=x.device)
        torch.ops.aten.leaky_relu.out(xn, slope, out=pad[:, 1:-1, 1:-1, :])
        _lib.check(_lib.load().fmi_reflect_border_nhwc(_ptr(pad), b, c, h, w, _lib.MMA_TF32, _stream()), "fmi_reflect_border_nhwc")
        ctx.save_for_backward(xn)
        ctx.slope = float(slope)
        return pad.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, gy):
        xn, = ctx.saved_tensors
        g = _as_nhwc(gy, fresh=True)                      # [N, H+2, W+2, C]
        r = g[:, 1:-1].clone()                            # padded row 0 mirrors image row 1, padded row H+1 mirrors row H-2
        r[:, 1] += g[:, 0]
        r[:, -2] += g[:, -1]
        gx = r[:, :, 1:-1].contiguous()
        gx[:, :, 1] += r[:, :, 0]
        gx[:, :, -2] += r[:, :, -1]
        return torch.ops.aten.leaky_relu_backward(gx, xn, ctx.slope, False).permute(0, 3, 1, 2), None


def act_reflect_pad_supported(act, pad, x) -> bool:
    return (isinstance(act, (nn.LeakyReLU, nn.ReLU)) and isinstance(pad, nn.ReflectionPad2d) and tuple(pad.padding) == (1, 1, 1, 1)
            and _train_kernels_on(x) and x.shape[1] % 4 == 0 and x.shape[2] >= 2 and x.shape[3] >= 2)


def run_block_sequential(seq, x):
    """`seq(x)` for the `model` Sequential of a PICNet block (base_function.py:207-366) with every (InstanceNorm2d, activation)
    pair fused into `_NormAct` when the training kernels apply; the wrapped convolutions take their own kernel path in
    SpectralNorm.forward."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        if i + 1 < len(mods) and norm_act_supported(mods[i], mods[i + 1], x):
            x = norm_act(mods[i], mods[i + 1], x)
            i += 2
        elif i + 1 < len(mods) and act_reflect_pad_supported(mods[i], mods[i + 1], x):
            slope = float(mods[i].negative_slope) if isinstance(mods[i], nn.LeakyReLU) else 0.0
            x = _ActReflectPad.apply(x, slope)
            i += 2
        elif act_round_supported(mods[i], x):
            x = act_round(mods[i], x)
            i += 1
        elif isinstance(mods[i], nn.AvgPool2d):
            x = avg_pool2(mods[i], x)
            i += 1
        else:
            x = mods[i](x)
            i += 1
    return x


def _pow2(v):
    return v > 0 and (v & (v - 1)) == 0


def conv_train_supported(conv, x) -> bool:
    """True when `conv(x)` can run (forward and backward) through `_ConvShared` / `_ConvTShared`: a plain 3x3 / 1x1 stride-1
    nn.Conv2d or a ConvTranspose2d(3, 2, 1, 1) on a CUDA fp32 batch under autograd, TF32 convolutions allowed (with them switched
    off, or FMI_PRECISION=bf16, the reference formulation on cuDNN stays), channel counts and extents the GEMM kernels take."""
    if not _train_kernels_on(x):
        return False
    h, w = x.shape[2], x.shape[3]
    if isinstance(conv, nn.ConvTranspose2d):
        if (conv.kernel_size != (3, 3) or conv.stride != (2, 2) or conv.padding != (1, 1) or conv.output_padding != (1, 1)
                or conv.dilation != (1, 1) or conv.groups != 1):
            return False
    elif isinstance(conv, nn.Conv2d):
        k = conv.kernel_size[0]
        if (conv.kernel_size not in ((1, 1), (3, 3)) or conv.stride != (1, 1) or conv.padding != (k // 2, k // 2)
                or conv.dilation != (1, 1) or conv.groups != 1 or conv.padding_mode != "zeros"):
            return False
    else:
        return False
    i, o = conv.in_channels, conv.out_channels
    return (i == x.shape[1] and i % 32 == 0 and o % 32 == 0 and (i <= 256 or i % 256 == 0) and (o <= 128 or o % 128 == 0)
            and (o <= 256 or o % 256 == 0) and i <= 1024 and o <= 1024 and _pow2(h) and _pow2(w) and w >= 4 and h * w >= 16)


def conv_train(conv, x):
    """`conv.forward(x)` for a (SpectralNorm-wrapped) convolution whose `.weight` may be a plain tensor (external_function.py:57)."""
    fn = _ConvTShared if isinstance(conv, nn.ConvTranspose2d) else _ConvShared
    if not getattr(x, "_fmi_tf32", False) and x.shape[1] % 4 == 0 and os.environ.get("FMI_CONV_TRAIN_ROUND") != "0":
        x = _RoundTF32.apply(x)     # the MMA would truncate instead (a bias that survives the pixel sum of the weight gradient)
    return fn.apply(x, conv.weight, conv.bias)


# ------------------------------------------------------------------------------------------------
# f3: loss-side S x S / Gram products
# ------------------------------------------------------------------------------------------------
def _split3(t, order):
    """[B, R, K] fp32 -> [B, R, 3K]: x = hi + lo (hi = the TF32 part), laid out [hi | hi | lo] (order 0) or [hi | lo | hi]
    (order 1) along K, so that one TF32 GEMM over 3K sums hi*hi + hi*lo + lo*hi: fp32-class products (fmi_tf32_split3)."""
    b, r, k = t.shape
    out = torch.empty((b, r, 3 * k), dtype=torch.float32, device=t.device)
    _lib.check(_lib.load().fmi_tf32_split3(_ptr(t), k, _ptr(out), b * r, k, order, _stream()), "fmi_tf32_split3")
    return out


_CHUNK_VIEW_OK = None     # None: not tried yet; whether cuTensorMapEncodeTiled takes a batch stride below the row pitch


def _bmm_nt_raw(a, b):
    a, b = a.contiguous().float(), b.contiguous().float()
    bs, m, k = a.shape
    n = b.shape[1]
    lib, st = _lib.load(), _stream()
    a3, b3 = _split3(a, 0), _split3(b, 1)
    k3 = 3 * k
    # Gram matrices (C x C outputs over H*W up to 65536 pixels) have few output tiles and a very long contraction: split K over
    # `s` chunks that run as extra batch entries of the one launch (operands re-laid [image, chunk, row, k]), partial products
    # summed afterwards — without it one CTA per image walks the whole contraction (measured 1 ms per 64 x 64 Gram)
    tiles = bs * ((m + 127) // 128) * ((n + 127) // 128)
    s = 1
    while tiles * s < 148 and k3 % (2 * s * 32) == 0 and k3 // (2 * s) >= 1024:
        s *= 2
    if s == 1:
        c = torch.empty((bs, m, n), dtype=torch.float32, device=a.device)
        _lib.check(lib.fmi_gemm_nt(_ptr(a3), k3, m * k3, _ptr(b3), k3, n * k3, _ptr(c), n, m * n, bs, m, n, k3, 0, _lib.MMA_TF32, st),
                   "fmi_gemm_nt")
        return c
    part = torch.empty((bs, s, m, n), dtype=torch.float32, device=a.device)
    kc = k3 // s
    global _CHUNK_VIEW_OK
    if _CHUNK_VIEW_OK is not False:
        # chunks as batch entries of a strided VIEW: row pitch k3, batch stride kc (a tensor map whose outer stride is smaller
        # than its row pitch); one launch per image, no re-layout of the operands
        rc = 0
        for i in range(bs):
            rc = lib.fmi_gemm_nt(a3[i].data_ptr(), k3, kc, b3[i].data_ptr(), k3, kc, part[i].data_ptr(), n, m * n, s, m, n, kc, 0,
                                 _lib.MMA_TF32, st)
            if rc:
                break
        if rc == 0:
            _CHUNK_VIEW_OK = True
            return part.sum(1)
        if _CHUNK_VIEW_OK:      # worked before: a real error
            _lib.check(rc, "fmi_gemm_nt")
        _CHUNK_VIEW_OK = False  # the driver refused that tensor map: re-lay the operands instead
    ac = a3.view(bs, m, s, kc).permute(0, 2, 1, 3).contiguous()      # [image, chunk, row, kc]: chunks as batch entries
    bc = b3.view(bs, n, s, kc).permute(0, 2, 1, 3).contiguous()
    _lib.check(lib.fmi_gemm_nt(_ptr(ac), kc, m * kc, _ptr(bc), kc, n * kc, _ptr(part), n, m * n, bs * s, m, n, kc, 0, _lib.MMA_TF32, st),
               "fmi_gemm_nt")
    return part.sum(1)


class _BmmNT(Function):
    """C[b] = A[b] @ B[b]^T for [B,M,K] x [B,N,K] fp32 (K contiguous) on the tcgen05 GEMM with error-compensated 3xTF32 operands;
    backward dA = G @ B, dB = G^T @ A through the same kernel (on transposed copies)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _bmm_nt_raw(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _bmm_nt_raw(g, b.transpose(1, 2))          # [B,M,N] x [B,K,N]^T
        if ctx.needs_input_grad[1]:
            db = _bmm_nt_raw(g.transpose(1, 2), a.transpose(1, 2))
        return da, db


def bmm_nt_supported(a, b) -> bool:
    return (torch.is_tensor(a) and a.is_cuda and b.is_cuda and a.dim() == 3 and b.dim() == 3 and a.dtype == torch.float32
            and b.dtype == torch.float32 and a.shape[0] == b.shape[0] and a.shape[2] == b.shape[2] and a.shape[2] % 8 == 0
            and a.shape[1] % 8 == 0 and b.shape[1] % 8 == 0 and os.environ.get("FMI_LOSS_KERNELS") != "0")


def bmm_nt(a, b):
    """torch.bmm(a, b.transpose(1, 2)) (the reference's loss-side products) on the tensor cores at fp32-class accuracy."""
    _need_cuda(a, b)
    return _BmmNT.apply(a, b)


def gram_matrix(input):
    """GramMatrix (modules/pluralistic_model/external_function.py:180-185): features @ features^T / (C H W) per image."""
    s = input.size()
    f = input.reshape(s[0], s[1], s[2] * s[3])
    g = bmm_nt(f, f) if bmm_nt_supported(f, f) else torch.bmm(f, f.transpose(1, 2))
    return g.div(s[1] * s[2] * s[3])


def contextual_loss(x, y, h=0.5):
    """contextual_loss (external_function.py:231-274), same arithmetic; the S x S cosine similarities — torch.bmm there, an fp32
    SIMT GEMM — on the tcgen05 GEMM (3xTF32 operands: d / (d_min + 1e-5) amplifies any error of the similarities)."""
    assert x.size() == y.size()
    n, c, hh, ww = x.size()
    y_mu = y.mean(3).mean(2).mean(0).reshape(1, -1, 1, 1)
    x_centered, y_centered = x - y_mu, y - y_mu
    x_normalized = x_centered / torch.norm(x_centered, p=2, dim=1, keepdim=True)
    y_normalized = y_centered / torch.norm(y_centered, p=2, dim=1, keepdim=True)
    xt = x_normalized.reshape(n, c, -1).transpose(1, 2)       # [N, S, C]
    yt = y_normalized.reshape(n, c, -1).transpose(1, 2)
    if bmm_nt_supported(xt, yt):
        cosine_sim = bmm_nt(xt.contiguous(), yt.contiguous())
    else:
        cosine_sim = torch.bmm(xt, yt.transpose(1, 2))
    d = 1 - cosine_sim
    d_min, _ = torch.min(d, dim=2, keepdim=True)
    d_tilde = d / (d_min + 1e-5)
    w = torch.exp((1 - d_tilde) / h)
    cx_ij = w / torch.sum(w, dim=2, keepdim=True)
    cx = torch.mean(torch.max(cx_ij, dim=1)[0], dim=1)
    return torch.mean(-torch.log(cx + 1e-5))


_PAD_BETA = 4.0   # exactly representable in bf16: the marker channel adds no rounding error to the logits


def _needs_padding(s, c0, c1):
    cv = c0 + c1
    return s % 128 != 0 or c0 % 32 != 0 or c1 % 32 != 0 or c0 < 32 or (cv > 256 and cv % 256 != 0)


def _pad_attention_args(x, wq, bq, v0, v1, mask):
    """Shapes the kernels do not take directly — H*W not a multiple of the 128-row tile (CelebA 218x178 inputs give 6x5 and
    24x20 feature maps), value channels not a multiple of 32 — are embedded in the next shape they do take; the reference
    accepts any size (example_guided_att.py:21-41, base_function.py:420-448). Pixels: zero-padded to S' = ceil(S / 128) * 128.
    A padded KEY must get softmax weight 0: two marker channels are appended to x — t1 = 1 on real pixels / 0 on padding
    (carries the query bias: column t1 of the augmented query weight is bq), t2 = +1 / -1 (row d of the augmented weight is
    beta * t2) — so q' = [Wq x + bq ; beta] on real pixels and [0 ; -beta] on padding: real-real logits gain the constant
    beta^2 (softmax-invariant), real-padding logits are -beta^2, i.e. at least e^(-2 beta^2) = e^-32 below the diagonal.
    Padded QUERY rows and padded value channels produce rows / channels that are sliced off. Returns the padded arguments
    and (S, S', c0', c1')."""
    n, c = x.shape[0], x.shape[1]
    s = x[0, 0].numel()
    sp = (s + 127) // 128 * 128
    d = wq.shape[0]
    c0 = v0.shape[1]
    c1 = v1.shape[1] if v1 is not None else 0
    up32 = lambda k: max(32, (k + 31) // 32 * 32)
    c0p, c1p = up32(c0), (up32(c1) if c1 else 0)
    if c0p + c1p > 256 and (c0p + c1p) % 256:
        extra = 256 - (c0p + c1p) % 256
        if c1:
            c1p += extra
        else:
            c0p += extra
    dev, dt = x.device, x.dtype
    xa = torch.zeros((n, c + 2, sp, 1), dtype=dt, device=dev)
    xa[:, :c, :s, 0] = x.reshape(n, c, s)
    xa[:, c, :s] = 1.0
    xa[:, c + 1] = -1.0
    xa[:, c + 1, :s] = 1.0
    wa = torch.zeros((d + 1, c + 2), dtype=torch.float32, device=dev)
    wa[:d, :c] = wq.reshape(d, c).float()
    if bq is not None:
        wa[:d, c] = bq.float()
    wa[d, c + 1] = _PAD_BETA

    def padv(v, cp):
        if v is None:
            return None
        out = torch.zeros((n, cp, sp, 1), dtype=dt, device=dev)
        out[:, :v.shape[1], :s, 0] = v.reshape(n, v.shape[1], s).to(dt)
        return out

    ma = None
    if mask is not None:
        ma = torch.zeros((n, 1, sp, 1), dtype=torch.float32, device=dev)
        ma[:, 0, :s, 0] = mask.reshape(n, s).float()
    return xa, wa, padv(v0, c0p), padv(v1, c1p), ma, (s, sp, c0, c1, c0p, c1p)


def attention_forward(x, wq, bq, v0, v1=None, mask=None, a0=None, b0=0.0, masked0=False, a1=None, b1=0.0,
                      masked1=False, order=(0, 1), need_lse=False, mma=None, save_o=False):
    """One fused pass of fmi_attn_fwd. x [N,C,H,W]; v0/v1 value groups [N,C0|C1,H,W]; mask [N,1,H,W] or None.
    Returns (out [N, C0+C1, H, W] with the groups concatenated in `order`, lse [N,S] or None, workspace)."""
    _need_cuda(x, wq, bq, v0, v1, mask, a0, a1)
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("fmi_b200 attention: fp32 or bf16 activations only")
    if _needs_padding(x[0, 0].numel(), v0.shape[1], v1.shape[1] if v1 is not None else 0):
        xa, wa, v0a, v1a, ma, (s, sp, c0, c1, c0p, c1p) = _pad_attention_args(x, wq, bq, v0, v1, mask)
        res = attention_forward(xa, wa, None, v0a, v1a, mask=ma, a0=a0, b0=b0, masked0=masked0, a1=a1, b1=b1, masked1=masked1,
                                order=(0, 1), need_lse=need_lse, mma=mma if mma is not None else mma_mode(x.dtype), save_o=save_o)
        g0, g1 = res[0][:, :c0, :s, 0], res[0][:, c0p:c0p + c1, :s, 0]
        out = torch.cat([g0, g1] if order[0] == 0 else [g1, g0], dim=1).reshape((x.shape[0], c0 + c1) + tuple(x.shape[2:]))
        # lse / o_saved stay in the padded layout: they are only handed back to attention_backward / attention_map
        return (out.contiguous(),) + tuple(res[1:])
    xc = x.contiguous()
    v0c = xc if v0 is x else v0.contiguous().to(xc.dtype)
    v1c = v1.contiguous().to(xc.dtype) if v1 is not None else None
    n, c = xc.shape[0], xc.shape[1]
    spatial = tuple(xc.shape[2:])
    s = xc[0, 0].numel()
    d = wq.shape[0]
    c0 = v0c.shape[1]
    c1 = v1c.shape[1] if v1c is not None else 0
    w = wq.reshape(d, c).contiguous().float()
    b = bq.contiguous().float() if bq is not None else None
    m = mask.reshape(n, s).contiguous().float() if mask is not None else None
    if mma is None:
        mma = mma_mode(xc.dtype)
    lib = _lib.load()
    ws_bytes = lib.fmi_attn_workspace_bytes(n, c, d, c0, c1, s, mma)
    if ws_bytes < 0:
        raise RuntimeError(f"fmi_attn_workspace_bytes: {_lib.last_error()}")
    ws = _workspace(xc.device, ws_bytes)
    out = torch.empty((n, c0 + c1) + spatial, dtype=xc.dtype, device=xc.device)
    esz = out.element_size()
    bs = (c0 + c1) * s
    off0 = 0 if order[0] == 0 else c1 * s
    off1 = c0 * s if order[0] == 0 else 0
    lse = torch.empty((n, s), dtype=torch.float32, device=xc.device) if need_lse else None
    o_saved = torch.empty((n, c0 + c1, s), dtype=xc.dtype, device=xc.device) if save_o else None
    a0c = a0.reshape(1).float().contiguous() if a0 is not None else None
    a1c = a1.reshape(1).float().contiguous() if a1 is not None else None
    _lib.check(lib.fmi_attn_fwd(_ptr(xc), _ptr(w), _ptr(b), _ptr(v0c), _ptr(v1c), _ptr(m), _ptr(a0c), float(b0),
                                int(masked0), _ptr(a1c), float(b1), int(masked1), out.data_ptr() + off0 * esz, bs,
                                (out.data_ptr() + off1 * esz) if c1 else None, bs, _ptr(lse), _ptr(o_saved), n, c, d, c0,
                                c1, s,
                                _dt(xc), mma, ws.data_ptr(), ws.numel(), _stream()), "fmi_attn_fwd")
    if save_o:
        return out, lse, ws, o_saved
    return out, lse, ws


def attention_map(ws, lse, n, d, s, mma):
    """Opt-in S x S map (what Auto_Attn returns, base_function.py:448) from the staged q and the row lse."""
    if lse.shape[1] != s:      # padded call (_pad_attention_args): one marker channel more, S' rows; padded keys have weight 0
        sp = lse.shape[1]
        return attention_map(ws, lse, n, d + 1, sp, mma)[:, :s, :s].contiguous()
    attn = torch.empty((n, s, s), dtype=torch.float32, device=lse.device)
    _lib.check(_lib.load().fmi_attn_materialize(ws.data_ptr(), _ptr(lse), _ptr(attn), n, d, s, mma, _stream()),
               "fmi_attn_materialize")
    return attn


def attention_backward(x, wq, bq, v0, v1, mask, a0, b0, masked0, a1, b1, masked1, o_saved, lse, grad_out, order=(0, 1),
                       need_dv0=True, need_dv1=True, mma=None):
    """fmi_attn_bwd for the forward call with the same arguments; grad_out is [N, C0+C1, H, W] (groups concatenated in
    `order`). Returns (dq [N,d,H,W] fp32, dv0, dv1 (fp32 or None), da0, da1 (fp32 scalars))."""
    _need_cuda(x, wq, v0, grad_out)
    if _needs_padding(x[0, 0].numel(), v0.shape[1], v1.shape[1] if v1 is not None else 0):
        xa, wa, v0a, v1a, ma, (s, sp, c0, c1, c0p, c1p) = _pad_attention_args(x, wq, bq, v0, v1, mask)
        n, d = x.shape[0], wq.shape[0]
        spatial = tuple(x.shape[2:])
        g = grad_out.reshape(n, c0 + c1, s)
        g0, g1 = (g[:, :c0], g[:, c0:]) if order[0] == 0 else (g[:, c1:], g[:, :c1])
        ga = torch.zeros((n, c0p + c1p, sp, 1), dtype=x.dtype, device=x.device)
        ga[:, :c0, :s, 0] = g0
        if c1:
            ga[:, c0p:c0p + c1, :s, 0] = g1
        dq, dv0, dv1, da0, da1 = attention_backward(xa, wa, None, v0a, v1a, ma, a0, b0, masked0, a1, b1, masked1, o_saved, lse, ga,
                                                    order=(0, 1), need_dv0=need_dv0, need_dv1=need_dv1,
                                                    mma=mma if mma is not None else mma_mode(x.dtype))
        # dq of the marker row and of padded pixels is dropped: the query conv's own backward (qconv_backward) sees the
        # real q gradient, from which dWq, dbq and dx follow as in the unpadded case
        dq = dq[:, :d, :s, 0].reshape((n, d) + spatial).contiguous()
        dv0 = dv0[:, :c0, :s, 0].reshape((n, c0) + spatial).contiguous() if dv0 is not None else None
        dv1 = dv1[:, :c1, :s, 0].reshape((n, c1) + spatial).contiguous() if dv1 is not None else None
        return dq, dv0, dv1, da0, da1
    xc = x.contiguous()
    v0c = xc if v0 is x else v0.contiguous().to(xc.dtype)
    v1c = v1.contiguous().to(xc.dtype) if v1 is not None else None
    n, c = xc.shape[0], xc.shape[1]
    spatial = tuple(xc.shape[2:])
    s = xc[0, 0].numel()
    d = wq.shape[0]
    c0 = v0c.shape[1]
    c1 = v1c.shape[1] if v1c is not None else 0
    w = wq.reshape(d, c).contiguous().float()
    b = bq.contiguous().float() if bq is not None else None
    m = mask.reshape(n, s).contiguous().float() if mask is not None else None
    if mma is None:
        mma = mma_mode(xc.dtype)
    g = grad_out.contiguous().to(xc.dtype)
    esz = g.element_size()
    bs = (c0 + c1) * s
    off0 = 0 if order[0] == 0 else c1 * s
    off1 = c0 * s if order[0] == 0 else 0
    lib = _lib.load()
    ws_bytes = lib.fmi_attn_bwd_workspace_bytes(n, c, d, c0, c1, s, mma)
    if ws_bytes < 0:
        raise RuntimeError(f"fmi_attn_bwd_workspace_bytes: {_lib.last_error()}")
    ws = _workspace(xc.device, ws_bytes)
    dpad = (d + 63) // 64 * 64
    dq = torch.empty((n, s, dpad), dtype=torch.float32, device=xc.device)
    dv0 = torch.empty((n, c0) + spatial, dtype=torch.float32, device=xc.device) if need_dv0 else None
    dv1 = torch.empty((n, c1) + spatial, dtype=torch.float32, device=xc.device) if (need_dv1 and c1) else None
    da = torch.zeros(2, dtype=torch.float32, device=xc.device)
    a0c = a0.reshape(1).float().contiguous() if a0 is not None else None
    a1c = a1.reshape(1).float().contiguous() if a1 is not None else None
    _lib.check(lib.fmi_attn_bwd(_ptr(xc), _ptr(w), _ptr(b), _ptr(v0c), _ptr(v1c), _ptr(m), _ptr(a0c), float(b0),
                                int(masked0), _ptr(a1c), float(b1), int(masked1), _ptr(o_saved), _ptr(lse),
                                g.data_ptr() + off0 * esz, bs, (g.data_ptr() + off1 * esz) if c1 else None, bs,
                                _ptr(dq), _ptr(dv0), _ptr(dv1), da.data_ptr(), da.data_ptr() + 4, n, c, d, c0, c1, s,
                                _dt(xc), mma, ws.data_ptr(), ws.numel(), _stream()), "fmi_attn_bwd")
    dq = dq[:, :, :d].permute(0, 2, 1).reshape((n, d) + spatial)
    return dq, dv0, dv1, da[0], da[1]


def qconv_backward(dq, x, weight, need_bias):
    """Gradients of the 1x1 query conv from dq [N,d,H,W]: plain dense products (library GEMMs, not hot-path kernels):
    dW[d,C] = sum_n dq_n x_n^T, dx = W^T dq, db = sum dq."""
    n, d = dq.shape[0], dq.shape[1]
    c = x.shape[1]
    dqf = dq.reshape(n, d, -1)
    xf = x.reshape(n, c, -1).float()
    dw = torch.einsum('nds,ncs->dc', dqf, xf).reshape(weight.shape)
    dx = torch.einsum('dc,nds->ncs', weight.reshape(d, c).float(), dqf).reshape(x.shape)
    db = dqf.sum((0, 2)) if need_bias else None
    return dw, dx, db
