"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain PyTorch (CPU, fp32 or fp64) restatement of the reference's algorithm for the generator hot
path of syncdoth/face_mask_inpaint. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this package, and only as the checker / the timed
CPU baseline. The product (`face_mask_inpaint_b200`) never imports it and has no CPU fallback.

Why torch and not numpy/C: the reference's arithmetic for this path lives in PyTorch ATen
(`torch.bmm`, `torch.softmax`, `F.conv2d`, `F.conv_transpose2d`, `F.interpolate`; unpinned dependency,
scripts/env_setup.sh:32 — container has torch 2.11.0) and in two in-tree CUDA kernels that cannot run
on a CPU (modules/psp/stylegan2/op/*.cu need ATen + a GPU: unbuildable here). Each function
below cites the reference file:line (relative to the reference root) that it restates.

Pinning: the reference ships NO tests / golden vectors / KATs. This oracle is pinned
instead against outputs of the reference's own modules imported from /root/reference in the build
container: `tests/golden/make_golden.py` generated `tests/golden/*.npz`, and
`tests/test_oracle_golden.py` checks every function here against them. `fused_bias_act` and `upfirdn2d`,
whose reference implementations are CUDA kernels, are additionally pinned on the GPU against the reference's
own compiled ops (oracle/build_ref.py -> oracle/_ref/*.so, tests/test_reference_ops_gpu.py).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.nn.functional as F

SQRT2 = 2 ** 0.5


# --------------------------------------------------------------------------------------------
# a5  fused bias + leaky relu
# --------------------------------------------------------------------------------------------
def fused_bias_act(x: torch.Tensor, b: Optional[torch.Tensor], ref: Optional[torch.Tensor], act: int, grad: int,
                   alpha: float, scale: float) -> torch.Tensor:
    """modules/psp/stylegan2/op/fused_bias_act_kernel.cu:18-49 — y = act'(x + b[(i/step_b)%size_b]) * scale."""
    if b is not None and b.numel():
        shape = [1, -1] + [1] * (x.ndim - 2)
        x = x + b.view(*shape).to(x.dtype)
    code = act * 10 + grad
    if code == 30:
        y = torch.where(x > 0, x, x * alpha)
    elif code == 31:
        y = torch.where(ref > 0, x, x * alpha)
    elif code in (12, 32):
        y = torch.zeros_like(x)
    else:
        y = x
    return y * scale


def fused_leaky_relu(x: torch.Tensor, bias: torch.Tensor, negative_slope: float = 0.2, scale: float = SQRT2):
    """modules/psp/stylegan2/op/fused_act.py:84-85,52-59 (act=3, grad=0)."""
    return fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)


def fused_leaky_relu_backward(grad_out: torch.Tensor, out: torch.Tensor, negative_slope: float = 0.2,
                              scale: float = SQRT2):
    """modules/psp/stylegan2/op/fused_act.py:18-38 — (grad_input, grad_bias)."""
    gi = fused_bias_act(grad_out, None, out, 3, 1, negative_slope, scale)
    dims = [0] + list(range(2, gi.ndim))
    return gi, gi.sum(dims)


# --------------------------------------------------------------------------------------------
# a4  upfirdn2d
# --------------------------------------------------------------------------------------------
def upfirdn2d_native(inp: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
                     pad_x0: int, pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    """modules/psp/stylegan2/op/upfirdn2d.py:150-184 (the reference's own CPU restatement of
    upfirdn2d_kernel.cu:52-137; as shipped it forgets to import F). inp is [major, in_h, in_w, minor]."""
    _, in_h, in_w, minor = inp.shape
    kernel_h, kernel_w = kernel.shape
    out = inp.reshape(-1, in_h, 1, in_w, 1, minor)
    out = F.pad(out, [0, 0, 0, up_x - 1, 0, 0, 0, up_y - 1])
    out = out.reshape(-1, in_h * up_y, in_w * up_x, minor)
    out = F.pad(out, [0, 0, max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    out = out[:, max(-pad_y0, 0):out.shape[1] - max(-pad_y1, 0), max(-pad_x0, 0):out.shape[2] - max(-pad_x1, 0), :]
    out = out.permute(0, 3, 1, 2)
    out = out.reshape([-1, 1, in_h * up_y + pad_y0 + pad_y1, in_w * up_x + pad_x0 + pad_x1])
    w = torch.flip(kernel, [0, 1]).view(1, 1, kernel_h, kernel_w).to(out)
    out = F.conv2d(out, w)
    out = out.reshape(-1, minor, in_h * up_y + pad_y0 + pad_y1 - kernel_h + 1,
                      in_w * up_x + pad_x0 + pad_x1 - kernel_w + 1)
    out = out.permute(0, 2, 3, 1)
    return out[:, ::down_y, ::down_x, :]


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up: int = 1, down: int = 1, pad: Sequence[int] = (0, 0)):
    """modules/psp/stylegan2/op/upfirdn2d.py:142-147,87-121 — NCHW wrapper (minor = 1)."""
    n, c, h, w = x.shape
    out = upfirdn2d_native(x.reshape(-1, h, w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return out.reshape(n, c, out.shape[1], out.shape[2])


def upfirdn2d_backward(grad_out: torch.Tensor, kernel: torch.Tensor, up: int, down: int, pad: Sequence[int],
                       in_size: Sequence[int]):
    """modules/psp/stylegan2/op/upfirdn2d.py:17-57,108-113 — grad wrt input: same op with up<->down swapped,
    flipped kernel and g_pad."""
    n, c, in_h, in_w = in_size
    kh, kw = kernel.shape
    out_h, out_w = grad_out.shape[-2:]
    g_pad_x0 = kw - pad[0] - 1
    g_pad_y0 = kh - pad[0] - 1
    g_pad_x1 = in_w * up - out_w * down + pad[0] - up + 1
    g_pad_y1 = in_h * up - out_h * down + pad[0] - up + 1
    g = upfirdn2d_native(grad_out.reshape(-1, out_h, out_w, 1), torch.flip(kernel, [0, 1]), down, down, up, up,
                         g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
    return g.reshape(n, c, in_h, in_w)


def make_kernel(k) -> torch.Tensor:
    """modules/psp/stylegan2/model.py:19-27."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    k = k / k.sum()
    return k


# --------------------------------------------------------------------------------------------
# a7  compositing
# --------------------------------------------------------------------------------------------
def scale_img(img: torch.Tensor, size) -> torch.Tensor:
    """modules/model.py:10-12."""
    return F.interpolate(img, size=size, mode='bilinear', align_corners=True)


def composite(src: torch.Tensor, ref: torch.Tensor, mask_full: torch.Tensor) -> torch.Tensor:
    """modules/model.py:98-99 — (1 - m) * src + m * ref with m = scale_img(mask[N,1,Hm,Wm], feature size).
    (psp_encoders.py:135-138 writes the same blend as m * ref + (1 - m) * src.)"""
    m = scale_img(mask_full.to(torch.float32), src.shape[-2:]).to(src.dtype)
    return (1 - m) * src + m * ref


# --------------------------------------------------------------------------------------------
# a1  ExampleGuidedAttention
# --------------------------------------------------------------------------------------------
def example_guided_attention(src_mask: torch.Tensor, src_feature: torch.Tensor, ref_feature: torch.Tensor,
                             conv_weight: torch.Tensor, out_conv_weight: Optional[torch.Tensor] = None,
                             out_conv_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """modules/example_guided_att.py:21-41 (apply_attention_map :15-19)."""
    n, c, h, w = src_feature.shape
    query = F.conv2d(src_feature, conv_weight)                       # :27
    query = query.reshape(n, query.shape[1], -1)                     # :28
    att_map = torch.softmax(query.permute(0, 2, 1) @ query, dim=-1)  # :30  (no 1/sqrt(d))
    src_att = (src_feature.reshape(n, c, -1) @ att_map.permute(0, 2, 1)).reshape(n, c, h, w)  # :31
    ref_att = (ref_feature.reshape(n, c, -1) @ att_map.permute(0, 2, 1)).reshape(n, c, h, w)  # :32
    flow = (1 - src_mask) * ref_att + src_mask * ref_feature         # :34
    out = torch.cat([flow, src_att], dim=1)                          # :36
    if out_conv_weight is not None:
        out = F.conv2d(out, out_conv_weight, out_conv_bias)          # :38-39
    return out


# --------------------------------------------------------------------------------------------
# a2  Auto_Attn
# --------------------------------------------------------------------------------------------
def auto_attn(x: torch.Tensor, query_weight: torch.Tensor, query_bias: torch.Tensor, gamma: torch.Tensor,
              pre: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
              alpha: Optional[torch.Tensor] = None, return_attention: bool = False):
    """modules/pluralistic_model/base_function.py:420-448, up to (not including) the ResBlock at :446.
    Returns (out, context_flow or None, attention or None)."""
    b, c, w_, h_ = x.shape
    q = F.conv2d(x, query_weight, query_bias).view(b, -1, w_ * h_)   # :429
    energy = torch.bmm(q.permute(0, 2, 1), q)                        # :432
    attention = torch.softmax(energy, dim=-1)                        # :433
    value = x.view(b, -1, w_ * h_)                                   # :434
    out = torch.bmm(value, attention.permute(0, 2, 1)).view(b, c, w_, h_)  # :436-437
    out = gamma * out + x                                            # :439
    ctx = None
    if pre is not None:
        ctx = torch.bmm(pre.view(b, -1, w_ * h_), attention.permute(0, 2, 1)).view(b, -1, w_, h_)  # :443-444
        ctx = alpha * (1 - mask) * ctx + mask * pre                  # :445
    return out, ctx, (attention if return_attention else None)


# --------------------------------------------------------------------------------------------
# a3  ModulatedConv2d, a6 StyledConv / ToRGB / NoiseInjection / Generator synthesis loop
# --------------------------------------------------------------------------------------------
def equal_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], lr_mul: float = 1.0,
                 activation: bool = False) -> torch.Tensor:
    """modules/psp/stylegan2/model.py:159-167 (EqualLinear.forward)."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        out = F.linear(x, weight * scale)
        return fused_leaky_relu(out, bias * lr_mul)
    return F.linear(x, weight * scale, bias=bias * lr_mul if bias is not None else None)


def modulated_conv2d(x: torch.Tensor, style: torch.Tensor, weight: torch.Tensor, mod_weight: torch.Tensor,
                     mod_bias: torch.Tensor, demodulate: bool = True, upsample: bool = False,
                     blur_kernel: Sequence[int] = (1, 3, 3, 1), downsample: bool = False) -> torch.Tensor:
    """modules/psp/stylegan2/model.py:241-279 (the downsample branch :265-272 is not reachable from the repo's scripts).
    weight is the parameter [1, O, I, k, k]; style is the latent [B, style_dim]."""
    batch, in_channel, height, width = x.shape
    _, out_channel, _, ksize, _ = weight.shape
    scale = 1 / math.sqrt(in_channel * ksize ** 2)                                      # :225-226
    s = equal_linear(style, mod_weight, mod_bias).view(batch, 1, in_channel, 1, 1)      # :244
    w = scale * weight * s                                                              # :245
    if demodulate:
        demod = torch.rsqrt(w.pow(2).sum([2, 3, 4]) + 1e-8)                             # :248
        w = w * demod.view(batch, out_channel, 1, 1, 1)                                 # :249
    w = w.view(batch * out_channel, in_channel, ksize, ksize)                           # :251
    if upsample:
        xin = x.reshape(1, batch * in_channel, height, width)                           # :255
        w = w.view(batch, out_channel, in_channel, ksize, ksize).transpose(1, 2).reshape(
            batch * in_channel, out_channel, ksize, ksize)                              # :256-259
        out = F.conv_transpose2d(xin, w, padding=0, stride=2, groups=batch)             # :260
        out = out.view(batch, out_channel, out.shape[-2], out.shape[-1])                # :261-262
        factor = 2
        p = (len(blur_kernel) - factor) - (ksize - 1)                                   # :207-209
        pad0, pad1 = (p + 1) // 2 + factor - 1, p // 2 + 1
        k = make_kernel(list(blur_kernel)) * (factor ** 2)                              # :78-82
        out = upfirdn2d(out, k.to(out), pad=(pad0, pad1))                         # :263
    elif downsample:
        factor = 2
        p = (len(blur_kernel) - factor) + (ksize - 1)                                   # :217-221
        xb = upfirdn2d(x, make_kernel(list(blur_kernel)).to(x), pad=((p + 1) // 2, p // 2))   # :266
        xin = xb.reshape(1, batch * in_channel, xb.shape[-2], xb.shape[-1])             # :267-268
        out = F.conv2d(xin, w, padding=0, stride=2, groups=batch)                       # :269
        out = out.view(batch, out_channel, out.shape[-2], out.shape[-1])                # :270-271
    else:
        xin = x.reshape(1, batch * in_channel, height, width)                           # :274
        out = F.conv2d(xin, w, padding=ksize // 2, groups=batch)                        # :275
        out = out.view(batch, out_channel, out.shape[-2], out.shape[-1])                # :276-277
    return out


def styled_conv(x, style, weight, mod_weight, mod_bias, noise_weight, act_bias, noise, upsample=False,
                blur_kernel=(1, 3, 3, 1)):
    """modules/psp/stylegan2/model.py:340-346 with NoiseInjection :289-294 (noise given explicitly) and
    FusedLeakyReLU op/fused_act.py:80-85."""
    out = modulated_conv2d(x, style, weight, mod_weight, mod_bias, True, upsample, blur_kernel)
    out = out + noise_weight * noise
    return fused_leaky_relu(out, act_bias)


def to_rgb(x, style, weight, mod_weight, mod_bias, bias, skip=None, blur_kernel=(1, 3, 3, 1)):
    """modules/psp/stylegan2/model.py:360-369; Upsample :30-49 (kernel*4, pad (2,1))."""
    out = modulated_conv2d(x, style, weight, mod_weight, mod_bias, demodulate=False)
    out = out + bias
    if skip is not None:
        k = make_kernel(list(blur_kernel)) * 4
        skip = upfirdn2d(skip, k.to(skip), up=2, down=1, pad=(2, 1))
        out = out + skip
    return out


def _sc_args(sd, prefix):
    return (sd[f'{prefix}.conv.weight'], sd[f'{prefix}.conv.modulation.weight'], sd[f'{prefix}.conv.modulation.bias'],
            sd[f'{prefix}.noise.weight'], sd[f'{prefix}.activate.bias'])


def _rgb_args(sd, prefix):
    return (sd[f'{prefix}.conv.weight'], sd[f'{prefix}.conv.modulation.weight'], sd[f'{prefix}.conv.modulation.bias'],
            sd[f'{prefix}.bias'])


def generator_synthesis(sd: dict, latent: torch.Tensor, noises: Optional[Sequence[torch.Tensor]] = None):
    """modules/psp/stylegan2/model.py:528-543 — Generator.forward with input_is_latent=True, one [B, n_latent, 512]
    latent and noise taken from the registered buffers (randomize_noise=False, :498-500) unless given.
    `sd` is the Generator state_dict (parameter names of :372-447)."""
    batch = latent.shape[0]
    n_latent = latent.shape[1]
    num_layers = n_latent - 1
    if noises is None:
        noises = [sd[f'noises.noise_{i}'] for i in range(num_layers)]
    out = sd['input.input'].repeat(batch, 1, 1, 1)                                          # :528, :304-306
    out = styled_conv(out, latent[:, 0], *_sc_args(sd, 'conv1'), noise=noises[0])           # :529
    skip = to_rgb(out, latent[:, 1], *_rgb_args(sd, 'to_rgb1'))                             # :531
    i = 1
    n_blocks = (num_layers - 1) // 2
    for blk in range(n_blocks):                                                             # :534-541
        out = styled_conv(out, latent[:, i], *_sc_args(sd, f'convs.{2 * blk}'), noise=noises[1 + 2 * blk],
                          upsample=True)
        out = styled_conv(out, latent[:, i + 1], *_sc_args(sd, f'convs.{2 * blk + 1}'), noise=noises[2 + 2 * blk])
        skip = to_rgb(out, latent[:, i + 2], *_rgb_args(sd, f'to_rgbs.{blk}'), skip=skip)
        i += 2
    return skip


# --------------------------------------------------------------------------------------------
# f1  PICNet decoder conv blocks
# --------------------------------------------------------------------------------------------
def spectral_norm_weight(w_bar: torch.Tensor, u: torch.Tensor, v: torch.Tensor, power_iterations: int = 1):
    """SpectralNorm._update_u_v (modules/pluralistic_model/external_function.py:44-57): one power iteration, then
    w = w_bar / sigma. Returns (w, u', v'); the reference stores u', v' back into the module."""
    height = w_bar.shape[0]
    wm = w_bar.reshape(height, -1)
    for _ in range(power_iterations):
        v = torch.mv(wm.t(), u)
        v = v / (v.norm() + 1e-12)                     # l2normalize, :11-12
        u = torch.mv(wm, v)
        u = u / (u.norm() + 1e-12)
    sigma = u.dot(wm.mv(v))
    return w_bar / sigma, u, v


def res_block_decoder(x, conv1_w, conv1_b, conv2_w, conv2_b, bypass_w, bypass_b, norm1=None, norm2=None, slope=0.1):
    """ResBlockDecoder.forward (base_function.py:308-366) given the EFFECTIVE conv weights (after SpectralNorm):
    [IN] lrelu conv3x3 [IN] lrelu convT3x3(s2,p1,op1)  +  convT3x3(s2,p1,op1) shortcut. norm1 / norm2 = (gamma, beta) of
    nn.InstanceNorm2d(affine=True) (eps 1e-5, instance statistics also in eval) or None."""
    up = dict(stride=2, padding=1, output_padding=1)
    h = x
    if norm1 is not None:
        h = F.instance_norm(h, weight=norm1[0], bias=norm1[1], eps=1e-5)
    h = F.conv2d(F.leaky_relu(h, slope), conv1_w, conv1_b, padding=1)
    if norm2 is not None:
        h = F.instance_norm(h, weight=norm2[0], bias=norm2[1], eps=1e-5)
    h = F.conv_transpose2d(F.leaky_relu(h, slope), conv2_w, conv2_b, **up)
    return h + F.conv_transpose2d(x, bypass_w, bypass_b, **up)


def output_block(x, conv_w, conv_b, slope=0.1):
    """Output.forward with norm_layer=None (base_function.py:369-398): lrelu, ReflectionPad2d(1), conv3x3, tanh."""
    h = F.pad(F.leaky_relu(x, slope), (1, 1, 1, 1), mode='reflect')
    return torch.tanh(F.conv2d(h, conv_w, conv_b))


# --------------------------------------------------------------------------------------------
# SSIM (north_star: "plus SSIM parity on generated images")
# --------------------------------------------------------------------------------------------
def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True):
    """modules/evaluations/ssim.py:8-68 (`ssim`): per-channel 11x11 Gaussian (sigma 1.5) window, zero padding, C1 = 0.01^2,
    C2 = 0.03^2 (images in [0, 1] or [-1, 1]: the constants are used as they are, as in the reference)."""
    channel = img1.shape[1]
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])   # :8-10
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).float().expand(channel, 1, window_size, window_size).contiguous().to(img1)                # :12-16
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)                                                       # :19-20
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2                                                     # :22-24
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq                                 # :26-28
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2                                                                                   # :30-31
    ssim_map = ((2 * mu1_mu2 + c1) * (2 * sigma12 + c2)) / ((mu1_sq + mu2_sq + c1) * (sigma1_sq + sigma2_sq + c2))  # :33
    return ssim_map.mean() if size_average else ssim_map.mean(1).mean(1).mean(1)                                    # :35-38


# --------------------------------------------------------------------------------------------
# bounded samples of the attention workload for the CPU baseline legs of bench.py
# --------------------------------------------------------------------------------------------
def example_guided_attention_rows(src_mask, src_feature, ref_feature, conv_weight, rows: torch.Tensor):
    """modules/example_guided_att.py:21-36 restricted to the query pixels `rows` (a 1-D index tensor): the same
    arithmetic per output pixel (q-conv for all pixels, softmax over all S keys, both value products, the masked
    blend), only fewer rows of the S x S map. Returns [N, 2C, len(rows)]."""
    n, c, h, w = src_feature.shape
    query = F.conv2d(src_feature, conv_weight).reshape(n, -1, h * w)                 # :27-28
    att = torch.softmax(query[:, :, rows].permute(0, 2, 1) @ query, dim=-1)           # :30   [N, R, S]
    src_att = src_feature.reshape(n, c, -1) @ att.permute(0, 2, 1)                    # :18,:31
    ref_att = ref_feature.reshape(n, c, -1) @ att.permute(0, 2, 1)                    # :18,:32
    m = src_mask.reshape(n, 1, -1)[:, :, rows]
    flow = (1 - m) * ref_att + m * ref_feature.reshape(n, c, -1)[:, :, rows]          # :34
    return torch.cat([flow, src_att], dim=1)                                          # :36


# --------------------------------------------------------------------------------------------
# f2  pSp encoder pieces: IR-SE unit, map2style head, FPN add — eval mode
# --------------------------------------------------------------------------------------------
def batch_norm_eval(x, weight, bias, mean, var, eps=1e-5):
    """nn.BatchNorm2d in eval mode (running statistics): modules/psp/encoders/helpers.py:107,114."""
    return F.batch_norm(x, mean, var, weight, bias, False, 0.0, eps)


def se_module(x, fc1_w, fc2_w):
    """modules/psp/encoders/helpers.py:56-74 — x * sigmoid(fc2(relu(fc1(avgpool(x)))))."""
    m = x.mean(dim=(2, 3), keepdim=True)
    return x * torch.sigmoid(F.conv2d(torch.relu(F.conv2d(m, fc1_w)), fc2_w))


def bottleneck_ir_se(x, sd, stride):
    """modules/psp/encoders/helpers.py:97-119 with the module's state_dict `sd` (keys as the reference names them:
    res_layer.{0,1,2,3,4,5.fc1,5.fc2}, shortcut_layer.{0,1}); shortcut = MaxPool2d(1, stride) when in_channel == depth."""
    bn = lambda t, p: batch_norm_eval(t, sd[f"{p}.weight"], sd[f"{p}.bias"], sd[f"{p}.running_mean"], sd[f"{p}.running_var"])
    r = bn(x, "res_layer.0")
    r = F.conv2d(r, sd["res_layer.1.weight"], None, 1, 1)
    r = F.prelu(r, sd["res_layer.2.weight"])
    r = F.conv2d(r, sd["res_layer.3.weight"], None, stride, 1)
    r = bn(r, "res_layer.4")
    if "res_layer.5.fc1.weight" in sd:
        r = se_module(r, sd["res_layer.5.fc1.weight"], sd["res_layer.5.fc2.weight"])
    if "shortcut_layer.0.weight" in sd:
        sc = bn(F.conv2d(x, sd["shortcut_layer.0.weight"], None, stride), "shortcut_layer.1")
    else:
        sc = x[:, :, ::stride, ::stride]                     # MaxPool2d(kernel 1, stride)
    return r + sc


def gradual_style_block(x, sd, n_convs):
    """modules/psp/encoders/psp_encoders.py:13-37 — n x [Conv2d(3, stride 2, padding 1) + LeakyReLU(0.01)] down to 1x1, then
    EqualLinear (stylegan2/model.py:135-171: F.linear(x, weight * scale, bias * lr_mul), scale = 1/sqrt(in), lr_mul = 1)."""
    for i in range(n_convs):
        x = F.leaky_relu(F.conv2d(x, sd[f"convs.{2 * i}.weight"], sd[f"convs.{2 * i}.bias"], 2, 1), 0.01)
    x = x.view(-1, x.shape[1])
    w = sd["linear.weight"]
    return F.linear(x, w * (1.0 / math.sqrt(w.shape[1])), sd["linear.bias"])


def upsample_add(x, y):
    """modules/psp/encoders/psp_encoders.py:83-98."""
    return F.interpolate(x, size=y.shape[-2:], mode="bilinear", align_corners=True) + y


def gram_matrix(x):
    """modules/pluralistic_model/external_function.py:180-185 (GramMatrix): features @ features^T / (C*H*W) per image."""
    n, c, h, w = x.shape
    f = x.reshape(n, c, h * w)
    return torch.bmm(f, f.transpose(1, 2)) / (c * h * w)


def style_loss(x, target):
    """external_function.py:188-192 (StyleLoss): L1 between the Gram matrices."""
    return F.l1_loss(gram_matrix(x), gram_matrix(target).detach())


def contextual_loss(x, y, h=0.5):
    """external_function.py:231-274: channel-centred (batch mean of y), L2-normalised features, S x S cosine similarities,
    d = 1 - cos, d~ = d / (min_j d + 1e-5), w = exp((1 - d~) / h), cx = w / sum_j w, loss = mean_n -log(mean_j max_i cx + 1e-5)."""
    n, c = x.shape[:2]
    y_mu = y.mean(3).mean(2).mean(0).reshape(1, -1, 1, 1)
    xc, yc = x - y_mu, y - y_mu
    xn = (xc / torch.norm(xc, p=2, dim=1, keepdim=True)).reshape(n, c, -1)
    yn = (yc / torch.norm(yc, p=2, dim=1, keepdim=True)).reshape(n, c, -1)
    d = 1 - torch.bmm(xn.transpose(1, 2), yn)
    d_min, _ = torch.min(d, dim=2, keepdim=True)
    w = torch.exp((1 - d / (d_min + 1e-5)) / h)
    cx_ij = w / torch.sum(w, dim=2, keepdim=True)
    cx = torch.mean(torch.max(cx_ij, dim=1)[0], dim=1)
    return torch.mean(-torch.log(cx + 1e-5))
