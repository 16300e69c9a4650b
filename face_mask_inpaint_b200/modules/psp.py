"""RefpSp (reference-guided pixel2style2pixel) around the hot-path kernels.

Host-side mirror of `pSp` (modules/psp/psp.py:22-130) with the `GradualStyleEncoder` (modules/psp/encoders/
psp_encoders.py:40-151; IR-SE50 units: encoders/helpers.py:56-119), same attribute / parameter names and shapes, so
reference checkpoints (`encoder.*`, `decoder.*`, `latent_avg`) load unchanged.

What runs where:
  * decoder = this package's StyleGAN2 `Generator` (modulated-conv / blur / ToRGB kernels, NHWC bf16 or tf32);
  * `attention1` (C=512 @16^2, out_conv) and `attention2` (C=256 @32^2): the fused attention kernel;
  * the masked source/reference blends (psp_encoders.py:127-138): `fmi_composite` with the mask sampled in the kernel;
  * the IR-SE50 trunk, the FPN adds and the 18 map2style heads: in inference the implicit-GEMM kernels of
    csrc/ir_encoder.cu (modules/psp_fast.py: eval-mode BatchNorm folded, source and reference as ONE 2N batch); under
    autograd / in train mode PyTorch + cuDNN as in the reference (two passes, because batch statistics differ).
"""
from __future__ import annotations

import math
from argparse import Namespace

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from .attention import ExampleGuidedAttention
from .stylegan2 import EqualLinear, Generator

# IR-50 layout (helpers.py:29-36): (input channels, depth, units); the first unit of every group has stride 2
_IR50 = ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3))
_TAPS = {6: 'c1', 20: 'c2', 23: 'c3'}   # trunk outputs used by the pyramid (psp_encoders.py:103-108)


def _trunk_layout(x):
    """The IR-SE50 trunk under autograd stays cuDNN / ATen (helpers.py:97-119 with BatchNorm batch statistics); fed an NCHW batch,
    cuDNN's TF32 kernels transpose every activation to NHWC and back (torch.profiler, RefpSp train step: 1 600 launches of its
    nchwToNhwc / nhwcToNchw kernels = 9.2 of 54 ms of GPU time). A channels_last input keeps the whole trunk — convolutions,
    BatchNorm, PReLU, pooling, SE — in NHWC. FMI_PSP_CHANNELS_LAST=0 keeps the reference's layout."""
    import os
    if x.is_cuda and x.dim() == 4 and torch.is_grad_enabled() and os.environ.get("FMI_PSP_CHANNELS_LAST") != "0":
        return x.contiguous(memory_format=torch.channels_last)
    return x



class _SqueezeExcite(nn.Module):
    """helpers.py:56-74."""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        return x * self.sigmoid(self.fc2(self.relu(self.fc1(self.avg_pool(x)))))


class _IRUnit(nn.Module):
    """bottleneck_IR / bottleneck_IR_SE (helpers.py:77-119): BN conv3x3 PReLU conv3x3(stride) BN [SE] + shortcut."""

    def __init__(self, in_channel, depth, stride, se):
        super().__init__()
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, 1, stride, bias=False), nn.BatchNorm2d(depth))
        layers = [nn.BatchNorm2d(in_channel), nn.Conv2d(in_channel, depth, 3, 1, 1, bias=False), nn.PReLU(depth),
                  nn.Conv2d(depth, depth, 3, stride, 1, bias=False), nn.BatchNorm2d(depth)]
        if se:
            layers.append(_SqueezeExcite(depth, 16))
        self.res_layer = nn.Sequential(*layers)

    def forward(self, x):
        return self.res_layer(x) + self.shortcut_layer(x)


class GradualStyleBlock(nn.Module):
    """psp_encoders.py:13-37: log2(spatial) stride-2 convs down to 1x1, then an EqualLinear."""

    def __init__(self, in_c, out_c, spatial):
        super().__init__()
        self.out_c, self.spatial = out_c, spatial
        layers, c = [], in_c
        for _ in range(int(math.log2(spatial))):
            layers += [nn.Conv2d(c, out_c, kernel_size=3, stride=2, padding=1), nn.LeakyReLU()]
            c = out_c
        self.convs = nn.Sequential(*layers)
        self.linear = EqualLinear(out_c, out_c, lr_mul=1)

    def forward(self, x):
        return self.linear(self.convs(x).view(-1, self.out_c))


class GradualStyleEncoder(nn.Module):
    """psp_encoders.py:40-151."""

    def __init__(self, num_layers, mode='ir', opts=None):
        super().__init__()
        if num_layers != 50 or mode not in ('ir', 'ir_se'):
            raise NotImplementedError("fmi_b200: the scripts build GradualStyleEncoder(50, 'ir_se'); other depths are not mirrored")
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, 3, 1, 1, bias=False), nn.BatchNorm2d(64), nn.PReLU(64))
        units = []
        for in_c, depth, n in _IR50:
            units += [_IRUnit(in_c if k == 0 else depth, depth, 2 if k == 0 else 1, mode == 'ir_se') for k in range(n)]
        self.body = nn.Sequential(*units)
        self.style_count = opts.n_styles
        self.coarse_ind, self.middle_ind = 3, 7
        self.styles = nn.ModuleList(
            GradualStyleBlock(512, 512, 16 if i < self.coarse_ind else 32 if i < self.middle_ind else 64)
            for i in range(self.style_count))
        self.latlayer1 = nn.Conv2d(256, 512, kernel_size=1, stride=1, padding=0)
        self.latlayer2 = nn.Conv2d(128, 512, kernel_size=1, stride=1, padding=0)
        self.use_attention = opts.use_attention
        if opts.use_attention:
            self.attention1 = ExampleGuidedAttention(512, out_channels=512)
            self.attention2 = ExampleGuidedAttention(256, out_channels=256)

    def _trunk(self, x):
        x = self.input_layer(_trunk_layout(x))
        taps = {}
        for i, unit in enumerate(self.body):
            x = unit(x)
            if i in _TAPS:
                taps[_TAPS[i]] = x
        return taps['c1'], taps['c2'], taps['c3']

    @staticmethod
    def _upsample_add(x, y):
        return F.interpolate(x, size=y.shape[-2:], mode='bilinear', align_corners=True) + y

    def forward(self, x, ref=None, mask=None):
        from . import psp_fast
        if psp_fast.supported(self, x, ref):      # inference: trunk, FPN and heads on the sm_100a kernels (psp_fast.py)
            return psp_fast.encoder_forward(self, x, ref, mask)
        if ref is None:
            c1, c2, c3 = self._trunk(x)
        else:
            if mask is None:
                raise AssertionError("ref and mask should both be provided")
            if self.training:
                (c1, c2, c3), (r1, r2, r3) = self._trunk(x), self._trunk(ref)
            else:
                n = x.shape[0]
                b1, b2, b3 = self._trunk(torch.cat([x, ref], dim=0))
                (c1, r1), (c2, r2), (c3, r3) = (t.split(n) for t in (b1, b2, b3))
            mask_full = mask.unsqueeze(1)  # [N, 1, 256, 256]
            if self.use_attention:
                c3 = self.attention1(ops.scale_img(mask_full, r3.shape[-2:]), c3, r3)  # psp_encoders.py:132
                c2 = self.attention2(ops.scale_img(mask_full, r2.shape[-2:]), c2, r2)  # :133
            else:
                c3 = ops.composite(c3.contiguous(), r3.contiguous(), mask_full)        # :135
                c2 = ops.composite(c2.contiguous(), r2.contiguous(), mask_full)        # :136
            c1 = ops.composite(c1.contiguous(), r1.contiguous(), mask_full)            # :138
        latents = [self.styles[j](c3) for j in range(self.coarse_ind)]
        p2 = self._upsample_add(c3, self.latlayer1(c2))
        latents += [self.styles[j](p2) for j in range(self.coarse_ind, self.middle_ind)]
        p1 = self._upsample_add(p2, self.latlayer2(c1))
        latents += [self.styles[j](p1) for j in range(self.middle_ind, self.style_count)]
        return torch.stack(latents, dim=1)


def _sub_state(d, prefix):
    d = d.get('state_dict', d)
    return {k[len(prefix) + 1:]: v for k, v in d.items() if k.startswith(prefix + '.')}


class pSp(nn.Module):
    """psp.py:22-130. `opts` needs: output_size, encoder_type, use_attention, train_decoder, start_from_latent_avg,
    learn_in_w, pt_ckpt_path, stylegan_weights (the fields the reference reads)."""

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        self.opts.n_styles = int(math.log(self.opts.output_size, 2)) * 2 - 2
        if self.opts.encoder_type != 'GradualStyleEncoder':
            raise NotImplementedError("fmi_b200: only the GradualStyleEncoder (the scripts' default) is mirrored")
        self.encoder = GradualStyleEncoder(50, 'ir_se', self.opts)
        self.decoder = Generator(self.opts.output_size, 512, 8)
        if not opts.train_decoder:
            for p in self.decoder.parameters():
                p.requires_grad = False
        self.face_pool = nn.AdaptiveAvgPool2d((256, 256))
        self.latent_avg = None
        self.load_weights()

    def load_weights(self):
        """psp.py:50-70. Without any checkpoint path the networks keep their random init (throughput runs, tests) and
        `latent_avg` is zero — the reference would fail here because its pretrained files are not shipped."""
        o = self.opts
        if getattr(o, 'pt_ckpt_path', None):
            ckpt = torch.load(o.pt_ckpt_path, map_location='cpu')
            self.encoder.load_state_dict(_sub_state(ckpt, 'encoder'), strict=False)
            self.decoder.load_state_dict(_sub_state(ckpt, 'decoder'), strict=True)
            self._load_latent_avg(ckpt, None)
        elif getattr(o, 'stylegan_weights', None):
            ckpt = torch.load(o.stylegan_weights, map_location='cpu')
            self.decoder.load_state_dict(ckpt['g_ema'], strict=False)
            self._load_latent_avg(ckpt, 1 if o.learn_in_w else o.n_styles)
        elif getattr(o, 'start_from_latent_avg', False):
            self.latent_avg = torch.zeros(1 if o.learn_in_w else o.n_styles, 512)

    def _load_latent_avg(self, ckpt, repeat):
        self.latent_avg = ckpt.get('latent_avg')
        if self.latent_avg is not None and repeat is not None:
            self.latent_avg = self.latent_avg.repeat(repeat, 1)

    def forward(self, x, ref=None, src_mask=None, resize=True, latent_mask=None, input_code=False, randomize_noise=True,
                inject_latent=None, return_latents=False, alpha=None):
        if input_code:
            codes = x
        else:
            codes = self.encoder(x, ref=ref, mask=src_mask)  # [N, n_styles, 512]
            if self.opts.start_from_latent_avg and self.latent_avg is not None:
                if self.latent_avg.device != codes.device:   # moved once and kept: no host copy on later calls (graph capture)
                    self.latent_avg = self.latent_avg.to(codes.device)
                avg = self.latent_avg
                codes = codes + (avg.repeat(codes.shape[0], 1) if self.opts.learn_in_w else avg.repeat(codes.shape[0], 1, 1))
        if latent_mask is not None:
            for i in latent_mask:
                if inject_latent is None:
                    codes[:, i] = 0
                elif alpha is not None:
                    codes[:, i] = alpha * inject_latent[:, i] + (1 - alpha) * codes[:, i]
                else:
                    codes[:, i] = inject_latent[:, i]
        images, result_latent = self.decoder([codes], input_is_latent=not input_code, randomize_noise=randomize_noise,
                                             return_latents=return_latents)
        if resize:
            images = ops.adaptive_avg_pool(images, self.face_pool.output_size)   # exact 4x4 mean: fmi_avgpool_planes
        return (images, result_latent) if return_latents else images


def refpsp_opts(output_size=1024, use_attention=1, train_decoder=0):
    """The configuration of BASELINE config 3 / SURVEY 8d (psp_inference.py defaults, no checkpoint files)."""
    return Namespace(output_size=output_size, encoder_type='GradualStyleEncoder', use_attention=use_attention,
                     train_decoder=train_decoder, start_from_latent_avg=1, learn_in_w=0, pt_ckpt_path=None,
                     stylegan_weights=None)
