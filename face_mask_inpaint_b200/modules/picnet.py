"""PICNet reference-fill generator around the hot-path kernels (the callers either side of the path).

Host-side mirror of `ReferenceFill` (modules/model.py:15-112) and of the PICNet networks it builds
(`ResEncoder`, `ResGenerator`: modules/pluralistic_model/network.py:73-293; their blocks: base_function.py:207-398),
with the SAME attribute / parameter names and shapes, so a reference checkpoint loads with strict=True.

What runs where:
  * `ExampleGuidedAttention` at 32^2 and `Auto_Attn` at 128^2 (58 % of the generator FLOPs), mask scaling and
    compositing: the sm_100a kernels of this package;
  * the spectral-normalised conv / conv-transpose blocks of the decoder and of the two encoders, and the Output block (SURVEY 8f
    rank 1): in inference one fused whole-network routine on the implicit-GEMM kernels of csrc/conv_blocks.cu (picnet_fast.py);
    under autograd (training) every wrapped convolution and every InstanceNorm + LeakyReLU pair runs forward AND backward through
    the per-layer Functions of ops.py (_ConvShared / _ConvTShared / _NormAct, channels_last between them). Only the 3-channel
    convolutions at the image boundary and anything with TF32 convolutions switched off stay PyTorch + cuDNN as in the reference.

Blocks are assembled from two small helpers instead of one class per variant: `_sn` wraps a conv in SpectralNorm
(external_function.py:16-72, one power iteration per forward, also in eval) and `_ResidualPair` is "main path + shortcut"
with optional average pooling — the reference's ResBlock / ResBlockEncoderOptimized / ResBlockDecoder differ only in
which convs, norms and pools go where. CoordConv (use_coord=True) is not reachable from the scripts and is refused.
"""
from __future__ import annotations

import functools

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from . import picnet_fast
from .attention import Auto_Attn, ExampleGuidedAttention
from .picnet_blocks import SpectralNorm


def _norm_factory(kind):
    """base_function.py:42-52."""
    if kind == 'none':
        return None
    if kind == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=True)
    if kind == 'batch':
        return functools.partial(nn.BatchNorm2d, momentum=0.1, affine=True)
    raise NotImplementedError(f"normalization layer [{kind}] is not found")


def _activation(kind):
    """base_function.py:55-67 (LeakyReLU slope is 0.1 there)."""
    table = {'ReLU': nn.ReLU, 'SELU': nn.SELU, 'PReLU': nn.PReLU, 'LeakyReLU': functools.partial(nn.LeakyReLU, 0.1)}
    if kind not in table:
        raise NotImplementedError(f"activation layer [{kind}] is not found")
    return table[kind]()


def _sn(conv, use_spect):
    return SpectralNorm(conv) if use_spect else conv


def _no_coord(use_coord):
    if use_coord:
        raise NotImplementedError("fmi_b200: CoordConv blocks are not reachable from the reference's scripts")


class _ResidualPair(nn.Module):
    """out = post(model(x)) + post(shortcut(x)); `model` and `shortcut` are nn.Sequential with the reference's layout,
    conv1 / conv2 / bypass are also direct attributes (that is how the reference registers them: the same module
    appears under two names in the state_dict, e.g. `conv1.module.weight_bar` and `model.1.module.weight_bar`)."""

    def __init__(self, conv1, conv2, bypass, main, shortcut, post=None):
        super().__init__()
        self.conv1, self.conv2, self.bypass = conv1, conv2, bypass
        self.model = nn.Sequential(*main)
        self.shortcut = nn.Sequential(*shortcut)
        if post is not None:
            self.pool = post
        self.sample = post is not None

    def forward(self, x):
        # under autograd the norm + activation pairs and the wrapped convolutions run forward AND backward on the kernels
        # (ops.run_block_sequential / SpectralNorm.forward -> ops.conv_train); otherwise this is the reference's Sequential
        main = ops.run_block_sequential(self.model, x)
        if self.sample:
            return ops.pool_sum(self.pool, main, ops.run_block_sequential(self.shortcut, x))
        return main + ops.run_block_sequential(self.shortcut, x)


def _pre_act(norm, act, channels):
    return ([norm(channels)] if norm is not None else []) + [act]


class ResBlock(_ResidualPair):
    """base_function.py:207-268: [norm] act conv3x3 [norm] act conv3x3 (+ 1x1 shortcut), optional 2x average pooling."""

    def __init__(self, input_nc, output_nc, hidden_nc=None, norm_layer=nn.BatchNorm2d, nonlinearity=None,
                 sample_type='none', use_spect=False, use_coord=False):
        _no_coord(use_coord)
        nonlinearity = nonlinearity if nonlinearity is not None else nn.LeakyReLU()
        hidden_nc = output_nc if hidden_nc is None else hidden_nc
        post = None
        if sample_type == 'down':
            post = nn.AvgPool2d(kernel_size=2, stride=2)
        elif sample_type == 'up':
            output_nc, post = output_nc * 4, nn.PixelShuffle(upscale_factor=2)
        elif sample_type != 'none':
            raise NotImplementedError(f"sample type [{sample_type}] is not found")
        conv1 = _sn(nn.Conv2d(input_nc, hidden_nc, 3, 1, 1), use_spect)
        conv2 = _sn(nn.Conv2d(hidden_nc, output_nc, 3, 1, 1), use_spect)
        bypass = _sn(nn.Conv2d(input_nc, output_nc, 1, 1, 0), use_spect)
        main = _pre_act(norm_layer, nonlinearity, input_nc) + [conv1] + _pre_act(norm_layer, nonlinearity, hidden_nc) + [conv2]
        super().__init__(conv1, conv2, bypass, main, [bypass], post)


class ResBlockEncoderOptimized(_ResidualPair):
    """base_function.py:271-305: conv [norm] act conv pool, shortcut = pool then 1x1 conv (first encoder layer)."""

    def __init__(self, input_nc, output_nc, norm_layer=nn.BatchNorm2d, nonlinearity=None, use_spect=False,
                 use_coord=False):
        _no_coord(use_coord)
        nonlinearity = nonlinearity if nonlinearity is not None else nn.LeakyReLU()
        conv1 = _sn(nn.Conv2d(input_nc, output_nc, 3, 1, 1), use_spect)
        conv2 = _sn(nn.Conv2d(output_nc, output_nc, 3, 1, 1), use_spect)
        bypass = _sn(nn.Conv2d(input_nc, output_nc, 1, 1, 0), use_spect)
        main = [conv1] + _pre_act(norm_layer, nonlinearity, output_nc) + [conv2, nn.AvgPool2d(kernel_size=2, stride=2)]
        super().__init__(conv1, conv2, bypass, main, [nn.AvgPool2d(kernel_size=2, stride=2), bypass])


class ResBlockDecoder(_ResidualPair):
    """base_function.py:308-366: [norm] act conv3x3 [norm] act convT3x3(stride 2) + convT3x3(stride 2) shortcut."""

    def __init__(self, input_nc, output_nc, hidden_nc=None, norm_layer=nn.BatchNorm2d, nonlinearity=None,
                 use_spect=False, use_coord=False):
        nonlinearity = nonlinearity if nonlinearity is not None else nn.LeakyReLU()
        hidden_nc = output_nc if hidden_nc is None else hidden_nc
        up = dict(kernel_size=3, stride=2, padding=1, output_padding=1)
        conv1 = _sn(nn.Conv2d(input_nc, hidden_nc, 3, 1, 1), use_spect)
        conv2 = _sn(nn.ConvTranspose2d(hidden_nc, output_nc, **up), use_spect)
        bypass = _sn(nn.ConvTranspose2d(input_nc, output_nc, **up), use_spect)
        main = _pre_act(norm_layer, nonlinearity, input_nc) + [conv1] + _pre_act(norm_layer, nonlinearity, hidden_nc) + [conv2]
        super().__init__(conv1, conv2, bypass, main, [bypass])


class Output(nn.Module):
    """base_function.py:369-398: [norm] act reflect-pad conv tanh."""

    def __init__(self, input_nc, output_nc, kernel_size=3, norm_layer=nn.BatchNorm2d, nonlinearity=None, use_spect=False,
                 use_coord=False):
        super().__init__()
        _no_coord(use_coord)
        nonlinearity = nonlinearity if nonlinearity is not None else nn.LeakyReLU()
        self.conv1 = _sn(nn.Conv2d(input_nc, output_nc, kernel_size, padding=0, bias=True), use_spect)
        self.model = nn.Sequential(*(_pre_act(norm_layer, nonlinearity, input_nc)
                                     + [nn.ReflectionPad2d(kernel_size // 2), self.conv1, nn.Tanh()]))

    def forward(self, x):
        return ops.run_block_sequential(self.model, x)     # training: activation + reflection padding without leaving NHWC


def _width(ngf, img_f, level):
    """Channel count at pyramid level `level`: ngf * 2^level capped at img_f (network.py:111-112, :213, :225-228)."""
    return ngf * min(2 ** level, img_f // ngf)


class ResEncoder(nn.Module):
    """network.py:73-172. Returns ([mu, softplus(std)], features)."""

    def __init__(self, input_nc=3, ngf=64, z_nc=128, img_f=1024, L=6, layers=6, norm='none', activation='ReLU',
                 use_spect=True, use_coord=False, encoder_type='src'):
        super().__init__()
        self.layers, self.z_nc, self.L = layers, z_nc, L
        self.ecnoder_type = encoder_type  # (sic) attribute name of the reference
        norm_layer, act = _norm_factory(norm), _activation(activation)
        self.block0 = ResBlockEncoderOptimized(input_nc, ngf, norm_layer, act, use_spect, use_coord)
        for i in range(layers - 1):
            cin, cout = _width(ngf, img_f, i), _width(ngf, img_f, i + 1)
            self.add_module(f'encoder{i}', ResBlock(cin, cout, cin, norm_layer, act, 'none' if i % 2 == 0 else 'down',
                                                    use_spect, use_coord))
        top = _width(ngf, img_f, layers - 1)
        if encoder_type == 'src':
            for i in range(L):
                self.add_module(f'infer_prior{i}', ResBlock(top, top, top, norm_layer, act, 'none', use_spect, use_coord))
            self.prior = ResBlock(top, 2 * z_nc, top, norm_layer, act, 'none', use_spect, use_coord)
        elif encoder_type == 'ref':
            self.posterior = ResBlock(top, 2 * z_nc, top, norm_layer, act, 'none', use_spect, use_coord)

    def _distribution(self, o):
        mu, std = torch.split(o, self.z_nc, dim=1)
        return [mu, F.softplus(std)]

    def prior_path(self, encoded):
        for i in range(self.L):
            encoded = getattr(self, f'infer_prior{i}')(encoded)
        return self._distribution(self.prior(encoded))

    def post_path(self, encoded):
        return self._distribution(self.posterior(encoded))

    def forward(self, img):
        if picnet_fast.encoder_supported(self, img):   # inference: the residual blocks on the implicit-GEMM kernels
            return picnet_fast.encoder_forward(self, img)
        out = self.block0(img)
        for i in range(self.layers - 1):
            out = getattr(self, f'encoder{i}')(out)
        if self.ecnoder_type == 'src':
            return self.prior_path(out), out
        if self.ecnoder_type == 'ref':
            return self.post_path(out), out
        return None


class ResGenerator(nn.Module):
    """network.py:175-293. `attn1` (after decoder1) is this package's Auto_Attn: the fused tcgen05 attention kernel."""

    def __init__(self, output_nc=3, ngf=64, z_nc=128, img_f=1024, L=1, layers=6, norm='batch', activation='ReLU',
                 use_spect=True, use_coord=False, use_attn=False):
        super().__init__()
        self.layers, self.L, self.use_attn = layers, L, use_attn
        norm_layer, act = _norm_factory(norm), _activation(activation)
        ch = _width(ngf, img_f, layers - 1)
        self.generator = ResBlock(z_nc, ch, ch, None, act, 'none', use_spect, use_coord)
        for i in range(L):
            self.add_module(f'generator{i}', ResBlock(ch, ch, ch, None, act, 'none', use_spect, use_coord))
        prev = ch
        for i in range(layers):
            ch = _width(ngf, img_f, layers - i - 1)
            self.add_module(f'decoder{i}', ResBlockDecoder(prev, ch, ch, norm_layer, act, use_spect, use_coord))
            if i > layers - 2:
                self.add_module(f'out{i}', Output(ch, output_nc, 3, None, act, use_spect, use_coord))
            if i == 1 and use_attn:
                self.add_module(f'attn{i}', Auto_Attn(ch, None))
            prev = ch

    def forward(self, encoded, z=None, f_e=None, mask=None, pool_to=None):
        """`pool_to` (not in the reference's signature; ReferenceFill passes its AdaptiveAvgPool2d size): return the pooled
        image, which the kernel path fuses into the Output block instead of writing and re-reading the full-size image."""
        if picnet_fast.supported(self, encoded):   # inference: the conv blocks on the implicit-GEMM kernels (csrc/conv_blocks.cu)
            if z is not None and not self._fmi_z_ok:     # unusual z -> f blocks: cuDNN, then the kernels
                f = self.generator(z)
                for i in range(self.L):
                    f = getattr(self, f'generator{i}')(f)
                encoded, z = encoded + f, None
            return picnet_fast.decoder_forward(self, encoded, f_e, mask, pool_to=pool_to, z=z)
        out = encoded
        if z is not None:
            f = self.generator(z)
            for i in range(self.L):
                f = getattr(self, f'generator{i}')(f)
            out = encoded + f
        output = None
        for i in range(self.layers):
            out = getattr(self, f'decoder{i}')(out)
            if i == 1 and self.use_attn:
                out, _ = getattr(self, f'attn{i}')(out, f_e, mask)
            if i > self.layers - 2:
                output = getattr(self, f'out{i}')(out)
                if i + 1 < self.layers:     # network.py:272 also concatenates after the LAST layer, where nothing reads it
                    out = torch.cat([out, output], dim=1)     # (35 channels at 1024^2, written and differentiated for nothing)
        if pool_to is not None:
            output = F.adaptive_avg_pool2d(output, pool_to)
        return output

    def get_z(self, src_distribution, ref_distribution, return_zq=False, mask=None):
        """network.py:270-293: reparameterised samples of the prior (source) and posterior (reference)."""
        p_mu, p_sigma = ref_distribution
        q_mu, q_sigma = src_distribution
        # argument validation reads a flag back to the host: kept in eager mode, skipped while a CUDA graph is capturing
        check = not (p_mu.is_cuda and torch.cuda.is_current_stream_capturing())
        z_p = torch.distributions.Normal(p_mu, p_sigma, validate_args=check).rsample()
        z_q = torch.distributions.Normal(q_mu, q_sigma, validate_args=check).rsample()
        return z_q if return_zq else torch.cat([z_q, z_p], dim=1)


def init_orthogonal(net, gain=0.02):
    """init_weights(init_type='orthogonal') of base_function.py:13-39, applied to every conv/linear weight the
    reference's class-name test would hit (SpectralNorm-wrapped convs have `weight_bar`, no `weight`: skipped there too)."""
    for m in net.modules():
        name = type(m).__name__
        w = getattr(m, 'weight', None)
        if isinstance(w, torch.Tensor) and ('Conv' in name or 'Linear' in name):
            nn.init.orthogonal_(w.data, gain=gain)
            if getattr(m, 'bias', None) is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif 'BatchNorm2d' in name and isinstance(w, torch.Tensor):
            nn.init.normal_(w.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0.0)
    return net


def define_e(encoder_type='src', input_nc=3, ngf=64, z_nc=512, img_f=512, L=6, layers=5, norm='none', activation='ReLU',
             use_spect=True, use_coord=False, init_type='orthogonal', gpu_ids=()):
    """network.py:10-27 (orthogonal init only; DataParallel wrapping is replaced by dist.py's batch sharding)."""
    if init_type != 'orthogonal':
        raise NotImplementedError("fmi_b200: only init_type='orthogonal' (the scripts' default) is mirrored")
    return init_orthogonal(ResEncoder(input_nc, ngf, z_nc, img_f, L, layers, norm, activation, use_spect, use_coord,
                                      encoder_type))


def define_g(output_nc=3, ngf=64, z_nc=512, img_f=512, L=1, layers=5, norm='instance', activation='ReLU', use_spect=True,
             use_coord=False, use_attn=True, init_type='orthogonal', gpu_ids=()):
    """network.py:30-47."""
    if init_type != 'orthogonal':
        raise NotImplementedError("fmi_b200: only init_type='orthogonal' (the scripts' default) is mirrored")
    return init_orthogonal(ResGenerator(output_nc, ngf, z_nc, img_f, L, layers, norm, activation, use_spect, use_coord,
                                        use_attn))


def two_encoders(model, src_image, ref_image):
    """model.py:87-88: the source and reference encoders are independent. In inference they run concurrently on two streams
    (each is a chain of ~60 small dependent kernels that leaves most of the GPU idle); under CUDA-graph capture the fork / join
    becomes two parallel branches of the graph. The side stream always waits for the current stream before it starts, so
    blocks of its allocator pool are only reused after every earlier consumer has been enqueued."""
    import os
    if (torch.is_grad_enabled() or not src_image.is_cuda or os.environ.get("FMI_ENCODER_STREAMS") == "0"):
        return model.src_encoder(src_image), model.ref_encoder(ref_image)
    cur = torch.cuda.current_stream()
    from ..graphs import module_cache
    cache = module_cache(model)
    side = cache.get("side_stream")
    if side is None or side.device != src_image.device:
        side = cache["side_stream"] = torch.cuda.Stream(device=src_image.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        ref_out = model.ref_encoder(ref_image)
    src_out = model.src_encoder(src_image)
    cur.wait_stream(side)
    return src_out, ref_out


class ReferenceFill(nn.Module):
    """modules/model.py:15-112 with encoder type 'pluralistic' (the README / PICNet_inference.py configuration)."""

    def __init__(self, mask_detector, encoder_params, decoder_params, use_att=True, out_size=(256, 256)):
        super().__init__()
        self.mask_detector = mask_detector
        encoder_params = dict(encoder_params)
        self.encoder_type = encoder_params.pop('type')
        if self.encoder_type != 'pluralistic':
            raise NotImplementedError("fmi_b200: ReferenceFill mirrors the 'pluralistic' encoders (drn is out of scope)")
        self.src_encoder = define_e(**encoder_params, encoder_type='src')
        self.ref_encoder = define_e(**encoder_params, encoder_type='ref')
        self.decoder = define_g(**dict(decoder_params))
        self.use_att = use_att
        if use_att:
            self.attention = ExampleGuidedAttention(encoder_params['img_f'])
        self.pool = nn.AdaptiveAvgPool2d(out_size)

    def forward(self, src_image, ref_image, src_mask=None, resize=True, no_prior=False):
        if src_mask is None:
            src_mask = self.mask_detector(src_image, mode='eval')
        (src_dist, src_features), (ref_dist, ref_features) = two_encoders(self, src_image, ref_image)
        mask_full = src_mask.unsqueeze(1)
        if self.use_att:
            scaled = ops.scale_img(mask_full, src_features.shape[-2:])           # model.py:96 -> fmi_scale_mask
            enc_features = self.attention(scaled, src_features, ref_features)    # :97 -> fmi_attn_fwd
        else:
            enc_features = ops.composite(src_features, ref_features, mask_full)  # :99 -> fmi_composite
        if no_prior:
            dec_image = self.decoder(enc_features)
        else:
            z = self.decoder.get_z(src_dist, ref_dist, return_zq=not self.use_att)
            # model.py:111 (`self.pool(dec_image)`) handed to the decoder so that its kernel path can fuse it
            dec_image = self.decoder(enc_features, z=z, pool_to=self.pool.output_size if resize else None)
        if resize and no_prior:
            dec_image = F.interpolate(dec_image, size=(218, 178), mode='bilinear', align_corners=True)
        return dec_image


# the configuration of README.md:58-70 / PICNet_inference.py:38-58 (BASELINE config 1)
PICNET_REF_ENCODER = dict(type='pluralistic', ngf=32, z_nc=128, img_f=128, layers=5, norm='none', activation='LeakyReLU',
                          init_type='orthogonal')
PICNET_REF_DECODER = dict(ngf=32, z_nc=256, img_f=256, L=0, layers=5, norm='instance', activation='LeakyReLU',
                          init_type='orthogonal')


def build_picnet_ref(use_att=True):
    return ReferenceFill(None, PICNET_REF_ENCODER, PICNET_REF_DECODER, use_att=use_att)
