"""Generate tests/golden/*.npz by running the REFERENCE's own modules (imported from /root/reference, CPU, fp32)
on seeded inputs. Run in the build container only (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py

What executes reference code, and what is stubbed:
  * modules/example_guided_att.py, modules/pluralistic_model/base_function.py (Auto_Attn), modules/model.py
    (scale_img) and modules/psp/stylegan2/model.py (ModulatedConv2d, StyledConv, ToRGB, Generator) are imported
    unmodified.
  * modules/psp/stylegan2/op is CUDA-only (JIT-built extensions). Its package is pre-seeded with:
      - upfirdn2d: the reference's OWN `upfirdn2d_native` (op/upfirdn2d.py:150-184), extracted from the file's AST
        and executed with `F` injected (the file forgets to import it);
      - fused_leaky_relu: scale * leaky_relu(x + bias) restated from op/fused_bias_act_kernel.cu:26-47 — this one op
        has no runnable reference on a CPU, its golden is therefore a restatement (said so in DESIGN.md).
"""
import ast
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
sys.path.insert(0, str(OUT.parent.parent))
sys.path.insert(0, str(OUT.parent))


def _load_upfirdn2d_native():
    src = (REF / "modules/psp/stylegan2/op/upfirdn2d.py").read_text()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "upfirdn2d_native"][0]
    ns = {"F": F, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "upfirdn2d_native", "exec"), ns)
    return ns["upfirdn2d_native"]


upfirdn2d_native = _load_upfirdn2d_native()


def _install_op_stub():
    def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
        n, c, h, w = input.shape
        out = upfirdn2d_native(input.reshape(-1, h, w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        return out.reshape(n, c, out.shape[1], out.shape[2])

    def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
        shape = [1, -1] + [1] * (input.ndim - 2)
        return scale * F.leaky_relu(input + bias.view(*shape), negative_slope)

    class FusedLeakyReLU(nn.Module):
        def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
            super().__init__()
            self.bias = nn.Parameter(torch.zeros(channel))
            self.negative_slope = negative_slope
            self.scale = scale

        def forward(self, input):
            return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)

    pkg = types.ModuleType("modules.psp.stylegan2.op")
    pkg.upfirdn2d, pkg.fused_leaky_relu, pkg.FusedLeakyReLU = upfirdn2d, fused_leaky_relu, FusedLeakyReLU
    pkg.__path__ = []
    sys.modules["modules.psp.stylegan2.op"] = pkg


_install_op_stub()
from modules.example_guided_att import ExampleGuidedAttention  # noqa: E402
from modules.pluralistic_model.base_function import Auto_Attn  # noqa: E402
from modules.model import scale_img  # noqa: E402
from modules.psp.stylegan2 import model as sg2  # noqa: E402

from golden_util import build_generator32, randomize  # noqa: E402  (deterministic parameter construction)


def np_(t):
    return t.detach().cpu().numpy()


def binary_mask(n, g):
    m = (torch.rand(n, 1, 256, 256, generator=g) < 0.3).float()
    m[:, :, 128:230, 50:206] = 1.0
    return m


def gen_attention():
    g = torch.Generator().manual_seed(100)
    out = {}
    # ExampleGuidedAttention, with and without out_conv
    for tag, c, hw, oc in [("ega_plain", 32, 16, None), ("ega_outconv", 64, 8, 48)]:
        torch.manual_seed(1)
        mod = ExampleGuidedAttention(c, oc)
        with torch.no_grad():
            mod.conv.weight.mul_(2.0)  # logit std ~ 4 (SURVEY 8d range): a near-uniform softmax would hide bugs
        src = torch.randn(2, c, hw, hw, generator=g)
        ref = torch.randn(2, c, hw, hw, generator=g)
        mask = scale_img(binary_mask(2, g), (hw, hw))
        y = mod(mask, src, ref)
        out.update({f"{tag}.src": np_(src), f"{tag}.ref": np_(ref), f"{tag}.mask": np_(mask), f"{tag}.out": np_(y),
                    f"{tag}.conv_w": np_(mod.conv.weight)})
        if oc:
            out.update({f"{tag}.oc_w": np_(mod.out_conv.weight), f"{tag}.oc_b": np_(mod.out_conv.bias)})
    # Auto_Attn: pre=None (the live call) and the pre/mask branch with `model` replaced by identity to expose its input
    torch.manual_seed(2)
    c, hw = 32, 16
    mod = Auto_Attn(c, None)
    with torch.no_grad():
        mod.query_conv.weight.mul_(2.0)
        mod.gamma.fill_(0.7)
        mod.alpha.fill_(1.3)
    x = torch.randn(2, c, hw, hw, generator=g)
    pre = torch.randn(2, c, hw, hw, generator=g)
    mask = scale_img(binary_mask(2, g), (hw, hw))
    y, attn = mod(x)
    mod.model = nn.Identity()
    cat, _ = mod(x, pre, mask)
    out.update({"auto.x": np_(x), "auto.pre": np_(pre), "auto.mask": np_(mask), "auto.out": np_(y),
                "auto.attn": np_(attn), "auto.cat": np_(cat), "auto.q_w": np_(mod.query_conv.weight),
                "auto.q_b": np_(mod.query_conv.bias), "auto.gamma": np_(mod.gamma), "auto.alpha": np_(mod.alpha)})
    np.savez_compressed(OUT / "attention.npz", **out)


def gen_upfirdn2d_and_composite():
    g = torch.Generator().manual_seed(200)
    out = {}
    cases = [("blur", [1, 3, 3, 1], 4.0, 1, 1, (1, 1), (2, 3, 9, 9)),
             ("blur_bwd", [1, 3, 3, 1], 4.0, 1, 1, (2, 2), (2, 3, 8, 8)),
             ("up", [1, 3, 3, 1], 4.0, 2, 1, (2, 1), (2, 3, 6, 6)),
             ("down", [1, 3, 3, 1], 1.0, 1, 2, (1, 1), (2, 3, 8, 8)),
             ("k3", [1, 2, 1], 1.0, 1, 1, (1, 1), (1, 2, 7, 5)),
             ("odd", [1, 4, 6, 4, 1], 1.0, 3, 2, (3, 2), (1, 2, 5, 6)),
             ("crop", [1, 3, 3, 1], 1.0, 1, 1, (-1, -2), (1, 2, 10, 10))]
    for tag, taps, gain, up, down, pad, shape in cases:
        k = sg2.make_kernel(taps) * gain
        if len(taps) == 5:
            k = k + 0.01 * torch.arange(25.).view(5, 5)
        x = torch.randn(*shape, generator=g)
        n, c, h, w = shape
        y = upfirdn2d_native(x.reshape(-1, h, w, 1), k, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        out.update({f"ufd.{tag}.x": np_(x), f"ufd.{tag}.k": np_(k), f"ufd.{tag}.y": np_(y.reshape(n, c, y.shape[1], y.shape[2])),
                    f"ufd.{tag}.cfg": np.array([up, down, pad[0], pad[1]])})
    # minor > 1 through the op-level signature
    x = torch.randn(3, 6, 7, 2, generator=g)
    k = sg2.make_kernel([1, 3, 3, 1])
    out.update({"ufd.minor.x": np_(x), "ufd.minor.k": np_(k),
                "ufd.minor.y": np_(upfirdn2d_native(x, k, 2, 1, 1, 2, 1, 2, 0, 1))})
    # compositing (modules/model.py:95-99)
    mask = binary_mask(2, g)
    for tag, shape in [("c32", (2, 8, 32, 32)), ("c16", (2, 4, 16, 16)), ("c7x9", (2, 3, 7, 9))]:
        src = torch.randn(*shape, generator=g)
        ref = torch.randn(*shape, generator=g)
        m = scale_img(mask, shape[-2:])
        out.update({f"comp.{tag}.src": np_(src), f"comp.{tag}.ref": np_(ref), f"comp.{tag}.m": np_(m),
                    f"comp.{tag}.out": np_((1 - m) * src + m * ref)})
    out["comp.mask"] = np_(mask)
    np.savez_compressed(OUT / "upfirdn2d_composite.npz", **out)


def gen_stylegan2():
    g = torch.Generator().manual_seed(300)
    out = {}
    for tag, up in [("plain", False), ("up", True)]:
        torch.manual_seed(3)
        mod = sg2.StyledConv(16, 32, 3, 24, upsample=up)
        randomize(mod, g)
        x = torch.randn(2, 16, 8, 8, generator=g)
        style = torch.randn(2, 24, generator=g)
        oh = 16 if up else 8
        noise = torch.randn(2, 1, oh, oh, generator=g)
        conv = mod.conv(x, style)
        y = mod(x, style, noise=noise)
        out.update({f"sc.{tag}.x": np_(x), f"sc.{tag}.style": np_(style), f"sc.{tag}.noise": np_(noise),
                    f"sc.{tag}.conv": np_(conv), f"sc.{tag}.out": np_(y)})
        out.update({f"sc.{tag}.sd.{k}": np_(v) for k, v in mod.state_dict().items()})
    # ModulatedConv2d(downsample=True) (model.py:265-272; not reachable from the scripts, part of the signature)
    torch.manual_seed(8)
    mod = sg2.ModulatedConv2d(16, 32, 3, 24, downsample=True)
    randomize(mod, g)
    x = torch.randn(2, 16, 12, 8, generator=g)
    style = torch.randn(2, 24, generator=g)
    out.update({"down.x": np_(x), "down.style": np_(style), "down.out": np_(mod(x, style))})
    out.update({f"down.sd.{k}": np_(v) for k, v in mod.state_dict().items()})
    torch.manual_seed(4)
    mod = sg2.ToRGB(16, 24)
    randomize(mod, g)
    x = torch.randn(2, 16, 8, 8, generator=g)
    style = torch.randn(2, 24, generator=g)
    skip = torch.randn(2, 3, 4, 4, generator=g)
    out.update({"rgb.x": np_(x), "rgb.style": np_(style), "rgb.skip": np_(skip), "rgb.out": np_(mod(x, style, skip)),
                "rgb.out_noskip": np_(mod(x, style))})
    out.update({f"rgb.sd.{k}": np_(v) for k, v in mod.state_dict().items()})
    np.savez_compressed(OUT / "stylegan2_layers.npz", **out)

    # Whole synthesis network at 32x32 (channels are 512 by construction, model.py:395-405): the 21 M parameters are not
    # stored — both sides rebuild them from the seed via our module mirror, and the reference loads them strict=True,
    # which also pins state_dict compatibility.
    my_gen = build_generator32()
    ref_gen = sg2.Generator(32, 512, 2)
    ref_gen.load_state_dict(my_gen.state_dict(), strict=True)
    latent = torch.randn(2, ref_gen.n_latent, 512, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        img, _ = ref_gen([latent], input_is_latent=True, randomize_noise=False)
    np.savez_compressed(OUT / "generator32.npz", latent=np_(latent), image=np_(img))


def gen_picnet():
    """BASELINE config 1: the reference's ReferenceFill (modules/model.py) with the README configuration, every parameter
    filled by name (golden_util.fill_by_name), get_z replaced by the distribution means, one forward at 256x256, batch 1.
    Stored: the 256x256 output, the 1024x1024 decoder output subsampled 8x, the attention input features' statistics
    and the reference's state_dict keys (the drop-in contract of modules/picnet.py)."""
    import types as _types
    from golden_util import fill_by_name, picnet_inputs, mean_z
    from modules.model import ReferenceFill
    enc = dict(type='pluralistic', ngf=32, z_nc=128, img_f=128, layers=5, norm='none', activation='LeakyReLU',
               init_type='orthogonal')
    dec = dict(ngf=32, z_nc=256, img_f=256, L=0, layers=5, norm='instance', activation='LeakyReLU', init_type='orthogonal')
    torch.manual_seed(0)
    model = ReferenceFill(None, dict(enc), dict(dec), use_att=True).eval()
    fill_by_name(model)
    model.decoder.get_z = _types.MethodType(mean_z, model.decoder)
    keys = sorted(model.state_dict().keys())
    src, ref, mask = picnet_inputs(1)
    with torch.no_grad():
        full = model(src, ref, mask, resize=False)
    # a second model instance: SpectralNorm's u/v advanced once per forward, so every forward starts from fresh state
    model2 = ReferenceFill(None, dict(enc), dict(dec), use_att=True).eval()
    fill_by_name(model2)
    model2.decoder.get_z = _types.MethodType(mean_z, model2.decoder)
    with torch.no_grad():
        out = model2(src, ref, mask)
    np.savez_compressed(OUT / "picnet_ref.npz", image=np_(out), full_sub8=np_(full[:, :, ::8, ::8]),
                        keys=np.array(keys), n_params=np.array(sum(p.numel() for p in model.parameters())))


def gen_picnet_blocks():
    """the reference's own ResBlockDecoder and Output (modules/pluralistic_model/base_function.py:308-398)
    with SpectralNorm, InstanceNorm2d(affine) and LeakyReLU(0.1) exactly as ResGenerator builds them (network.py:225-241),
    parameters filled by name, one eval forward on a seeded non-square input. Stored: input, every parameter (w_bar, u, v,
    biases, norm scales) BEFORE the forward, the block output, and Output applied to it."""
    import functools
    from golden_util import fill_by_name
    from modules.pluralistic_model import base_function as BF
    norm = functools.partial(nn.InstanceNorm2d, affine=True)
    act = nn.LeakyReLU(0.1)
    torch.manual_seed(0)
    blk = fill_by_name(BF.ResBlockDecoder(64, 32, 32, norm, act, True, False).eval(), seed=3)
    out = fill_by_name(BF.Output(32, 3, 3, None, act, True, False).eval(), seed=4)
    d = {}
    for tag, m in (("blk", blk), ("out", out)):
        for k, v in m.state_dict().items():   # convs are registered twice (conv1 and model.2, ...): keep the first name
            if ".module." in k and (k.startswith("model.") or k.startswith("shortcut.")):
                continue
            d[f"{tag}.{k}"] = np_(v.clone())
    x = torch.randn(2, 64, 12, 20, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        y = blk(x)
        img = out(y)
    d.update(x=np_(x), y=np_(y), img=np_(img))
    np.savez_compressed(OUT / "picnet_blocks.npz", **d)


def gen_ssim():
    """The reference's own SSIM (modules/evaluations/ssim.py) on two seeded image pairs: pins oracle.ssim."""
    from modules.evaluations.ssim import ssim as ref_ssim
    g = torch.Generator().manual_seed(77)
    a = torch.rand(2, 3, 40, 56, generator=g)
    b = (a + 0.05 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    np.savez_compressed(OUT / "ssim.npz", a=np_(a), b=np_(b), ssim=np_(ref_ssim(a, b)), ssim_same=np_(ref_ssim(a, a)),
                        ssim_per_image=np_(ref_ssim(a, b, size_average=False)))


def gen_loss_side():
    """f3: the reference's own GramMatrix / StyleLoss / contextual_loss (modules/pluralistic_model/external_function.py:180-192,
    231-274) on seeded VGG-like feature maps: pins oracle.gram_matrix / style_loss / contextual_loss."""
    from modules.pluralistic_model.external_function import GramMatrix, StyleLoss, contextual_loss
    g = torch.Generator().manual_seed(91)
    x = torch.relu(torch.randn(2, 64, 12, 8, generator=g))
    y = torch.relu(x + 0.5 * torch.randn(x.shape, generator=g))
    np.savez_compressed(OUT / "loss_side.npz", x=np_(x), y=np_(y), gram=np_(GramMatrix(x)), style=np_(StyleLoss(x, y)),
                        cx=np_(contextual_loss(x, y)), cx_h1=np_(contextual_loss(y, x, h=1.0)))


def gen_refpsp(size=256):
    """BASELINE config 3 (reduced output size for the fixture): the reference's pSp (modules/psp/psp.py) with
    GradualStyleEncoder(50, 'ir_se') + attention and its StyleGAN2 decoder, `load_weights` bypassed (no pretrained files
    offline), latent_avg = 0, every parameter filled by name, eval mode, randomize_noise=False, batch 1.
    Stored: codes [1, n_styles, 512], the face-pooled 256x256 image and the reference's state_dict keys."""
    from argparse import Namespace
    from golden_util import fill_by_name, refpsp_inputs
    _install_op_stub()
    import modules.psp.psp as ref_psp
    ref_psp.pSp.load_weights = lambda self: None
    opts = Namespace(output_size=size, encoder_type='GradualStyleEncoder', use_attention=1, train_decoder=0,
                     start_from_latent_avg=1, learn_in_w=0, pt_ckpt_path=None, stylegan_weights=None)
    torch.manual_seed(0)
    net = ref_psp.pSp(opts).eval()
    net.latent_avg = torch.zeros(opts.n_styles, 512)
    fill_by_name(net)
    keys = sorted(net.state_dict().keys())
    x, ref, mask = refpsp_inputs(1)
    with torch.no_grad():
        img, codes = net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False, return_latents=True)
    np.savez_compressed(OUT / f"refpsp{size}.npz", image=np_(img), codes=np_(codes), keys=np.array(keys),
                        n_params=np.array(sum(p.numel() for p in net.parameters())))


def gen_psp_encoder_pieces():
    """bottleneck_IR_SE / bottleneck_IR units, a GradualStyleBlock and _upsample_add of the reference's own encoder classes
    (modules/psp/encoders/helpers.py:77-119, psp_encoders.py:13-37,83-98), eval mode, randomised BatchNorm statistics."""
    from modules.psp.encoders import helpers, psp_encoders
    g = torch.Generator().manual_seed(300)
    out = {}

    def rand_bn(m):
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.copy_(0.5 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
                mod.weight.data.copy_(1 + 0.3 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.data.copy_(0.3 * torch.randn(mod.bias.shape, generator=g))

    for tag, cls, cin, depth, stride in [("se_s1", helpers.bottleneck_IR_SE, 32, 32, 1), ("se_s2", helpers.bottleneck_IR_SE, 32, 64, 2),
                                         ("se_pool", helpers.bottleneck_IR_SE, 32, 32, 2), ("ir_s1", helpers.bottleneck_IR, 16, 16, 1)]:
        torch.manual_seed(301)
        unit = cls(cin, depth, stride).eval()
        rand_bn(unit)
        x = torch.randn(2, cin, 12, 10, generator=g)
        with torch.no_grad():
            y = unit(x)
        out[f"{tag}.x"], out[f"{tag}.y"], out[f"{tag}.stride"] = np_(x), np_(y), np.array(stride)
        for k, v in unit.state_dict().items():
            if "num_batches" not in k:
                out[f"{tag}.sd.{k}"] = np_(v)
    torch.manual_seed(302)
    blk = psp_encoders.GradualStyleBlock(32, 32, 8).eval()
    x = torch.randn(3, 32, 8, 8, generator=g)
    with torch.no_grad():
        out["head.x"], out["head.y"] = np_(x), np_(blk(x))
    for k, v in blk.state_dict().items():
        out[f"head.sd.{k}"] = np_(v)
    a, b = torch.randn(2, 8, 5, 7, generator=g), torch.randn(2, 8, 10, 13, generator=g)
    out["fpn.x"], out["fpn.y"] = np_(a), np_(b)
    out["fpn.out"] = np_(psp_encoders.GradualStyleEncoder._upsample_add(None, a, b))
    np.savez_compressed(OUT / "psp_encoder.npz", **out)
    print("psp_encoder.npz", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] in ("picnet", "refpsp", "picnet_blocks", "ssim", "psp_encoder", "loss_side"):
        if sys.argv[1] == "psp_encoder":
            gen_psp_encoder_pieces()
        elif sys.argv[1] == "loss_side":
            gen_loss_side()
        elif sys.argv[1] == "picnet":
            gen_picnet()
        elif sys.argv[1] == "picnet_blocks":
            gen_picnet_blocks()
        elif sys.argv[1] == "ssim":
            gen_ssim()
        else:
            gen_refpsp()
        for f in sorted(OUT.glob("*.npz")):
            print(f.name, f.stat().st_size)
        sys.exit(0)
    gen_attention()
    gen_upfirdn2d_and_composite()
    gen_stylegan2()
    gen_picnet()
    gen_picnet_blocks()
    gen_ssim()
    gen_refpsp()
    gen_psp_encoder_pieces()
    gen_loss_side()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
