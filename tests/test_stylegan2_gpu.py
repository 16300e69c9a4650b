"""GPU parity of the StyleGAN2 decoder path (a3 ModulatedConv2d, a6 StyledConv / ToRGB / Generator) against the
CPU oracle. Tolerance: north_star's max|a-b|/max|b| <= 1e-3 for fp32 I/O (TF32 operands) and <= 2e-2 for bf16."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

MODES = [("fp32", torch.float32, 1e-3), ("bf16", torch.bfloat16, 2e-2)]


def _mods():
    from face_mask_inpaint_b200.modules import stylegan2 as SG
    return SG


def _randomize(mod, g):
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if name.endswith("noise.weight"):
                p.fill_(0.3)
            elif name.endswith("activate.bias") or name.endswith(".bias") and p.dim() == 4:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("modulation.bias"):
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g))


@pytest.mark.parametrize("cfg", [
    # (B, I, O, H, W, upsample)
    (2, 64, 64, 16, 16, False),
    (2, 64, 32, 16, 16, True),
    (3, 512, 512, 4, 4, False),     # conv1 of the generator (16 pixels per image, two N tiles)
    (2, 512, 512, 4, 4, True),      # first up layer: parity classes of 5x5 / 4x5 / 5x4 / 4x4
    (1, 128, 64, 40, 24, False),    # ragged tiles
    (1, 32, 32, 64, 64, True),      # 32 input channels = half a swizzle row (TMA zero fill along K)
    (1, 32, 32, 8, 256, False),     # halo mode (W >= 128): one 130-pixel box per kernel row, two tiles per row
    (2, 64, 64, 5, 130, False),     # halo mode, ragged second tile (2 valid pixels), top/bottom padding rows
    (1, 128, 128, 4, 128, False),   # halo mode, N tile 128, two K chunks
])
@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_modulated_conv2d(cfg, mode):
    SG = _mods()
    b, i, o, h, w, up = cfg
    _, dtype, tol = mode
    g = torch.Generator().manual_seed(0)
    mod = SG.ModulatedConv2d(i, o, 3, 512, upsample=up)
    _randomize(mod, g)
    x = torch.randn(b, i, h, w, generator=g).to(dtype).float()
    style = torch.randn(b, 512, generator=g)
    want = O.modulated_conv2d(x, style, mod.weight.detach(), mod.modulation.weight.detach(),
                              mod.modulation.bias.detach(), True, up)
    mod = mod.to(DEV)
    with torch.no_grad():
        got = mod(x.to(dtype).to(DEV), style.to(DEV))
    assert got.shape == want.shape and got.dtype == dtype
    err = rel_err(got, want)
    assert err <= tol, f"rel err {err:.3e}"


@pytest.mark.parametrize("cfg", [(2, 32, 64, 16, 16), (1, 64, 32, 12, 20)])
@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_modulated_conv2d_downsample(cfg, mode):
    """model.py:265-272 (blur pad (2,2), then stride-2 valid conv): served by the blur kernel + the stride-1 implicit GEMM."""
    SG = _mods()
    b, i, o, h, w = cfg
    _, dtype, tol = mode
    g = torch.Generator().manual_seed(2)
    mod = SG.ModulatedConv2d(i, o, 3, 512, downsample=True)
    _randomize(mod, g)
    x = torch.randn(b, i, h, w, generator=g).to(dtype).float()
    style = torch.randn(b, 512, generator=g)
    want = O.modulated_conv2d(x, style, mod.weight.detach(), mod.modulation.weight.detach(),
                              mod.modulation.bias.detach(), True, False, downsample=True)
    mod = mod.to(DEV)
    with torch.no_grad():
        got = mod(x.to(dtype).to(DEV), style.to(DEV))
    assert got.shape == want.shape == (b, o, h // 2, w // 2)
    err = rel_err(got, want)
    assert err <= tol, f"rel err {err:.3e}"


@pytest.mark.parametrize("up", [False, True])
@pytest.mark.parametrize("batched_noise", [False, True])
@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_styled_conv(up, batched_noise, mode):
    SG = _mods()
    _, dtype, tol = mode
    b, i, o, h, w = 2, 128, 64, 16, 16
    g = torch.Generator().manual_seed(1)
    mod = SG.StyledConv(i, o, 3, 512, upsample=up)
    _randomize(mod, g)
    x = torch.randn(b, i, h, w, generator=g).to(dtype).float()
    style = torch.randn(b, 512, generator=g)
    oh = 2 * h if up else h
    noise = torch.randn(b if batched_noise else 1, 1, oh, oh, generator=g)
    sd = {k: v.detach() for k, v in mod.state_dict().items()}
    want = O.styled_conv(x, style, sd['conv.weight'], sd['conv.modulation.weight'], sd['conv.modulation.bias'],
                         sd['noise.weight'], sd['activate.bias'], noise, upsample=up)
    mod = mod.to(DEV)
    with torch.no_grad():
        got = mod(x.to(dtype).to(DEV), style.to(DEV), noise=noise.to(DEV))
    err = rel_err(got, want)
    assert err <= tol, f"rel err {err:.3e}"


@pytest.mark.parametrize("cfg", [(2, 64, 32, 32, 40), (1, 128, 64, 24, 24), (2, 32, 32, 17, 33)])
@pytest.mark.parametrize("batched_noise", [False, True])
@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_styled_conv_upsample_narrow_layers(cfg, batched_noise, fused, mode, monkeypatch):
    """Up-sampling StyledConv with O <= 64: conv_transpose + blur as ONE merged-parity implicit GEMM with the combined 6x6
    weights (FMI_UPBLUR_FUSED=1, default) and as the two-pass path (=0), both against the oracle; ragged and odd extents."""
    SG = _mods()
    monkeypatch.setenv("FMI_UPBLUR_FUSED", fused)
    _, dtype, tol = mode
    b, i, o, h, w = cfg
    g = torch.Generator().manual_seed(11)
    mod = SG.StyledConv(i, o, 3, 512, upsample=True)
    _randomize(mod, g)
    x = torch.randn(b, i, h, w, generator=g).to(dtype).float()
    style = torch.randn(b, 512, generator=g)
    noise = torch.randn(b if batched_noise else 1, 1, 2 * h, 2 * w, generator=g)
    sd = {k: v.detach() for k, v in mod.state_dict().items()}
    want = O.styled_conv(x, style, sd['conv.weight'], sd['conv.modulation.weight'], sd['conv.modulation.bias'],
                         sd['noise.weight'], sd['activate.bias'], noise, upsample=True)
    mod = mod.to(DEV)
    n0 = _launches()
    with torch.no_grad():
        got = mod(x.to(dtype).to(DEV), style.to(DEV), noise=noise.to(DEV))
    used = _launches() - n0
    assert got.shape == want.shape == (b, o, 2 * h, 2 * w)
    err = rel_err(got, want)
    assert err <= tol, f"rel err {err:.3e}"
    # style modulation + weight prep + layout changes are common; the conv itself is 2 launches (combined weights + one GEMM)
    # fused, 5 (four parity-class GEMMs + blur) two-pass
    base = 4                                            # nchw_to_nhwc, style_mod, weight_prep, nhwc_to_nchw
    assert used == base + (2 if fused == "1" else 5), used


def _launches():
    from face_mask_inpaint_b200 import _lib
    return _lib.load().fmi_kernel_launch_count()


@pytest.mark.parametrize("with_skip", [False, True])
@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_to_rgb(with_skip, mode):
    SG = _mods()
    _, dtype, tol = mode
    b, i, h, w = 2, 64, 32, 32
    g = torch.Generator().manual_seed(2)
    mod = SG.ToRGB(i, 512)
    _randomize(mod, g)
    x = torch.randn(b, i, h, w, generator=g).to(dtype).float()
    style = torch.randn(b, 512, generator=g)
    skip = torch.randn(b, 3, h // 2, w // 2, generator=g) if with_skip else None
    sd = {k: v.detach() for k, v in mod.state_dict().items()}
    want = O.to_rgb(x, style, sd['conv.weight'], sd['conv.modulation.weight'], sd['conv.modulation.bias'], sd['bias'],
                    skip)
    mod = mod.to(DEV)
    with torch.no_grad():
        got = mod(x.to(dtype).to(DEV), style.to(DEV), skip.to(DEV) if with_skip else None)
    err = rel_err(got, want)
    assert err <= tol, f"rel err {err:.3e}"


@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_generator_synthesis(mode, monkeypatch):
    """Whole synthesis network (Generator.forward, input_is_latent=True, randomize_noise=False) at 64x64:
    9 StyledConv + 5 ToRGB chained in the native NHWC layout."""
    SG = _mods()
    name, dtype, tol = mode
    if name == "bf16":
        monkeypatch.setenv("FMI_PRECISION", "bf16")
    g = torch.Generator().manual_seed(3)
    torch.manual_seed(3)
    gen = SG.Generator(64, 512, 2)
    _randomize(gen, g)
    latent = torch.randn(2, gen.n_latent, 512, generator=g)
    sd = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    want = O.generator_synthesis(sd, latent)
    gen = gen.to(DEV)
    with torch.no_grad():
        got, _ = gen([latent.to(DEV)], input_is_latent=True, randomize_noise=False)
    assert got.shape == (2, 3, 64, 64)
    err = rel_err(got, want)
    assert err <= tol, f"rel err {err:.3e}"


@pytest.mark.parametrize("cfg", [(1, 32, 32, 8, 256), (2, 64, 64, 5, 130), (1, 128, 128, 4, 128)])
@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
def test_modulated_conv2d_halo_mode(cfg, mode, monkeypatch):
    """Opt-in halo mode of the implicit GEMM (FMI_MODCONV_HALO=1): one 130-pixel TMA box per kernel row, the three
    horizontal taps read row-shifted windows of it (tools/umma_probe.cu tests 9-12 establish that the MMA unit swizzles on
    address bits, so a descriptor may start at any 128-byte row of a SWIZZLE_128B tile)."""
    monkeypatch.setenv("FMI_MODCONV_HALO", "1")
    SG = _mods()
    b, i, o, h, w = cfg
    _, dtype, tol = mode
    g = torch.Generator().manual_seed(21)
    mod = SG.ModulatedConv2d(i, o, 3, 512)
    _randomize(mod, g)
    x = torch.randn(b, i, h, w, generator=g).to(dtype).float()
    style = torch.randn(b, 512, generator=g)
    want = O.modulated_conv2d(x, style, mod.weight.detach(), mod.modulation.weight.detach(),
                              mod.modulation.bias.detach(), True, False)
    mod = mod.to(DEV)
    with torch.no_grad():
        got = mod(x.to(dtype).to(DEV), style.to(DEV))
    assert rel_err(got, want) <= tol
