"""Pins the CPU oracle (oracle/ref_ops.py) against outputs of the reference's OWN modules, recorded by
tests/golden/make_golden.py from /root/reference (the reference ships no tests or golden vectors of its own).
CPU only; fp32 vs fp32 with the same ATen kernels underneath, so agreement is to rounding (1e-6)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as O

GOLD = Path(__file__).resolve().parent / "golden"


def load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(GOLD / name).items()}


def test_example_guided_attention_matches_reference():
    g = load("attention.npz")
    for tag, has_oc in [("ega_plain", False), ("ega_outconv", True)]:
        got = O.example_guided_attention(g[f"{tag}.mask"], g[f"{tag}.src"], g[f"{tag}.ref"], g[f"{tag}.conv_w"],
                                         g.get(f"{tag}.oc_w") if has_oc else None,
                                         g.get(f"{tag}.oc_b") if has_oc else None)
        assert rel_err(got, g[f"{tag}.out"]) <= 1e-6


def test_auto_attn_matches_reference():
    g = load("attention.npz")
    out, _, attn = O.auto_attn(g["auto.x"], g["auto.q_w"], g["auto.q_b"], g["auto.gamma"], return_attention=True)
    assert rel_err(out, g["auto.out"]) <= 1e-6
    assert rel_err(attn, g["auto.attn"]) <= 1e-6
    out, ctx, _ = O.auto_attn(g["auto.x"], g["auto.q_w"], g["auto.q_b"], g["auto.gamma"], g["auto.pre"], g["auto.mask"],
                              g["auto.alpha"])
    assert rel_err(torch.cat([out, ctx], 1), g["auto.cat"]) <= 1e-6


@pytest.mark.parametrize("tag", ["blur", "blur_bwd", "up", "down", "k3", "odd", "crop"])
def test_upfirdn2d_matches_reference_native(tag):
    g = load("upfirdn2d_composite.npz")
    up, down, p0, p1 = [int(v) for v in g[f"ufd.{tag}.cfg"]]
    got = O.upfirdn2d(g[f"ufd.{tag}.x"], g[f"ufd.{tag}.k"], up=up, down=down, pad=(p0, p1))
    assert got.shape == g[f"ufd.{tag}.y"].shape
    assert rel_err(got, g[f"ufd.{tag}.y"]) <= 1e-6


def test_upfirdn2d_minor_dim_matches_reference_native():
    g = load("upfirdn2d_composite.npz")
    got = O.upfirdn2d_native(g["ufd.minor.x"], g["ufd.minor.k"], 2, 1, 1, 2, 1, 2, 0, 1)
    assert rel_err(got, g["ufd.minor.y"]) <= 1e-6


def test_upfirdn2d_backward_restatement_is_the_adjoint():
    """op/upfirdn2d.py:17-57: the gradient op equals autograd through the forward restatement."""
    g = torch.Generator().manual_seed(0)
    for up, down, pad, taps in [(1, 1, (1, 1), [1, 3, 3, 1]), (2, 1, (2, 1), [1, 3, 3, 1]), (1, 2, (1, 1), [1, 3, 3, 1])]:
        x = torch.randn(2, 3, 8, 8, generator=g, dtype=torch.float64, requires_grad=True)
        k = (O.make_kernel(taps) * up ** 2).double()
        y = O.upfirdn2d(x, k, up=up, down=down, pad=pad)
        go = torch.randn(y.shape, generator=g, dtype=torch.float64)
        y.backward(go)
        assert rel_err(O.upfirdn2d_backward(go, k, up, down, pad, x.shape), x.grad) <= 1e-12


@pytest.mark.parametrize("tag", ["c32", "c16", "c7x9"])
def test_composite_matches_reference(tag):
    g = load("upfirdn2d_composite.npz")
    h, w = g[f"comp.{tag}.src"].shape[-2:]
    assert rel_err(O.scale_img(g["comp.mask"], (h, w)), g[f"comp.{tag}.m"]) <= 1e-6
    assert rel_err(O.composite(g[f"comp.{tag}.src"], g[f"comp.{tag}.ref"], g["comp.mask"]), g[f"comp.{tag}.out"]) <= 1e-6


@pytest.mark.parametrize("tag,up", [("plain", False), ("up", True)])
def test_styled_conv_matches_reference(tag, up):
    g = load("stylegan2_layers.npz")
    sd = {k[len(f"sc.{tag}.sd."):]: v for k, v in g.items() if k.startswith(f"sc.{tag}.sd.")}
    conv = O.modulated_conv2d(g[f"sc.{tag}.x"], g[f"sc.{tag}.style"], sd["conv.weight"], sd["conv.modulation.weight"],
                              sd["conv.modulation.bias"], True, up)
    assert rel_err(conv, g[f"sc.{tag}.conv"]) <= 1e-5
    out = O.styled_conv(g[f"sc.{tag}.x"], g[f"sc.{tag}.style"], sd["conv.weight"], sd["conv.modulation.weight"],
                        sd["conv.modulation.bias"], sd["noise.weight"], sd["activate.bias"], g[f"sc.{tag}.noise"],
                        upsample=up)
    assert rel_err(out, g[f"sc.{tag}.out"]) <= 1e-5


def test_modulated_conv_downsample_matches_reference():
    g = load("stylegan2_layers.npz")
    sd = {k[len("down.sd."):]: v for k, v in g.items() if k.startswith("down.sd.")}
    out = O.modulated_conv2d(g["down.x"], g["down.style"], sd["weight"], sd["modulation.weight"], sd["modulation.bias"], True,
                             False, downsample=True)
    assert out.shape == g["down.out"].shape == (2, 32, 6, 4)
    assert rel_err(out, g["down.out"]) <= 1e-5


def test_to_rgb_matches_reference():
    g = load("stylegan2_layers.npz")
    sd = {k[len("rgb.sd."):]: v for k, v in g.items() if k.startswith("rgb.sd.")}
    args = (sd["conv.weight"], sd["conv.modulation.weight"], sd["conv.modulation.bias"], sd["bias"])
    assert rel_err(O.to_rgb(g["rgb.x"], g["rgb.style"], *args, skip=g["rgb.skip"]), g["rgb.out"]) <= 1e-5
    assert rel_err(O.to_rgb(g["rgb.x"], g["rgb.style"], *args), g["rgb.out_noskip"]) <= 1e-5


def test_generator_synthesis_matches_reference():
    """Whole StyleGAN2 synthesis at 32x32; parameters are rebuilt from the seed (tests/golden_util.py) — the
    reference loaded the same state_dict with strict=True when the golden was made."""
    from golden_util import build_generator32
    g = load("generator32.npz")
    torch.set_num_threads(max(torch.get_num_threads(), 4))
    gen = build_generator32()
    sd = {k: v.detach() for k, v in gen.state_dict().items()}
    got = O.generator_synthesis(sd, g["latent"])
    assert rel_err(got, g["image"]) <= 1e-5


def test_fused_bias_act_semantics():
    """op/fused_bias_act_kernel.cu:18-49 has no CPU-runnable reference: check the restatement's switch table."""
    x = torch.tensor([[-2.0, 3.0], [0.5, -0.25]])
    b = torch.tensor([1.0, -1.0])
    r = torch.tensor([[1.0, -1.0], [-1.0, 1.0]])
    assert torch.allclose(O.fused_bias_act(x, b, None, 3, 0, 0.2, 2.0), torch.tensor([[-0.4, 4.0], [3.0, -0.5]]))
    assert torch.allclose(O.fused_bias_act(x, None, r, 3, 1, 0.2, 2.0), torch.tensor([[-4.0, 1.2], [0.2, -0.5]]))
    assert torch.allclose(O.fused_bias_act(x, b, r, 3, 2, 0.2, 2.0), torch.zeros(2, 2))
    assert torch.allclose(O.fused_bias_act(x, b, None, 1, 0, 0.2, 2.0), (x + b) * 2)
    gi, gb = O.fused_leaky_relu_backward(torch.ones(2, 2, 1, 1), torch.tensor([[[[1.0]], [[-1.0]]], [[[-1.0]], [[1.0]]]]))
    assert torch.allclose(gb, torch.tensor([1.2 * 2 ** 0.5, 1.2 * 2 ** 0.5]))


def _sn(g, prefix):
    w, _, _ = O.spectral_norm_weight(g[f"{prefix}.module.weight_bar"], g[f"{prefix}.module.weight_u"],
                                     g[f"{prefix}.module.weight_v"])
    return w, g[f"{prefix}.module.bias"]


def test_picnet_decoder_blocks_match_reference():
    """f1: SpectralNorm + ResBlockDecoder + Output restatements against the reference's own classes
    (base_function.py:308-398, external_function.py:44-57)."""
    g = load("picnet_blocks.npz")
    w1, b1 = _sn(g, "blk.conv1")
    w2, b2 = _sn(g, "blk.conv2")
    ws, bs = _sn(g, "blk.bypass")
    y = O.res_block_decoder(g["x"], w1, b1, w2, b2, ws, bs, (g["blk.model.0.weight"], g["blk.model.0.bias"]),
                            (g["blk.model.3.weight"], g["blk.model.3.bias"]), slope=0.1)
    assert rel_err(y, g["y"]) <= 1e-6
    wo, bo = _sn(g, "out.conv1")
    assert rel_err(O.output_block(y, wo, bo, slope=0.1), g["img"]) <= 1e-6


def test_ssim_matches_reference():
    """oracle.ssim against the reference's own modules/evaluations/ssim.py (the metric behind north_star's SSIM parity)."""
    g = load("ssim.npz")
    assert abs(float(O.ssim(g["a"], g["b"])) - float(g["ssim"])) <= 1e-6
    assert abs(float(O.ssim(g["a"], g["a"])) - float(g["ssim_same"])) <= 1e-6 and float(g["ssim_same"]) > 0.999999
    assert rel_err(O.ssim(g["a"], g["b"], size_average=False), g["ssim_per_image"]) <= 1e-6


# ---- f2: pSp encoder pieces, recorded from the reference's own bottleneck_IR(_SE), GradualStyleBlock and _upsample_add
def _sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


@pytest.mark.parametrize("tag", ["se_s1", "se_s2", "se_pool", "ir_s1"])
def test_ir_se_unit_matches_reference(tag):
    g = load("psp_encoder.npz")
    got = O.bottleneck_ir_se(g[f"{tag}.x"], _sub(g, f"{tag}.sd."), int(g[f"{tag}.stride"]))
    assert got.shape == g[f"{tag}.y"].shape and rel_err(got, g[f"{tag}.y"]) <= 1e-6


def test_gradual_style_block_and_fpn_add_match_reference():
    g = load("psp_encoder.npz")
    assert rel_err(O.gradual_style_block(g["head.x"], _sub(g, "head.sd."), 3), g["head.y"]) <= 1e-6
    assert rel_err(O.upsample_add(g["fpn.x"], g["fpn.y"]), g["fpn.out"]) <= 1e-6


@pytest.mark.parametrize("tag", ["se_s1", "se_s2", "se_pool", "ir_s1"])
def test_folded_unit_of_the_kernel_path_matches_the_golden(tag, monkeypatch):
    """The BatchNorm folding of modules/psp_fast.py (border-class bias, per-output scale) applied to the GOLDEN unit's weights and
    evaluated densely on CPU reproduces the reference's own output: ties the host-side algebra to the reference, not only to
    this package's mirror (tests/test_psp_fast_cpu.py)."""
    from face_mask_inpaint_b200.modules import psp as P
    from face_mask_inpaint_b200.modules import psp_fast as PF
    from test_psp_fast_cpu import _conv_from_taps, _planes_conv
    monkeypatch.setattr(PF, "_operand", lambda w, mma: w.contiguous())
    g = load("psp_encoder.npz")
    sd = _sub(g, f"{tag}.sd.")
    stride = int(g[f"{tag}.stride"])
    cin, depth = sd["res_layer.1.weight"].shape[1], sd["res_layer.1.weight"].shape[0]
    unit = P._IRUnit(cin, depth, stride, "res_layer.5.fc1.weight" in sd).eval()
    unit.load_state_dict(sd, strict=False)
    u = PF._prep_unit(unit, 0)
    x = g[f"{tag}.x"]
    with torch.no_grad():
        a1 = _conv_from_taps(x, u.w1, u.b1)
        a1 = torch.where(a1 > 0, a1, a1 * u.slope.view(1, -1, 1, 1))
        r = _planes_conv(a1, u.w2, u.b2) if stride == 2 else _conv_from_taps(a1, u.w2, u.b2)
        xs = x[:, :, ::stride, ::stride]
        sc = xs if u.ws is None else _conv_from_taps(xs, u.ws, u.bs)
        if u.se1 is not None:
            gate = torch.sigmoid(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(r.mean(dim=(2, 3)), u.se1)), u.se2))
            r = r * gate.view(*gate.shape, 1, 1)
        got = r + sc
    assert rel_err(got, g[f"{tag}.y"]) <= 2e-5


def test_loss_side_ops_match_reference():
    """f3: oracle.gram_matrix / style_loss / contextual_loss against the reference's own functions
    (modules/pluralistic_model/external_function.py:180-192, 231-274; golden from tests/golden/make_golden.py loss_side)."""
    g = load("loss_side.npz")
    assert rel_err(O.gram_matrix(g["x"]), g["gram"]) <= 1e-6
    assert rel_err(O.style_loss(g["x"], g["y"]), g["style"]) <= 1e-6
    assert rel_err(O.contextual_loss(g["x"], g["y"]), g["cx"]) <= 1e-6
    assert rel_err(O.contextual_loss(g["y"], g["x"], h=1.0), g["cx_h1"]) <= 1e-6
