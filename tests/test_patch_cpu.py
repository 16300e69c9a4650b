"""The drop-in installer against the real reference tree (build container only: /root/reference does not travel to
the GPU box). Checks that the patched classes are the ones the reference's assemblies instantiate and that
parameter/buffer names and shapes are unchanged (strict state_dict compatibility, psp.py:55)."""
import sys
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present")


def _fresh_reference_modules():
    for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
        del sys.modules[k]


def test_state_dict_layout_is_unchanged_by_the_patch():
    import types
    import torch.nn.functional as F
    from face_mask_inpaint_b200 import patch
    # 1) reference classes, with a throw-away stub op package so the stylegan2 module imports without its JIT build
    _fresh_reference_modules()
    sys.path.insert(0, str(REF))
    stub = types.ModuleType("modules.psp.stylegan2.op")
    stub.__path__ = []

    class _FLR(torch.nn.Module):
        def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
            super().__init__()
            self.bias = torch.nn.Parameter(torch.zeros(channel))

    stub.FusedLeakyReLU, stub.fused_leaky_relu, stub.upfirdn2d = _FLR, None, None
    sys.modules["modules.psp.stylegan2.op"] = stub
    from modules.example_guided_att import ExampleGuidedAttention as RefEGA
    from modules.pluralistic_model.base_function import Auto_Attn as RefAA
    from modules.psp.stylegan2.model import Generator as RefGen
    want = {
        "ega": {k: tuple(v.shape) for k, v in RefEGA(64, 32).state_dict().items()},
        "aa": {k: tuple(v.shape) for k, v in RefAA(64, None).state_dict().items()},
        "gen": {k: tuple(v.shape) for k, v in RefGen(32, 512, 2).state_dict().items()},
    }
    # 2) patched
    _fresh_reference_modules()
    patch.uninstall_flag_for_tests()
    patch.install(str(REF))
    import modules.model as ref_model
    import modules.pluralistic_model.base_function as bf
    import modules.psp.stylegan2.model as sg
    from face_mask_inpaint_b200.modules import attention as my_att, stylegan2 as my_sg
    assert ref_model.ExampleGuidedAttention is my_att.ExampleGuidedAttention
    assert bf.Auto_Attn is my_att.Auto_Attn
    assert sg.Generator is my_sg.Generator and sg.ModulatedConv2d is my_sg.ModulatedConv2d
    got = {
        "ega": {k: tuple(v.shape) for k, v in ref_model.ExampleGuidedAttention(64, 32).state_dict().items()},
        "aa": {k: tuple(v.shape) for k, v in bf.Auto_Attn(64, None).state_dict().items()},
        "gen": {k: tuple(v.shape) for k, v in sg.Generator(32, 512, 2).state_dict().items()},
    }
    assert got == want
    # 3) the reference's own assemblies pick up the drop-ins (PICNet decoder's attn1, network.py:243-245)
    from modules.pluralistic_model import network
    g = network.define_g(ngf=8, z_nc=16, img_f=32, L=0, layers=3, norm='instance', activation='LeakyReLU',
                         init_type='orthogonal')
    assert isinstance(g.attn1, my_att.Auto_Attn)
    # 4) f1: the reference's ResGenerator.forward is the installer's (kernel path in CUDA inference); its blocks are recognised
    #    by the kernel path (the reference's own SpectralNorm / ResBlockDecoder / Output classes, duck-typed), and off the GPU
    #    or under autograd it is the reference's forward, pooled on request
    from face_mask_inpaint_b200.modules import picnet_fast
    assert network.ResGenerator.forward.__name__ == "res_generator_forward"
    g32 = network.define_g(ngf=32, z_nc=64, img_f=128, L=0, layers=3, norm='instance', activation='LeakyReLU',
                           init_type='orthogonal', use_attn=False).eval()
    assert isinstance(g32.decoder0.conv1, picnet_fast.SpectralNorm) and not isinstance(g32.decoder0.conv1.module,
                                                                                     picnet_fast.SpectralNorm)
    assert picnet_fast._block_layout(g32.decoder0) is not None
    x = torch.randn(1, 128, 4, 4)
    assert not picnet_fast.supported(g32, x)                      # CPU tensor: never the kernel path
    with torch.no_grad():
        full = g32(x)
    assert full.shape == (1, 3, 32, 32)
    g32b = network.define_g(ngf=32, z_nc=64, img_f=128, L=0, layers=3, norm='instance', activation='LeakyReLU',
                            init_type='orthogonal', use_attn=False).eval()
    g32b.load_state_dict(g32.state_dict())                        # incl. the SpectralNorm u / v advanced by the call above
    with torch.no_grad():
        pooled = g32b(x, pool_to=(8, 8))
        want_pooled = F.adaptive_avg_pool2d(g32(x), (8, 8))
    assert torch.allclose(pooled, want_pooled, atol=1e-6)
    _fresh_reference_modules()
    patch.uninstall_flag_for_tests()


def test_standalone_auto_attn_keeps_reference_names():
    """Without the reference importable, Auto_Attn builds its own `model` block with the same parameter names."""
    from face_mask_inpaint_b200.modules.attention import Auto_Attn
    saved = Auto_Attn.resblock_factory
    Auto_Attn.resblock_factory = None
    try:
        keys = set(Auto_Attn(16, None).state_dict().keys())
    finally:
        Auto_Attn.resblock_factory = saved
    for k in ["query_conv.weight", "query_conv.bias", "gamma", "alpha", "model.conv1.module.weight_bar",
              "model.conv1.module.weight_u", "model.conv1.module.weight_v", "model.conv1.module.bias",
              "model.bypass.module.weight_bar", "model.model.1.module.weight_bar", "model.shortcut.0.module.bias"]:
        assert k in keys, k
