"""Host logic of the training path (ops.run_block_sequential & friends, graphs.release_autograd_state) on CPU tensors: with no CUDA
tensor in sight every helper must fall through to the reference's own PyTorch formulation — same modules, same results — and
never touch the library (there is no CPU kernel path)."""
import torch
from torch import nn

from face_mask_inpaint_b200 import graphs, ops
from face_mask_inpaint_b200.modules.picnet import Output, ResBlock, ResBlockDecoder
from face_mask_inpaint_b200.modules.picnet_blocks import SpectralNorm


def test_support_predicates_refuse_cpu_tensors():
    x = torch.randn(2, 32, 16, 16, requires_grad=True)
    assert not ops.conv_train_supported(nn.Conv2d(32, 32, 3, 1, 1), x)
    assert not ops.conv_train_supported(nn.ConvTranspose2d(32, 32, 3, 2, 1, 1), x)
    assert not ops.norm_act_supported(nn.InstanceNorm2d(32, affine=True), nn.LeakyReLU(0.1), x)
    assert not ops.act_round_supported(nn.LeakyReLU(0.1), x)
    assert not ops.act_reflect_pad_supported(nn.LeakyReLU(0.1), nn.ReflectionPad2d(1), x)
    assert not ops.bmm_nt_supported(torch.randn(2, 8, 16), torch.randn(2, 8, 16))


def test_block_runner_is_the_sequential_on_cpu():
    torch.manual_seed(0)
    x = torch.randn(2, 8, 12, 12)
    seq = nn.Sequential(nn.InstanceNorm2d(8, affine=True), nn.LeakyReLU(0.1), nn.Conv2d(8, 8, 3, 1, 1), nn.AvgPool2d(2, 2),
                        nn.LeakyReLU(0.1), nn.ReflectionPad2d(1), nn.Conv2d(8, 3, 3), nn.Tanh())
    assert torch.equal(ops.run_block_sequential(seq, x), seq(x))
    pool = nn.AvgPool2d(kernel_size=2, stride=2)
    a, b = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    assert torch.equal(ops.pool_sum(pool, a, b), pool(a) + pool(b))
    assert torch.equal(ops.avg_pool2(pool, a), pool(a))


def test_mirror_blocks_unchanged_on_cpu():
    """The block forwards go through run_block_sequential: on CPU they must equal the plain Sequential composition (the reference's
    formulation, base_function.py:262-268, 361-364, 395-398)."""
    torch.manual_seed(1)
    x = torch.randn(1, 8, 8, 8)
    blk = ResBlock(8, 8, 8, norm_layer=None, nonlinearity=nn.LeakyReLU(0.1), sample_type='down', use_spect=False)
    assert torch.allclose(blk(x), blk.pool(blk.model(x)) + blk.pool(blk.shortcut(x)))
    dec = ResBlockDecoder(8, 8, 8, norm_layer=None, nonlinearity=nn.LeakyReLU(0.1), use_spect=False)
    assert torch.allclose(dec(x), dec.model(x) + dec.shortcut(x))
    out = Output(8, 3, 3, None, nn.LeakyReLU(0.1), False)
    assert torch.allclose(out(x), out.model(x))


def test_release_autograd_state_detaches_spectral_norm_weight():
    """SpectralNorm leaves its normalised weight (a non-leaf tensor with a grad_fn) on the wrapped module between forwards
    (external_function.py:57); graphs.release_autograd_state detaches it so that a capture does not inherit last iteration's graph."""
    sn = SpectralNorm(nn.Conv2d(4, 4, 3, 1, 1))
    sn(torch.randn(1, 4, 6, 6))
    w = sn.module.__dict__.get("weight")
    assert w is not None and w.grad_fn is not None
    graphs.release_autograd_state(nn.Sequential(sn))
    w2 = sn.module.__dict__["weight"]
    assert w2.grad_fn is None and torch.equal(w2, w.detach())
    y = sn(torch.randn(1, 4, 6, 6))          # the next forward rebuilds it
    assert y.requires_grad and sn.module.__dict__["weight"].grad_fn is not None


def test_nhwc_view_and_cache_rules():
    x = torch.randn(2, 4, 3, 5)
    v = ops._as_nhwc(x.contiguous(memory_format=torch.channels_last))
    assert v.shape == (2, 3, 5, 4) and v.is_contiguous()
    leaf = torch.randn(2, 4, 3, 5)
    a, b = ops._as_nhwc(leaf), ops._as_nhwc(leaf)
    assert a.data_ptr() != b.data_ptr()                 # a persistent leaf (a model input, a static graph buffer) is converted every time
    inter = (torch.randn(2, 4, 3, 5, requires_grad=True) * 2.0)
    c, d = ops._as_nhwc(inter), ops._as_nhwc(inter)
    assert c.data_ptr() == d.data_ptr()                 # an intermediate of ONE iteration is converted once
    g1, g2 = ops._as_nhwc(leaf, fresh=True), ops._as_nhwc(leaf, fresh=True)
    assert g1.data_ptr() == g2.data_ptr()
    leaf.add_(1.0)                                      # in-place change: the remembered copy is stale
    assert ops._as_nhwc(leaf, fresh=True).data_ptr() != g1.data_ptr()
