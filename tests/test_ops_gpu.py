"""GPU parity of the HBM-bound ops (a4 upfirdn2d, a5 fused bias-act, a7 compositing) through the C ABI
against the CPU oracle on the same seeded inputs. fp32 results: <= 1e-5 relative (summation order only);
bf16: <= 2e-2 (north_star tolerance), in practice one bf16 ulp."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from face_mask_inpaint_b200 import ops
    return ops


@pytest.mark.parametrize("shape", [(2, 8, 4, 4), (3, 5, 7, 9), (2, 16, 64, 64), (1, 32, 128, 128), (4, 12), (1, 3, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_fused_leaky_relu_fwd_bwd(shape, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    go = torch.randn(*shape, generator=g)
    xr = x.to(dtype).float().clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = O.fused_leaky_relu(xr, br)
    yr.backward(go.to(dtype).float())
    xd = x.to(dtype).to(DEV).detach().requires_grad_(True)
    bd = b.to(DEV).requires_grad_(True)
    yd = ops.fused_leaky_relu(xd, bd)
    yd.backward(go.to(dtype).to(DEV))
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert yd.dtype == dtype and yd.shape == xd.shape
    assert rel_err(yd, yr) <= tol
    assert rel_err(xd.grad, xr.grad) <= tol
    assert rel_err(bd.grad, br.grad) <= (1e-5 if dtype == torch.float32 else 2e-2)


def test_fused_bias_act_codes_and_empty():
    ops = _ops()
    x = torch.randn(2, 6, 5, 5)
    b = torch.randn(6)
    r = torch.randn(2, 6, 5, 5)
    for act, grad in [(1, 0), (1, 1), (1, 2), (3, 0), (3, 1), (3, 2)]:
        want = O.fused_bias_act(x, b, r, act, grad, 0.2, 1.5)
        got = ops.fused_bias_act(x.to(DEV), b.to(DEV), r.to(DEV), act, grad, 0.2, 1.5)
        assert rel_err(got, want) <= 1e-6 or want.abs().max() == 0
    # empty bias / ref tensors mean "absent" (fused_bias_act_kernel.cu:62-63)
    e = torch.empty(0, device=DEV)
    got = ops.fused_bias_act(x.to(DEV), e, e, 3, 0, 0.2, 2 ** 0.5)
    assert rel_err(got, O.fused_bias_act(x, None, None, 3, 0, 0.2, 2 ** 0.5)) <= 1e-6
    assert ops.fused_bias_act(torch.empty(0, 4, device=DEV), torch.zeros(4, device=DEV), e, 3, 0, 0.2, 1.0).numel() == 0


UFD_CASES = [
    # (C, H, W, kernel taps, up, down, pad)
    (4, 9, 9, [1, 3, 3, 1], 1, 1, (1, 1)),       # Blur after up-modconv   (mode 1)
    (4, 8, 8, [1, 3, 3, 1], 1, 1, (2, 2)),       # its backward
    (3, 4, 4, [1, 3, 3, 1], 2, 1, (2, 1)),       # Upsample of the RGB skip (mode 3)
    (3, 8, 8, [1, 3, 3, 1], 1, 2, (1, 1)),       # its backward / Downsample (mode 5)
    (2, 7, 5, [1, 2, 1], 1, 1, (1, 1)),          # 3x3 taps (mode 2)
    (2, 6, 6, [1, 1], 2, 1, (1, 0)),             # 2x2 taps up (mode 4)
    (2, 6, 6, [1, 1], 1, 2, (0, 0)),             # 2x2 taps down (mode 6)
    (2, 5, 6, [1, 4, 6, 4, 1], 3, 2, (3, 2)),    # outside every reference mode
    (2, 10, 10, [1, 3, 3, 1], 1, 1, (-1, -2)),   # negative pad = crop
    (5, 65, 131, [1, 3, 3, 1], 1, 1, (1, 1)),    # ragged tile edges
    (2, 300, 300, [1, 3, 3, 1], 1, 1, (2, 2)),   # several tiles
    (3, 150, 70, [1, 3, 3, 1], 2, 1, (2, 1)),    # polyphase up=2: several tiles, ragged edges
    (2, 33, 65, [1, 3, 3, 1], 2, 1, (1, 2)),     # up=2 with odd leading pad (the other tap phase)
    (2, 20, 20, [1, 3, 3, 1], 2, 1, (3, 0)),
    (2, 40, 40, [1, 3, 3, 1], 2, 1, (-1, 1)),    # up=2 with a crop
    (3, 300, 140, [1, 3, 3, 1], 1, 2, (1, 1)),   # down=2: several tiles
    (2, 67, 131, [1, 3, 3, 1], 1, 2, (2, 1)),    # down=2, odd extents
    (2, 64, 64, [1, 2, 1], 1, 2, (0, 1)),        # down=2 with 3 taps
]


@pytest.mark.parametrize("case", UFD_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upfirdn2d_fwd_bwd(case, dtype):
    ops = _ops()
    C, H, W, taps, up, down, pad = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, C, H, W, generator=g).to(dtype).float()
    k = O.make_kernel(taps) * (up ** 2)
    if len(taps) == 5:
        k = k + 0.01 * torch.arange(25.).view(5, 5)  # asymmetric: exercises the tap flip
    xr = x.clone().requires_grad_(True)
    yr = O.upfirdn2d(xr, k, up=up, down=down, pad=pad)
    go = torch.randn(yr.shape, generator=g).to(dtype).float()
    yr.backward(go)
    xd = x.to(dtype).to(DEV).detach().requires_grad_(True)
    yd = ops.upfirdn2d(xd, k.to(DEV), up=up, down=down, pad=pad)
    assert yd.shape == yr.shape
    yd.backward(go.to(dtype).to(DEV))
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_err(yd, yr) <= tol
    assert rel_err(xd.grad, xr.grad) <= tol
    # explicit backward restatement (op/upfirdn2d.py:17-57)
    gb = O.upfirdn2d_backward(go, k, up, down, pad, x.shape)
    assert rel_err(gb, xr.grad) <= 1e-5


def test_upfirdn2d_minor_dim_and_errors():
    ops = _ops()
    x = torch.randn(3, 6, 7, 2)
    k = O.make_kernel([1, 3, 3, 1])
    want = O.upfirdn2d_native(x, k, 2, 1, 1, 2, 1, 2, 0, 1)
    got = ops.upfirdn2d_op(x.to(DEV), k.to(DEV), 2, 1, 1, 2, 1, 2, 0, 1)
    assert got.shape == want.shape and rel_err(got, want) <= 1e-5
    with pytest.raises(RuntimeError):
        ops.upfirdn2d(torch.randn(1, 1, 2, 2, device=DEV), torch.ones(4, 4, device=DEV))  # empty extent


@pytest.mark.parametrize("shape", [(4, 128, 32, 32), (2, 128, 64, 64), (2, 512, 16, 16), (3, 5, 7, 9)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_composite_fwd_bwd(shape, dtype):
    ops = _ops()
    n, c, h, w = shape
    g = torch.Generator().manual_seed(2)
    src = torch.randn(*shape, generator=g).to(dtype).float()
    ref = torch.randn(*shape, generator=g).to(dtype).float()
    mask = (torch.rand(n, 1, 256, 256, generator=g) < 0.3).float()
    mask[:, :, 128:230, 50:206] = 1.0
    sr, rr = src.clone().requires_grad_(True), ref.clone().requires_grad_(True)
    want = O.composite(sr, rr, mask)
    go = torch.randn(*shape, generator=g).to(dtype).float()
    want.backward(go)
    sd = src.to(dtype).to(DEV).detach().requires_grad_(True)
    rd = ref.to(dtype).to(DEV).detach().requires_grad_(True)
    got = ops.composite(sd, rd, mask.to(DEV))
    got.backward(go.to(dtype).to(DEV))
    tol = 2e-6 if dtype == torch.float32 else 2e-2
    assert rel_err(got, want) <= tol
    assert rel_err(sd.grad, sr.grad) <= tol and rel_err(rd.grad, rr.grad) <= tol
    m_got = ops.scale_img(mask.to(DEV), (h, w))
    assert rel_err(m_got, O.scale_img(mask, (h, w))) <= 2e-6
