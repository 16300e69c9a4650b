"""a4 / a5 against the REFERENCE'S OWN CUDA OPS. oracle/build_ref.py compiles the reference's four source files
(modules/psp/stylegan2/op/{fused_bias_act,upfirdn2d}{.cpp,_kernel.cu}) where they lie under /root/reference into
oracle/_ref/*.so (build container; the shared objects travel to the GPU box). Here the reference's pybind ops
`fused.fused_bias_act` / `upfirdn2d.upfirdn2d` run on the same inputs as this package's drop-in ops (same signatures,
face_mask_inpaint_b200.ops.fused_bias_act / upfirdn2d_op) and as the CPU oracle — which pins the oracle's restatement of
fused_bias_act (its only reference implementation is that CUDA kernel). fp32 and fp16 (the reference has no bf16 path)."""
import sys
from pathlib import Path

import pytest
import torch

from conftest import rel_err

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import build_ref  # noqa: E402
from oracle import ref_ops as O  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(len(build_ref.built()) < 2, reason="oracle/_ref/*.so not built (python oracle/build_ref.py)")]
SQRT2 = 2 ** 0.5


@pytest.fixture(scope="module")
def ref_fused():
    return build_ref.load_built("fmi_ref_fused")


@pytest.fixture(scope="module")
def ref_upfirdn():
    return build_ref.load_built("fmi_ref_upfirdn2d")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize("shape", [(2, 32, 16, 16), (1, 512, 4, 4), (3, 7, 5, 9), (4, 24)])
def test_fused_bias_act_forward_and_backward_vs_reference_kernel(ref_fused, dtype, tol, shape):
    from face_mask_inpaint_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).cuda().to(dtype)
    b = torch.randn(shape[1], generator=g).cuda().to(dtype)
    empty = x.new_empty(0)
    want = ref_fused.fused_bias_act(x, b, empty, 3, 0, 0.2, SQRT2)                 # fused_act.py:54
    got = ops.fused_bias_act(x, b, empty, 3, 0, 0.2, SQRT2)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert rel_err(got, want) <= tol, rel_err(got, want)
    # the CPU oracle's restatement (fp32 arithmetic on the same rounded inputs)
    orc = O.fused_bias_act(x.float().cpu(), b.float().cpu(), None, 3, 0, 0.2, SQRT2)
    assert rel_err(want.float().cpu(), orc) <= tol
    # backward: grad_input = fused_bias_act(grad, empty, out, act 3, grad 1) (fused_act.py:27-29), bias grad = its sum
    gy = torch.randn(shape, generator=g).cuda().to(dtype)
    want_gx = ref_fused.fused_bias_act(gy, empty, want, 3, 1, 0.2, SQRT2)
    got_gx = ops.fused_bias_act(gy, empty, want, 3, 1, 0.2, SQRT2)
    assert rel_err(got_gx, want_gx) <= tol
    orc_gx = O.fused_bias_act(gy.float().cpu(), None, want.float().cpu(), 3, 1, 0.2, SQRT2)
    assert rel_err(want_gx.float().cpu(), orc_gx) <= tol
    if dtype == torch.float32:       # whole autograd path of the module-level op
        xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ops.fused_leaky_relu(xr, br, 0.2, SQRT2).backward(gy)
        dims = [0] + list(range(2, len(shape)))
        assert rel_err(xr.grad, want_gx) <= 1e-6 and rel_err(br.grad, want_gx.sum(dims)) <= 1e-5


# (up, down, pad0, pad1, kernel scale): the live call sites of the scripts + their backward configurations
SITES = [(1, 1, 1, 1, 4.0),     # Blur after the up-sampling modulated conv
         (1, 1, 2, 2, 4.0),     # its backward
         (2, 1, 2, 1, 4.0),     # Upsample of the RGB skip
         (1, 2, 1, 1, 4.0),     # its backward
         (1, 2, 2, 2, 1.0)]     # Blur + stride of the down-sampling branch (model.py:217-223)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize("site", SITES)
@pytest.mark.parametrize("shape", [(2, 16, 33, 33), (1, 3, 64, 48), (2, 8, 9, 17)])
def test_upfirdn2d_vs_reference_kernel(ref_upfirdn, dtype, tol, site, shape):
    from face_mask_inpaint_b200 import ops
    up, down, p0, p1, ks = site
    n, c, h, w = shape
    g = torch.Generator().manual_seed(h * w + up + down)
    x = torch.randn(shape, generator=g).cuda().to(dtype)
    k = (O.make_kernel([1, 3, 3, 1]) * ks).cuda()
    xin = x.reshape(-1, h, w, 1)                                                    # upfirdn2d.py:95
    want = ref_upfirdn.upfirdn2d(xin, k.to(dtype) if dtype != torch.float32 else k, up, up, down, down, p0, p1, p0, p1)
    got = ops.upfirdn2d_op(xin, k, up, up, down, down, p0, p1, p0, p1)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert rel_err(got, want) <= tol, rel_err(got, want)
    orc = O.upfirdn2d(x.float().cpu(), k.float().cpu(), up=up, down=down, pad=(p0, p1))
    assert rel_err(want.float().reshape(n, c, want.shape[1], want.shape[2]).cpu(), orc) <= tol
