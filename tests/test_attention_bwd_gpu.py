"""GPU parity of the attention BACKWARD (fmi_attn_bwd through the autograd Functions of the module mirrors) against
autograd through the CPU oracle on the same seeded inputs. Same tolerance metric as the forward:
max|a-b|/max|b| <= 1e-3 (fp32 I/O, TF32 operands) / 2e-2 (bf16) — gradients of every input and parameter."""
import pytest
import torch

from conftest import rel_err
from oracle import ref_ops as O
from test_attention_gpu import _mask, _scaled_query_weight

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("case", [(2, 128, 32, 32, None), (1, 64, 16, 16, 48), (2, 256, 16, 16, 256)])
@pytest.mark.parametrize("logit_std", [1.0, 4.0])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_example_guided_attention_backward(case, logit_std, dtype):
    from face_mask_inpaint_b200.modules import ExampleGuidedAttention
    n, c, h, w, oc = case
    g = torch.Generator().manual_seed(10)
    src = torch.randn(n, c, h, w, generator=g).to(dtype).float()
    ref = torch.randn(n, c, h, w, generator=g).to(dtype).float()
    mask = _mask(n, h, w, g)
    mod = ExampleGuidedAttention(c, oc)
    with torch.no_grad():
        mod.conv.weight.copy_(_scaled_query_weight(c, c // 4, src, logit_std, g))
        if oc is not None:
            mod.out_conv.weight.copy_(torch.randn(mod.out_conv.weight.shape, generator=g) / (2 * c) ** 0.5)
            mod.out_conv.bias.copy_(torch.randn(oc, generator=g))
    go = torch.randn(n, oc or 2 * c, h, w, generator=g).to(dtype).float()
    # oracle + autograd on CPU
    ps = [src.clone().requires_grad_(True), ref.clone().requires_grad_(True), mod.conv.weight.detach().clone().requires_grad_(True)]
    ocw = mod.out_conv.weight.detach().clone().requires_grad_(True) if oc else None
    ocb = mod.out_conv.bias.detach().clone().requires_grad_(True) if oc else None
    O.example_guided_attention(mask, ps[0], ps[1], ps[2], ocw, ocb).backward(go)
    # ours
    mod = mod.to(DEV)
    sd = src.to(dtype).to(DEV).requires_grad_(True)
    rd = ref.to(dtype).to(DEV).requires_grad_(True)
    mod(mask.to(DEV), sd, rd).backward(go.to(dtype).to(DEV))
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert rel_err(sd.grad, ps[0].grad) <= tol, "d src"
    assert rel_err(rd.grad, ps[1].grad) <= tol, "d ref"
    # parameter gradients are sums over all pixels with heavy cancellation: their error relative to max|grad| is the
    # accumulated per-pixel error (<= 1e-3 each, see dq above) over a small total — bounded separately
    assert rel_err(mod.conv.weight.grad, ps[2].grad) <= 30 * tol, "d Wq"
    if oc:
        assert rel_err(mod.out_conv.weight.grad, ocw.grad) <= tol and rel_err(mod.out_conv.bias.grad, ocb.grad) <= tol


@pytest.mark.parametrize("case", [(2, 128, 32, 32), (1, 256, 32, 32), (2, 64, 16, 16)])
@pytest.mark.parametrize("logit_std", [1.0, 4.0])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_auto_attn_backward(case, logit_std, dtype):
    from face_mask_inpaint_b200.modules import Auto_Attn
    n, c, h, w = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, c, h, w, generator=g).to(dtype).float()
    mod = Auto_Attn(c, None)
    with torch.no_grad():
        mod.query_conv.weight.copy_(_scaled_query_weight(c, c // 4, x, logit_std, g))
        mod.query_conv.bias.copy_(0.05 * torch.randn(c // 4, generator=g))
        mod.gamma.fill_(0.7)
    go = torch.randn(n, c, h, w, generator=g).to(dtype).float()
    xr = x.clone().requires_grad_(True)
    wr = mod.query_conv.weight.detach().clone().requires_grad_(True)
    br = mod.query_conv.bias.detach().clone().requires_grad_(True)
    gr = mod.gamma.detach().clone().requires_grad_(True)
    O.auto_attn(xr, wr, br, gr)[0].backward(go)
    mod = mod.to(DEV)
    xd = x.to(dtype).to(DEV).requires_grad_(True)
    out, _ = mod(xd)
    out.backward(go.to(dtype).to(DEV))
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert rel_err(xd.grad, xr.grad) <= tol, "dx"
    # the attention-only part of dx (without the identity path of `+ x`)
    if dtype == torch.float32:
        assert rel_err(xd.grad.cpu() - go, xr.grad - go) <= 3e-3, "dx (attention part)"
    assert rel_err(mod.query_conv.weight.grad, wr.grad) <= 30 * tol, "dWq"
    assert rel_err(mod.query_conv.bias.grad, br.grad) <= 30 * tol, "dbq"
    assert rel_err(mod.gamma.grad, gr.grad) <= 3 * tol, "dgamma"


def test_auto_attn_pre_branch_backward():
    """Both value groups, masked blend, alpha: gradients through cat[out, context_flow] (base_function.py:439-445)."""
    from face_mask_inpaint_b200.modules.attention import _AutoAttnFunction
    n, c, h, w = 2, 64, 16, 16
    g = torch.Generator().manual_seed(12)
    x = torch.randn(n, c, h, w, generator=g)
    pre = torch.randn(n, c, h, w, generator=g)
    mask = _mask(n, h, w, g)
    wq = _scaled_query_weight(c, c // 4, x, 1.0, g)
    bq = 0.05 * torch.randn(c // 4, generator=g)
    go = torch.randn(n, 2 * c, h, w, generator=g)
    ps = [t.clone().requires_grad_(True) for t in (x, wq, bq, torch.tensor([0.7]), pre, torch.tensor([1.3]))]
    out, ctx_flow, _ = O.auto_attn(ps[0], ps[1], ps[2], ps[3], ps[4], mask, ps[5])
    torch.cat([out, ctx_flow], 1).backward(go)
    qs = [t.detach().to(DEV).requires_grad_(True) for t in (x, wq, bq, torch.tensor([0.7]), pre, torch.tensor([1.3]))]
    cat, _ = _AutoAttnFunction.apply(qs[0], qs[1], qs[2], qs[3], qs[4], mask.to(DEV), qs[5])
    cat.backward(go.to(DEV))
    for name, a, b in zip(["dx", "dWq", "dbq", "dgamma", "dpre", "dalpha"], qs, ps):
        assert rel_err(a.grad, b.grad) <= 2e-3, name
