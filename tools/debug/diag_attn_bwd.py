"""Debug: print gradient errors of the attention backward by component. GPU box only."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent / "tests"))
from oracle import ref_ops as O
from test_attention_gpu import _mask, _scaled_query_weight
from face_mask_inpaint_b200.modules import ExampleGuidedAttention, Auto_Attn
from face_mask_inpaint_b200 import ops
DEV = "cuda"
def rel(a, b): return ((a.detach().double().cpu() - b.detach().double().cpu()).abs().max() / b.detach().double().abs().max()).item()
for std in (0.3, 1.0, 4.0):
    n, c, h, w = 2, 128, 32, 32
    g = torch.Generator().manual_seed(10)
    src = torch.randn(n, c, h, w, generator=g); ref = torch.randn(n, c, h, w, generator=g); mask = _mask(n, h, w, g)
    wq = _scaled_query_weight(c, c // 4, src, std, g)
    go = torch.randn(n, 2 * c, h, w, generator=g)
    ps = [src.double().requires_grad_(True), ref.double().requires_grad_(True), wq.double().requires_grad_(True)]
    out64 = O.example_guided_attention(mask.double(), *ps)
    out64.backward(go.double())
    mod = ExampleGuidedAttention(c).to(DEV)
    with torch.no_grad(): mod.conv.weight.copy_(wq)
    sd = src.to(DEV).requires_grad_(True); rd = ref.to(DEV).requires_grad_(True)
    out = mod(mask.to(DEV), sd, rd); out.backward(go.to(DEV))
    # pieces: call the op directly
    _, lse, _, o_saved = ops.attention_forward(sd.detach(), mod.conv.weight.detach(), None, sd.detach(), rd.detach(), mask=mask.to(DEV), b0=0.0, masked1=True, order=(1, 0), need_lse=True, save_o=True)
    dq, dv0, dv1, _, _ = ops.attention_backward(sd.detach(), mod.conv.weight.detach(), None, sd.detach(), rd.detach(), mask.to(DEV), None, 0.0, False, None, 0.0, True, o_saved, lse, go.to(DEV), order=(1, 0))
    # fp64 pieces
    s64, r64, w64 = src.double(), ref.double(), wq.double()
    q = torch.nn.functional.conv2d(s64, w64).flatten(2).requires_grad_(True)
    vs = s64.flatten(2).clone().requires_grad_(True); vr = r64.flatten(2).clone().requires_grad_(True)
    P = torch.softmax(q.transpose(1, 2) @ q, -1)
    m = mask.double().flatten(2)
    outp = torch.cat([(1 - m) * (vr @ P.transpose(1, 2)) + m * vr, vs @ P.transpose(1, 2)], 1)
    outp.backward(go.double().flatten(2))
    lse64 = torch.logsumexp(q.transpose(1, 2) @ q, -1)
    print(f"std {std}: fwd out {rel(out, out64):.2e} lse {rel(lse, lse64):.2e} | dq {rel(dq.flatten(2), q.grad):.2e} dv0 {rel(dv0.flatten(2), vs.grad):.2e} dv1 {rel(dv1.flatten(2), vr.grad):.2e} | d src {rel(sd.grad, ps[0].grad):.2e} d ref {rel(rd.grad, ps[1].grad):.2e} dWq {rel(mod.conv.weight.grad, ps[2].grad):.2e}")
