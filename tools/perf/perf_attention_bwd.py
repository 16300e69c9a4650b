"""Attention forward+backward timing (training path of BASELINE configs 4/5) — run on the GPU box.
Ours (fmi_attn_fwd + fmi_attn_bwd via autograd) vs the reference formulation with PyTorch autograd on the same GPU."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200.modules import Auto_Attn, ExampleGuidedAttention  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


def time_cuda(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = "cuda"
    torch.manual_seed(0)
    print(f"{'case':40s} {'ours fwd+bwd ms':>16s} {'torch fwd+bwd ms':>17s} {'speedup':>8s} {'dx relerr':>10s}")
    for kind, n, c, hw in [("auto", 4, 256, 128), ("auto", 1, 256, 128), ("auto", 4, 128, 32), ("ega", 4, 128, 32),
                           ("ega", 8, 256, 64), ("auto", 8, 256, 64)]:
        d, s = c // 4, hw * hw
        x = torch.randn(n, c, hw, hw, device=dev)
        ref = torch.randn(n, c, hw, hw, device=dev)
        mask = torch.rand(n, 1, hw, hw, device=dev)
        wq = torch.randn(d, c, 1, 1, device=dev) / c ** 0.5
        q = torch.nn.functional.conv2d(x[:1, :, :32, :32], wq).flatten(2)
        wq = wq * (1.0 / (q.transpose(1, 2) @ q).std()) ** 0.5
        go = torch.randn(n, c if kind == "auto" else 2 * c, hw, hw, device=dev)
        if kind == "auto":
            mod = Auto_Attn(c, None).to(dev)
            with torch.no_grad():
                mod.query_conv.weight.copy_(wq)
                mod.gamma.fill_(0.7)
            params = [mod.query_conv.weight, mod.query_conv.bias, mod.gamma]

            def ours():
                xi = x.detach().requires_grad_(True)
                mod(xi)[0].backward(go)
                return xi.grad

            def torch_ref():
                xi = x.detach().requires_grad_(True)
                O.auto_attn(xi, mod.query_conv.weight, mod.query_conv.bias, mod.gamma)[0].backward(go)
                return xi.grad
        else:
            mod = ExampleGuidedAttention(c).to(dev)
            with torch.no_grad():
                mod.conv.weight.copy_(wq)
            params = [mod.conv.weight]

            def ours():
                xi = x.detach().requires_grad_(True)
                mod(mask, xi, ref).backward(go)
                return xi.grad

            def torch_ref():
                xi = x.detach().requires_grad_(True)
                O.example_guided_attention(mask, xi, ref, mod.conv.weight).backward(go)
                return xi.grad
        g_ours = ours()
        t = time_cuda(ours)
        err, t_ref = float("nan"), float("nan")
        if n * s * s * 4 * 4 <= 40 << 30:
            for p in params:
                p.grad = None
            g_ref = torch_ref()
            err = ((g_ours - g_ref).abs().max() / g_ref.abs().max()).item()
            t_ref = time_cuda(torch_ref, 1, 3)
        print(f"{kind + f' C={c} {hw}x{hw} N={n} fp32':40s} {t:16.3f} {t_ref:17.3f} {t_ref / t:8.1f} {err:10.2e}", flush=True)


if __name__ == "__main__":
    main()
