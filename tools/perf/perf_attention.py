"""Attention microbenchmark (BASELINE config 2) — NOT a pytest file; run on the GPU box:

    python tools/perf/perf_attention.py > gpurun_out/perf_attention.txt

Times fmi_attn_fwd (prologue kernels + main kernel) with CUDA events and, beside it, the reference's PyTorch
formulation (oracle functions executed on the same GPU: cuBLAS fp32 bmm + ATen softmax, the composition the
unmodified reference runs on a B200). Lives under tests/ because it imports the oracle.
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import _lib, ops  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


def time_cuda(fn, warmup=3, iters=10):
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
    print(f"{'case':34s} {'mma':5s} {'ms':>9s} {'img/s':>10s} {'TFLOP/s':>9s} {'torch ms':>9s} {'speedup':>8s} {'relerr':>9s}")
    for kind in ("auto", "ega"):
        for (n, hw) in [(32, 32), (8, 64), (32, 64), (1, 128), (4, 128), (8, 128)]:
            c, d, s = 256, 64, hw * hw
            x = torch.randn(n, c, hw, hw, device=dev)
            ref = torch.randn(n, c, hw, hw, device=dev)
            wq = torch.randn(d, c, 1, 1, device=dev) / c ** 0.5
            q = torch.nn.functional.conv2d(x[:1], wq).flatten(2)
            e = (q.transpose(1, 2)[:, :1024] @ q[:, :, :1024])
            wq = wq * (1.0 / e.std()) ** 0.5
            mask = torch.rand(n, 1, hw, hw, device=dev)
            gamma = torch.tensor([0.7], device=dev)
            cv = c if kind == "auto" else 2 * c
            flops = n * (2.0 * s * s * d * (cv // 256) + 2.0 * s * s * cv)  # QK is recomputed per 256-ch slice
            algo_flops = n * (2.0 * s * s * d + 2.0 * s * s * cv)
            want = None
            t_ref = float("nan")
            if n * s * s * 4 <= 12 << 30:
                if kind == "auto":
                    fn_ref = lambda: O.auto_attn(x, wq, None, gamma)[0]
                else:
                    fn_ref = lambda: O.example_guided_attention(mask, x, ref, wq)
                want = fn_ref()
                t_ref = time_cuda(fn_ref, 2, 5)
            for dtype, mma, name in [(torch.float32, _lib.MMA_TF32, "tf32"), (torch.float32, _lib.MMA_BF16, "bf16*"),
                                     (torch.bfloat16, _lib.MMA_BF16, "bf16")]:
                xd, rd = x.to(dtype), ref.to(dtype)
                if kind == "auto":
                    fn = lambda: ops.attention_forward(xd, wq, None, xd, None, a0=gamma, b0=1.0, mma=mma)[0]
                else:
                    fn = lambda: ops.attention_forward(xd, wq, None, xd, rd, mask=mask, masked1=True, order=(1, 0), mma=mma)[0]
                got = fn()
                err = float("nan")
                if want is not None:
                    err = ((got.float() - want).abs().max() / want.abs().max()).item()
                t = time_cuda(fn)
                print(f"{kind + f' C=256 {hw}x{hw} N={n}':34s} {name:5s} {t:9.3f} {n / t * 1e3:10.1f} "
                      f"{algo_flops / t / 1e9:9.1f} {t_ref:9.3f} {t_ref / t:8.1f} {err:9.2e}", flush=True)
            del want


if __name__ == "__main__":
    main()
