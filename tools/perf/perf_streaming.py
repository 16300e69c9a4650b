"""HBM-bound kernels against the measured copy bandwidth — NOT a pytest file; run on the GPU box:

    python tools/perf/perf_streaming.py > gpurun_out/perf_streaming.txt

Each case is one C-ABI call at a StyleGAN2-1024 / PICNet shape; `GB/s` = algorithmic bytes (every input
element read once, every output element written once) / CUDA-event time; `frac` = GB/s over MEASURED_PEAKS.json hbm_gbs.
Working sets are >= 256 MB (larger than the 126 MB L2) so no flush is needed between iterations.
"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from face_mask_inpaint_b200 import ops  # noqa: E402
from face_mask_inpaint_b200.modules import stylegan2 as SG  # noqa: E402


def peak_gbs():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6456.8, "fallback (round-1 measured copy bandwidth)"


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
    peak, src = peak_gbs()
    print(f"# HBM peak {peak:.1f} GB/s ({src})")
    print(f"{'case':62s} {'ms':>8s} {'GB/s':>8s} {'frac':>6s}")
    torch.manual_seed(0)
    k4 = SG.make_kernel([1, 3, 3, 1]).to(dev)

    only = os.environ.get("PERF_ONLY")  # substring filter (ncu captures)

    def report(name, nbytes, fn):
        if only and only not in name:
            return
        t = time_cuda(fn, *((1, 1) if os.environ.get("PERF_QUICK") else ()))
        gbs = nbytes / t * 1e-6
        print(f"{name:62s} {t:8.3f} {gbs:8.1f} {gbs / peak:6.2f}", flush=True)

    with torch.no_grad():
        for dt in (torch.float32, torch.bfloat16):
            es = torch.empty(0, dtype=dt).element_size()
            tag = "fp32" if dt == torch.float32 else "bf16"
            # a5: fused bias + leaky relu, the 1024^2 layer of the generator (B=8, 32 channels)
            x = torch.randn(8, 32, 1024, 1024, device=dev, dtype=dt)
            b = torch.randn(32, device=dev)
            report(f"fused_leaky_relu fwd [8,32,1024,1024] {tag}", 2 * x.numel() * es, lambda: ops.fused_leaky_relu(x, b))
            y = ops.fused_leaky_relu(x, b)
            g = torch.randn_like(x)
            report(f"fused_leaky_relu bwd (+bias grad) [8,32,1024,1024] {tag}", 3 * x.numel() * es,
                   lambda: ops.FusedLeakyReLUFunctionBackward.apply(g, y, 0.2, 2 ** 0.5, 32))
            del y, g
            # a4: Blur after the up-conv (mode 1): [C,2H+1,2H+1] -> [C,2H,2H]
            xm = torch.randn(8, 32, 1025, 1025, device=dev, dtype=dt)
            report(f"upfirdn2d blur pad(1,1) [8,32,1025,1025]->[.,1024,1024] {tag}",
                   (xm.numel() + 8 * 32 * 1024 * 1024) * es, lambda: ops.upfirdn2d(xm, k4 * 4, pad=(1, 1)))
            # its backward: pad (2,2), [.,1024,1024] -> [.,1025,1025]
            report(f"upfirdn2d blur-bwd pad(2,2) [8,32,1024,1024]->[.,1025,1025] {tag}",
                   (xm.numel() + x.numel()) * es, lambda: ops.upfirdn2d(x, k4 * 4, pad=(2, 2)))
            del xm
            # a4: down=2 (backward of the RGB-skip upsample at channel width; Downsample), mode 5
            report(f"upfirdn2d down=2 pad(1,1) [8,32,1024,1024]->[.,512,512] {tag}",
                   (x.numel() + x.numel() // 4) * es, lambda: ops.upfirdn2d(x, k4, down=2, pad=(1, 1)))
            del x
            # a4: Upsample of the RGB skip (mode 3) — 3 planes only, so use a wide batch to exceed L2
            xs = torch.randn(64, 3, 512, 512, device=dev, dtype=dt)
            report(f"upfirdn2d up=2 pad(2,1) [64,3,512,512]->[.,1024,1024] {tag}", 5 * xs.numel() * es,
                   lambda: ops.upfirdn2d(xs, k4 * 4, up=2, pad=(2, 1)))
            del xs
            # a7: compositing with the mask bilinear-sampled in the kernel (pSp c1 at width 128 / PICNet-like large map)
            src_f = torch.randn(32, 128, 128, 128, device=dev, dtype=dt)
            ref_f = torch.randn_like(src_f)
            mask = (torch.rand(32, 1, 256, 256, device=dev) > 0.5).float()
            report(f"composite [32,128,128,128] mask 256^2 {tag}", 3 * src_f.numel() * es + mask.numel() * 4,
                   lambda: ops.composite(src_f, ref_f, mask))
            del src_f, ref_f


if __name__ == "__main__":
    main()
