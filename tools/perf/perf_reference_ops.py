"""a4 / a5: this package's kernels next to the REFERENCE'S OWN CUDA kernels (oracle/_ref/*.so, built by oracle/build_ref.py from
modules/psp/stylegan2/op/*.cu) on the same B200 and the same tensors — NOT a pytest file:

    python tools/perf/perf_reference_ops.py > gpurun_out/perf_reference_ops.txt

Shapes are the largest live call sites of the 1024^2 generator at batch 8. GB/s = algorithmic bytes (read + write
once) / time; "of HBM" against the measured copy bandwidth in MEASURED_PEAKS.json (6456.8 GB/s)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from face_mask_inpaint_b200 import ops  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle import ref_ops as O  # noqa: E402

HBM = 6456.8
SQRT2 = 2 ** 0.5


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


def row(name, nbytes, t_ours, t_ref):
    print(f"{name:64s} ours {t_ours * 1e3:8.1f} us {nbytes / t_ours / 1e6:7.0f} GB/s ({nbytes / t_ours / 1e6 / HBM:4.2f} of HBM) | "
          f"reference kernel {t_ref * 1e3:8.1f} us {nbytes / t_ref / 1e6:7.0f} GB/s | {t_ref / t_ours:5.2f}x", flush=True)


def main():
    fused = build_ref.load_built("fmi_ref_fused")
    upf = build_ref.load_built("fmi_ref_upfirdn2d")
    for dtype in (torch.float32, torch.float16):
        esz = 4 if dtype == torch.float32 else 2
        tag = "fp32" if esz == 4 else "fp16"
        # a5 fused bias + leaky relu, forward and its grad = 1 backward
        for shape in ((8, 32, 1024, 1024), (8, 64, 512, 512), (8, 512, 64, 64)):
            x = torch.randn(shape, device="cuda", dtype=dtype)
            b = torch.randn(shape[1], device="cuda", dtype=dtype)
            empty = x.new_empty(0)
            n = x.numel() * esz
            row(f"fused_bias_act fwd {tag} {list(shape)}", 2 * n,
                time_cuda(lambda: ops.fused_bias_act(x, b, empty, 3, 0, 0.2, SQRT2)),
                time_cuda(lambda: fused.fused_bias_act(x, b, empty, 3, 0, 0.2, SQRT2)))
            y = ops.fused_bias_act(x, b, empty, 3, 0, 0.2, SQRT2)
            row(f"fused_bias_act bwd (grad=1) {tag} {list(shape)}", 3 * n,
                time_cuda(lambda: ops.fused_bias_act(x, empty, y, 3, 1, 0.2, SQRT2)),
                time_cuda(lambda: fused.fused_bias_act(x, empty, y, 3, 1, 0.2, SQRT2)))
            del x, y
        # a4 upfirdn2d at the live call sites
        k4 = (O.make_kernel([1, 3, 3, 1]) * 4).cuda()
        kk = k4 if dtype == torch.float32 else k4.to(dtype)
        for name, (c, h, w), (up, down, p0, p1) in (("blur after up-conv", (32, 1025, 1025), (1, 1, 1, 1)),
                                                    ("blur after up-conv", (64, 513, 513), (1, 1, 1, 1)),
                                                    ("blur backward", (32, 1024, 1024), (1, 1, 2, 2)),
                                                    ("RGB skip upsample", (3, 512, 512), (2, 1, 2, 1)),
                                                    ("upsample backward (down 2)", (3, 1024, 1024), (1, 2, 1, 1))):
            x = torch.randn(8 * c, h, w, 1, device="cuda", dtype=dtype)
            out = ops.upfirdn2d_op(x, k4, up, up, down, down, p0, p1, p0, p1)
            n = (x.numel() + out.numel()) * esz
            row(f"upfirdn2d {name} {tag} [8x{c},{h},{w}] up{up} down{down} pad({p0},{p1})", n,
                time_cuda(lambda: ops.upfirdn2d_op(x, k4, up, up, down, down, p0, p1, p0, p1)),
                time_cuda(lambda: upf.upfirdn2d(x, kk, up, up, down, down, p0, p1, p0, p1)))
            del x, out


if __name__ == "__main__":
    main()
