"""Micro-benchmark of fmi_conv_nhwc (the implicit-GEMM kernel) over batch sizes: separates the per-launch overhead from the per-tile
cost.  python tools/perf/perf_conv_nhwc.py [bf16|tf32]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
os.environ["FMI_PRECISION"] = sys.argv[1] if len(sys.argv) > 1 else "bf16"
from face_mask_inpaint_b200.modules import psp_fast as PF  # noqa: E402

k = PF._Ctx(torch.device("cuda", 0))
print(f"mode {os.environ['FMI_PRECISION']}; columns: B I O HxW tiles | us/launch | TFLOP/s | us per wave of 148 tiles")
for (i, o, hw) in [(256, 256, 32), (128, 128, 64), (64, 64, 128), (512, 512, 16), (64, 64, 256), (32, 32, 512)]:
    for b in (2, 4, 8, 16, 32, 64):
        if b * hw * hw * max(i, o) * 4 > (3 << 30):
            continue
        x = PF._operand(torch.randn(b, hw, hw, i, device=k.dev), k.mma)
        w = PF._operand(torch.randn(9, o, i, device=k.dev) / (3 * i ** 0.5), k.mma)
        bias = torch.zeros(o, device=k.dev)
        y = k.empty(b, hw, hw, o)
        for _ in range(3):
            k.conv(x, i, w, bias, y, b, i, o, hw, hw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            k.conv(x, i, w, bias, y, b, i, o, hw, hw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        tiles = b * hw * hw // 128 * max(1, o // 256)
        fl = 2.0 * b * hw * hw * 9 * i * o
        print(f"B={b:3d} I={i:3d} O={o:3d} {hw:3d}^2 tiles={tiles:6d} | {us:8.1f} | {fl / us / 1e6:7.1f} | {us / max(1.0, tiles / 148):7.1f}")
