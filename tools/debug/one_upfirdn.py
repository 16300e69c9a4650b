"""The four stand-alone upfirdn2d blur launches of perf_streaming.py (forward pad (1,1) on [8,32,1025,1025], backward pad (2,2) on
[8,32,1024,1024], fp32 and bf16), 3 launches each, for ncu captures: python tools/debug/one_upfirdn.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import ops  # noqa: E402

k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
k4 = (k1[:, None] * k1[None, :] / k1.sum() ** 2).cuda()
for dt in (torch.float32, torch.bfloat16):
    xm = torch.randn(8, 32, 1025, 1025, device="cuda", dtype=dt)
    x = torch.randn(8, 32, 1024, 1024, device="cuda", dtype=dt)
    for _ in range(3):
        y = ops.upfirdn2d(xm, k4 * 4, pad=(1, 1))
    for _ in range(3):
        y = ops.upfirdn2d(x, k4 * 4, pad=(2, 2))
    torch.cuda.synchronize()
    del xm, x, y
print("ok")
