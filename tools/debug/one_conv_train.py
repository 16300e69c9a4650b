"""One training convolution of the PICNet decoder (64 -> 32 @512^2, batch 4) forward + backward through ops._ConvShared, 3 times,
for ncu captures of the weight-gradient / data-gradient GEMMs: python tools/debug/one_conv_train.py [B I O H W k]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import ops  # noqa: E402

b, i, o, h, w, k = (int(v) for v in sys.argv[1:7]) if len(sys.argv) > 6 else (4, 64, 32, 512, 512, 3)
x = torch.randn(b, i, h, w, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
wt = (torch.randn(o, i, k, k, device="cuda") / (k * i ** 0.5)).requires_grad_(True)
bias = torch.zeros(o, device="cuda", requires_grad=True)
gy = torch.randn(b, o, h, w, device="cuda").contiguous(memory_format=torch.channels_last)
for _ in range(3):
    y = ops._ConvShared.apply(x, wt, bias)
    y.backward(gy)
torch.cuda.synchronize()
print("ok", float(wt.grad.abs().mean()))
