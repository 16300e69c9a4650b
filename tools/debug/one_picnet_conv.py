"""One fmi_conv3x3_nhwc configuration of the PICNet conv blocks, 3 launches (ncu captures): 
python tools/debug/one_picnet_conv.py B I O H W MODE [fp32|bf16|tf32x3]   (modes: csrc/conv_blocks.cu, fmi_conv3x3_nhwc)"""
import os
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
b, i, o, h, w, mode = (int(v) for v in sys.argv[1:7])
os.environ["FMI_PRECISION"] = sys.argv[7] if len(sys.argv) > 7 else "fp32"
from face_mask_inpaint_b200.modules import picnet_fast as PF  # noqa: E402

k = PF._Ctx(torch.device("cuda", 0))
tr = mode in (2, 3)
x = torch.randn(b, h, w, i, device="cuda").to(k.dt)
wt = torch.randn((i, o, 3, 3) if tr else (o, i, 3, 3), device="cuda") / (3 * i ** 0.5)
wp = k.weights([(wt, tr)], o, merged=mode == 3)
bias = torch.zeros(o, device="cuda")
oh, ow = (2 * h, 2 * w) if tr else (h, w)
y = k.empty(b, oh, ow, o)
for _ in range(3):
    k.conv(x.data_ptr(), i, wp, bias, y.data_ptr(), o, 0, None, 0, b, i, o, h, w, mode, 1, 0.1)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(20):
    k.conv(x.data_ptr(), i, wp, bias, y.data_ptr(), o, 0, None, 0, b, i, o, h, w, mode, 1, 0.1)
e[1].record()
torch.cuda.synchronize()
us = e[0].elapsed_time(e[1]) / 20 * 1e3
byt = (x.numel() + y.numel()) * x.element_size()
print(f"ok {float(y.float().abs().mean()):.4f}  {us:.1f} us  {byt / us / 1e3:.0f} GB/s algorithmic  {2.0 * b * oh * ow * o * i * (9 if not tr else 2.25) / us / 1e6:.1f} TFLOP/s")
k.finish()
