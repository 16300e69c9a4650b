"""One fmi_conv_nhwc configuration, 3 launches, for ncu source-level captures: python tools/debug/one_conv.py B I O HW [bf16|tf32]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
b, i, o, hw = (int(v) for v in sys.argv[1:5])
os.environ["FMI_PRECISION"] = sys.argv[5] if len(sys.argv) > 5 else "bf16"
from face_mask_inpaint_b200.modules import psp_fast as PF  # noqa: E402

k = PF._Ctx(torch.device("cuda", 0))
x = PF._operand(torch.randn(b, hw, hw, i, device=k.dev), k.mma)
w = PF._operand(torch.randn(9, o, i, device=k.dev) / (3 * i ** 0.5), k.mma)
bias = torch.zeros(o, device=k.dev)
y = k.empty(b, hw, hw, o)
for _ in range(3):
    k.conv(x, i, w, bias, y, b, i, o, hw, hw)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
