"""One RefpSp 1024^2 forward (batch 8 by default, bf16 operands) after two warm-up forwards, for ncu launch lists:
    python tools/debug/one_refpsp.py [batch] [bf16|fp32]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
os.environ["FMI_PRECISION"] = sys.argv[2] if len(sys.argv) > 2 else "bf16"
from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts  # noqa: E402

torch.manual_seed(11)
net = pSp(refpsp_opts(output_size=1024)).eval().cuda()
g = torch.Generator().manual_seed(0)
x = (torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1).cuda()
ref = (torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1).cuda()
mask = torch.zeros(batch, 256, 256).cuda()
mask[:, 128:230, 50:206] = 1
with torch.no_grad():
    for _ in range(2):
        net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measured")
    out = net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print(out.shape, float(out.abs().mean()))
