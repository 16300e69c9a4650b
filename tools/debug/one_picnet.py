"""One PICNet-ref 256^2 forward (batch 8 by default) after two warm-up forwards, for ncu launch lists:
    python tools/debug/one_picnet.py [batch] [fp32|bf16]"""
import os
import sys
import types
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent / "tests"))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
os.environ["FMI_PRECISION"] = sys.argv[2] if len(sys.argv) > 2 else "fp32"
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref  # noqa: E402
from golden_util import fill_by_name, mean_z, picnet_inputs  # noqa: E402

net = fill_by_name(build_picnet_ref()).eval().cuda()
net.decoder.get_z = types.MethodType(mean_z, net.decoder)
src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
with torch.no_grad():
    for _ in range(2):
        net(src, ref, mask)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("measured")
    out = net(src, ref, mask)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print(out.shape, float(out.abs().mean()))
