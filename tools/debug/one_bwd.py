"""One Auto_Attn forward+backward at 128x128 (for an ncu launch list). GPU box only."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200.modules import Auto_Attn
dev = "cuda"
torch.manual_seed(0)
c, hw = 256, 128
x = torch.randn(1, c, hw, hw, device=dev)
mod = Auto_Attn(c, None).to(dev)
with torch.no_grad():
    mod.query_conv.weight.mul_(0.5); mod.gamma.fill_(0.7)
go = torch.randn(1, c, hw, hw, device=dev)
for _ in range(2):
    xi = x.detach().requires_grad_(True)
    mod(xi)[0].backward(go)
torch.cuda.synchronize()
print("ok")
