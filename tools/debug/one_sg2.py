"""One StyleGAN2-1024 forward (bf16 operands, batch 8) for ncu captures: python tools/debug/one_sg2.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
os.environ.setdefault("FMI_PRECISION", "bf16")
from face_mask_inpaint_b200.modules import stylegan2 as SG  # noqa: E402

torch.manual_seed(0)
gen = SG.Generator(1024, 512, 8).cuda().eval()
latent = torch.randn(8, gen.n_latent, 512, device="cuda")
with torch.no_grad():
    img = gen([latent], input_is_latent=True, randomize_noise=False)[0]
torch.cuda.synchronize()
print(img.shape, float(img.abs().mean()))
