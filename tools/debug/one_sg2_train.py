"""One StyleGAN2-1024 forward+backward (bf16 operands, batch 2) for ncu launch lists: python tools/debug/one_sg2_train.py"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
os.environ.setdefault("FMI_PRECISION", "bf16")
from face_mask_inpaint_b200.modules import stylegan2 as SG  # noqa: E402

torch.manual_seed(0)
gen = SG.Generator(1024, 512, 8).cuda().train()
latent = torch.randn(2, gen.n_latent, 512, device="cuda", requires_grad=True)
gout = torch.randn(2, 3, 1024, 1024, device="cuda")
img = gen([latent], input_is_latent=True, randomize_noise=False)[0]
img.backward(gout)
torch.cuda.synchronize()
print(img.shape, float(latent.grad.abs().mean()))
