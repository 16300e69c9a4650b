"""One StyleGAN2-1024 synthesis forward, batch 8 (for ncu launch lists). GPU box only. FMI_PRECISION selects the mode."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from face_mask_inpaint_b200.modules import stylegan2 as SG
dev = "cuda"
torch.manual_seed(0)
gen = SG.Generator(1024, 512, 8).to(dev).eval()
latent = torch.randn(8, gen.n_latent, 512, device=dev)
with torch.no_grad():
    for _ in range(2):
        img, _ = gen([latent], input_is_latent=True, randomize_noise=False)
torch.cuda.synchronize()
print("ok", tuple(img.shape))
