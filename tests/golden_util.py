"""Shared by tests/golden/make_golden.py and the golden tests: deterministic parameter construction."""
import torch


def randomize(mod, g):
    """Give the zero/one-initialised parameters non-trivial values (random init hides bugs)."""
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if name.endswith("noise.weight"):
                p.fill_(0.3)
            elif name.endswith("activate.bias") or (name.endswith(".bias") and p.dim() == 4):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("modulation.bias"):
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g))


def build_generator32():
    """The 32x32 StyleGAN2 generator whose output is pinned in tests/golden/generator32.npz."""
    from face_mask_inpaint_b200.modules import stylegan2 as mine
    torch.manual_seed(5)
    gen = mine.Generator(32, 512, 2)
    randomize(gen, torch.Generator().manual_seed(6))
    return gen
