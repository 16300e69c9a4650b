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


def fill_by_name(module, seed=0):
    """Deterministic, name-keyed values for EVERY parameter (whole-model goldens: the reference model and this package's
    mirror get identical weights without shipping a checkpoint). SpectralNorm u/v are unit vectors, conv / linear weights
    ~ N(0, 1/fan_in), StyleGAN2 modulated-conv weights and the constant input ~ N(0, 1) (they are demodulated / scaled by
    the layer), norm scales ~ 1, PReLU slopes ~ 0.25, biases small, noise strengths 0.1, the zero-initialised attention
    gains (gamma, alpha) non-zero."""
    import zlib
    seen = set()
    with torch.no_grad():
        for mname, mod in module.named_modules():
            kind = type(mod).__name__
            for leaf, p in mod.named_parameters(recurse=False):
                if id(p) in seen:
                    continue
                seen.add(id(p))
                name = f"{mname}.{leaf}" if mname else leaf
                g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
                r = torch.randn(p.shape, generator=g)
                if leaf.endswith('_u') or leaf.endswith('_v'):
                    p.copy_(r / (r.norm() + 1e-12))
                elif leaf == 'gamma':
                    p.fill_(1.0)
                elif leaf == 'alpha':
                    p.fill_(0.5)
                elif kind == 'PReLU':
                    p.copy_(0.25 + 0.05 * r)
                elif kind == 'NoiseInjection':
                    p.fill_(0.1)
                elif kind in ('ModulatedConv2d', 'ConstantInput'):
                    p.copy_(r)
                elif p.dim() >= 2:
                    p.copy_(r / p[0].numel() ** 0.5)
                elif leaf == 'bias':
                    p.copy_((1.0 if mname.endswith('modulation') else 0.0) + 0.05 * r)
                else:  # norm scales
                    p.copy_(1 + 0.1 * r)
        # the StyleGAN2 noise maps are BUFFERS drawn at construction (model.py:441-443): name-keyed too
        for name, b in module.named_buffers():
            if '.noises.noise_' in f'.{name}':
                g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
                b.copy_(torch.randn(b.shape, generator=g))
    return module


def picnet_inputs(n=1, seed=7):
    """Synthetic masked-face / reference / binary-mask inputs of BASELINE config 1: U[0,1) images, lower-face
    rectangle mask."""
    g = torch.Generator().manual_seed(seed)
    src = torch.rand(n, 3, 256, 256, generator=g)
    ref = torch.rand(n, 3, 256, 256, generator=g)
    mask = torch.zeros(n, 256, 256)
    mask[:, 128:230, 50:206] = 1.0
    return src, ref, mask


def mean_z(self, src_distribution, ref_distribution, return_zq=False, mask=None):
    """Replacement for ResGenerator.get_z in the PICNet golden: the distribution means instead of rsample(), so the CPU
    reference and the GPU mirror do not depend on their (different) random generators."""
    q_mu, p_mu = src_distribution[0], ref_distribution[0]
    return q_mu if return_zq else torch.cat([q_mu, p_mu], dim=1)


def refpsp_inputs(n=1, seed=11):
    """Synthetic inputs of BASELINE config 3: U[-1,1) source / reference, binary lower-face mask."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 256, 256, generator=g) * 2 - 1
    ref = torch.rand(n, 3, 256, 256, generator=g) * 2 - 1
    mask = torch.zeros(n, 256, 256)
    mask[:, 128:230, 50:206] = 1.0
    return x, ref, mask
