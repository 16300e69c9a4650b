"""PICNet reference-fill generator TRAIN step (BASELINE config 4, generator side): ReferenceFill forward (z ~ N(mu, sigma) as in
train_reference_fill.py:342), L1 reconstruction loss, backward, Adam(1e-5) step — NOT a pytest file:

    python tools/perf/perf_picnet_train.py > gpurun_out/perf_picnet_train.txt

`ours`   : both attention modules forward AND backward on the sm_100a kernels (fmi_attn_fwd / fmi_attn_bwd, no S x S map);
           the conv blocks under autograd are cuDNN (their backward is not a kernel of this package yet).
`ref-GPU`: the same network with the attention computed the reference's way (bmm + softmax + bmm under ATen autograd, the
           S x S map of 1 GiB per image materialised and saved for backward), conv blocks on cuDNN.
The discriminator, VGG perceptual loss and GANOptimizer of the script are the reference's own PyTorch (out of scope) and are
not part of either arm.
"""
import copy
import os
import sys
from pathlib import Path

import torch
from torch import nn

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref  # noqa: E402
from golden_util import fill_by_name, picnet_inputs  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


class TorchEGA(nn.Module):
    def __init__(self, mod):
        super().__init__()
        self.conv = mod.conv

    def forward(self, mask, src, ref):
        return O.example_guided_attention(mask, src, ref, self.conv.weight)


class TorchAutoAttn(nn.Module):
    def __init__(self, mod):
        super().__init__()
        self.query_conv, self.gamma = mod.query_conv, mod.gamma

    def forward(self, x, pre=None, mask=None):
        return O.auto_attn(x, self.query_conv.weight, self.query_conv.bias, self.gamma)[0], None


def step_time(model, batch, iters=5):
    src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
    gt = torch.rand(batch, 3, 256, 256, device="cuda")
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-5)

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(src, ref, mask)
        loss = (out - gt).abs().mean()
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, torch.cuda.max_memory_allocated() / 2 ** 30, float(loss)


def main():
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = False   # the reference's bmm default
    base = fill_by_name(build_picnet_ref()).train()
    with torch.no_grad():
        base.decoder.attn1.gamma.fill_(1.0)
    print(f"{'case':60s} {'ms/step':>9s} {'img/s':>8s} {'peak GiB':>9s} {'loss':>8s}")
    for batch in (1, 2, 4, 8):
        ours = copy.deepcopy(base).cuda()
        t, mem, loss = step_time(ours, batch)
        print(f"{f'ours (attention fwd+bwd kernels) B={batch}':60s} {t:9.2f} {batch / t * 1e3:8.1f} {mem:9.2f} {loss:8.4f}", flush=True)
        del ours
        torch.cuda.empty_cache()
        if batch * 16384 * 16384 * 4 * 6 > 150e9:
            print(f"ref-GPU B={batch}: skipped (S x S maps and their autograd copies would not fit)")
            continue
        refm = copy.deepcopy(base)
        refm.attention = TorchEGA(refm.attention)
        refm.decoder.attn1 = TorchAutoAttn(refm.decoder.attn1)
        refm = refm.cuda()
        try:
            t, mem, loss = step_time(refm, batch, iters=3)
            print(f"{f'ref-GPU (bmm + softmax autograd) B={batch}':60s} {t:9.2f} {batch / t * 1e3:8.1f} {mem:9.2f} {loss:8.4f}", flush=True)
        except torch.cuda.OutOfMemoryError:
            print(f"ref-GPU B={batch}: out of memory")
        del refm
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
