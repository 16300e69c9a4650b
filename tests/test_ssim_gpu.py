"""north_star: "plus SSIM parity on generated images". Images generated on the GPU by this package's kernels against the images
the reference's own models produced (CPU fp32 goldens recorded from /root/reference), compared with the reference's own SSIM
(modules/evaluations/ssim.py, restated as oracle.ssim and pinned by tests/golden/ssim.npz). Images are mapped to [0, 1] first
(the scripts' convention, PICNet_inference.py:124-131)."""
import os
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from golden_util import build_generator32, fill_by_name, mean_z, picnet_inputs, refpsp_inputs
from oracle import ref_ops as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _unit(img):
    return (img.float().cpu().clamp(-1, 1) + 1) / 2


@pytest.mark.parametrize("prec,floor", [("fp32", 0.9995), ("bf16", 0.995)])
def test_picnet_generated_image_ssim(prec, floor, monkeypatch):
    """ReferenceFill 256^2 (every conv block and both attentions on the kernels, TF32 / bf16 operands) vs the reference's image."""
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    monkeypatch.setenv("FMI_PRECISION", prec)
    g = np.load(GOLD / "picnet_ref.npz")
    m = fill_by_name(build_picnet_ref()).eval().cuda()
    m.decoder.get_z = types.MethodType(mean_z, m.decoder)
    src, ref, mask = (t.cuda() for t in picnet_inputs(1))
    assert torch.backends.cudnn.allow_tf32          # PyTorch's default: the kernel path is the one under test
    with torch.no_grad():
        out = m(src, ref, mask)
    s = float(O.ssim(_unit(out), _unit(torch.from_numpy(g["image"]))))
    print(f"SSIM {prec}: {s:.5f}")
    assert s >= floor, s


@pytest.mark.parametrize("prec,floor", [("fp32", 0.9995), ("bf16", 0.995)])
def test_refpsp_generated_image_ssim(prec, floor, monkeypatch):
    from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts
    monkeypatch.setenv("FMI_PRECISION", prec)
    g = np.load(GOLD / "refpsp256.npz")
    net = fill_by_name(pSp(refpsp_opts(output_size=256))).eval().cuda()
    x, ref, mask = (t.cuda() for t in refpsp_inputs(1))
    with torch.no_grad():
        img = net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
    s = float(O.ssim(_unit(img), _unit(torch.from_numpy(g["image"]))))
    print(f"SSIM {prec}: {s:.5f}")
    assert s >= floor, s


@pytest.mark.parametrize("prec,floor", [("fp32", 0.9995), ("bf16", 0.995)])
def test_stylegan2_generated_image_ssim(prec, floor, monkeypatch):
    monkeypatch.setenv("FMI_PRECISION", prec)
    g = np.load(GOLD / "generator32.npz")
    gen = build_generator32().eval().cuda()
    with torch.no_grad():
        img, _ = gen([torch.from_numpy(g["latent"]).cuda()], input_is_latent=True, randomize_noise=False)
    want = torch.from_numpy(g["image"])
    scale = float(want.abs().max())      # random-init synthesis output is not confined to [-1, 1]: normalise both the same way
    s = float(O.ssim(_unit(img / scale), _unit(want / scale)))
    print(f"SSIM {prec}: {s:.5f}")
    assert s >= floor, s
