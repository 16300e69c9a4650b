"""RefpSp forward (BASELINE config 3 at a reduced output size) on the GPU against the reference's own pSp run on CPU
(tests/golden/refpsp256.npz from tests/golden/make_golden.py refpsp): IR-SE50 trunk + map2style heads on cuDNN (TF32 off for
the comparison), attention1/attention2, the masked blends and the whole StyleGAN2 decoder on this package's kernels."""
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import rel_err
from golden_util import fill_by_name, refpsp_inputs

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "refpsp256.npz"


def _net():
    from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts
    return fill_by_name(pSp(refpsp_opts(output_size=256))).eval().cuda()


def test_refpsp_forward_matches_reference():
    g = np.load(GOLD)
    x, ref, mask = refpsp_inputs(1)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            img, codes = _net()(x.cuda(), ref=ref.cuda(), src_mask=mask.cuda(), resize=True, randomize_noise=False,
                                return_latents=True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    e_codes = rel_err(codes, torch.from_numpy(g["codes"]))
    e_img = rel_err(img, torch.from_numpy(g["image"]))
    # codes: encoder (cuDNN fp32) + two attention kernels + compositing; image: + 13 StyledConv / 7 ToRGB with TF32 operands
    assert e_codes <= 1e-3 and e_img <= 3e-3, f"codes {e_codes:.3e} image {e_img:.3e}"


def test_refpsp_eval_batches_src_and_ref_together():
    """In eval mode source and reference share one trunk pass; the result must equal two separate passes."""
    net = _net()
    x, ref, mask = (t.cuda() for t in refpsp_inputs(2))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False   # cuDNN picks per-batch-size algorithms; keep them all fp32
    try:
        with torch.no_grad():
            a = net.encoder(x, ref=ref, mask=mask)
            c = [net.encoder._trunk(x), net.encoder._trunk(ref)]
            b1 = net.encoder._trunk(torch.cat([x, ref]))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for k in range(3):
        assert rel_err(b1[k][:2], c[0][k]) <= 1e-4 and rel_err(b1[k][2:], c[1][k]) <= 1e-4
    assert a.shape == (2, 14, 512)
