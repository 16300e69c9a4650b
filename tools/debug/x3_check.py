"""Strict-fp32 contract of the PICNet conv blocks (3xTF32 split operands) against cuDNN strict fp32 and the golden, and what it
costs: ReferenceFill forward, batch 4, three ways. Usage: python tools/debug/x3_check.py"""
import copy
import os
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import rel_err  # noqa: E402
from golden_util import fill_by_name, mean_z, picnet_inputs  # noqa: E402
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref  # noqa: E402

base = fill_by_name(build_picnet_ref()).eval()


def run(batch, force_cudnn, tf32, prec=None, reps=0):
    src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
    m = copy.deepcopy(base).cuda()
    m.decoder.get_z = types.MethodType(mean_z, m.decoder)
    os.environ["FMI_PICNET_CUDNN"] = "1" if force_cudnn else "0"
    if prec:
        os.environ["FMI_PRECISION"] = prec
    else:
        os.environ.pop("FMI_PRECISION", None)
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.no_grad():
        out = m(src, ref, mask, resize=False)
        ms = None
        if reps:
            for _ in range(3):
                m(src, ref, mask)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                m(src, ref, mask)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / reps * 1e3
    return out, ms


g = np.load(ROOT / "tests/golden/picnet_ref.npz")
o1, _ = run(1, False, False)
print(f"batch 1, x3 kernels vs golden (reference on CPU): full-res {rel_err(o1[:, :, ::8, ::8].cpu(), torch.from_numpy(g['full_sub8'])):.3e}")
o1c, _ = run(1, True, False)
print(f"batch 1, cuDNN strict fp32 vs golden:             full-res {rel_err(o1c[:, :, ::8, ::8].cpu(), torch.from_numpy(g['full_sub8'])):.3e}")
truth, t_c = run(4, True, False, reps=10)
truth2, _ = run(4, True, False)
x3, t_x3 = run(4, False, False, reps=10)
tf, t_tf = run(4, False, True, reps=10)
ref_tf, t_ctf = run(4, True, True, reps=10)
print(f"batch 4 eager ms: cuDNN fp32 {t_c:.2f}  cuDNN tf32 {t_ctf:.2f}  kernels x3 {t_x3:.2f}  kernels tf32 {t_tf:.2f}")
print(f"vs cuDNN strict fp32: second cuDNN fp32 run {rel_err(truth2, truth):.3e}  x3 {rel_err(x3, truth):.3e}  "
      f"tf32 kernels {rel_err(tf, truth):.3e}  cuDNN tf32 {rel_err(ref_tf, truth):.3e}")
