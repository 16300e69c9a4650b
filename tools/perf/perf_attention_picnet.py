"""Auto_Attn of the PICNet decoder (C = 64, d = 16, 128^2, batch 4: the largest single launch of the PICNet-ref forward) with the
fast kernel's debug knobs, to attribute its time:  python tools/perf/perf_attention_picnet.py
FMI_ATTN_DBG bits: 1 skip the TMEM load of S, 2 skip the exponentials, 4 skip the P store, 8 skip the PV MMAs."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import _lib, ops  # noqa: E402


def time_cuda(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
for (n, c, d, hw) in [(4, 64, 16, 128), (4, 64, 64, 128), (4, 256, 16, 128), (8, 256, 64, 32)]:
    x = torch.randn(n, c, hw, hw, device="cuda")
    wq = torch.randn(d, c, 1, 1, device="cuda") / c ** 0.5 * 0.5
    gamma = torch.tensor([0.7], device="cuda")
    for mma, name in ((_lib.MMA_TF32, "tf32"), (_lib.MMA_BF16, "bf16")):
        row = []
        for env in ({}, {"FMI_ATTN_KTRIM": "0"}, {"FMI_ATTN_DBG": "2"}, {"FMI_ATTN_DBG": "8"}, {"FMI_ATTN_DBG": "10"}, {"FMI_ATTN_DBG": "7"},
                    {"FMI_ATTN_CLUSTER": "0"}):
            for k in ("FMI_ATTN_DBG", "FMI_ATTN_KTRIM", "FMI_ATTN_CLUSTER"):
                os.environ.pop(k, None)
            os.environ.update(env)
            if "FMI_ATTN_KTRIM" in env or "FMI_ATTN_CLUSTER" in env:
                row.append(float("nan"))      # read once per process: see the separate runs
                continue
            t = time_cuda(lambda: ops.attention_forward(x, wq, None, x, None, a0=gamma, b0=1.0, mma=mma)[0])
            row.append(t)
        print(f"N={n} C={c} d={d} {hw}^2 {name}: full {row[0]:.3f}  no-exp {row[2]:.3f}  no-PV {row[3]:.3f}  no-exp-no-PV {row[4]:.3f}  "
              f"no-ld/exp/st {row[5]:.3f} ms", flush=True)
