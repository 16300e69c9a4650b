"""Debug: time the fast attention kernel with parts switched off (FMI_ATTN_DBG bit flags). GPU box only."""
import os, sys, subprocess
from pathlib import Path
if len(sys.argv) == 1:
    for dbg in [0, 1, 2, 4, 8, 16, 3, 7, 24, 5, 13]:
        env = dict(os.environ, FMI_ATTN_DBG=str(dbg))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(f"dbg={dbg:2d} (1=no ld, 2=no exp, 4=no st, 8=no PV mma, 16=no QK mma): {out.stdout.strip()} {out.stderr.strip()[-200:]}")
else:
    import torch
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
    from face_mask_inpaint_b200 import _lib, ops
    dev = "cuda"
    res = []
    for name, dtype, mma in [("tf32", torch.float32, _lib.MMA_TF32), ("bf16", torch.bfloat16, _lib.MMA_BF16)]:
        n, c, hw, d = 8, 256, 128, 64
        x = torch.randn(n, c, hw, hw, device=dev).to(dtype)
        ref = torch.randn(n, c, hw, hw, device=dev).to(dtype)
        wq = torch.randn(d, c, 1, 1, device=dev) / c ** 0.5 * 0.6
        mask = torch.rand(n, 1, hw, hw, device=dev)
        fn = lambda: ops.attention_forward(x, wq, None, x, ref, mask=mask, masked1=True, order=(1, 0), mma=mma)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        res.append(f"{name} {e0.elapsed_time(e1)/5:6.3f} ms")
    print("  ".join(res))
