"""Debug: where the issuing threads of attn_fwd3_kernel (CTA 0, the leader of pair 0) wait. Run on the GPU box."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import _lib, ops

lib = _lib.load()
lib.fmi_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"
for name, dtype, mma in [("tf32", torch.float32, _lib.MMA_TF32), ("bf16", torch.bfloat16, _lib.MMA_BF16)]:
    n, c, hw, d = 8, 256, 128, 64
    x = torch.randn(n, c, hw, hw, device=dev).to(dtype)
    ref = torch.randn(n, c, hw, hw, device=dev).to(dtype)
    wq = torch.randn(d, c, 1, 1, device=dev) / c ** 0.5 * 0.6
    mask = torch.rand(n, 1, hw, hw, device=dev)
    fn = lambda: ops.attention_forward(x, wq, None, x, ref, mask=mask, masked1=True, order=(1, 0), mma=mma)
    fn(); torch.cuda.synchronize()
    buf = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
    lib.fmi_debug_set_attn_trace(buf.data_ptr())
    fn(); torch.cuda.synchronize()
    lib.fmi_debug_set_attn_trace(None)
    t = buf.cpu()[:16].tolist()
    steps = hw * hw // 64
    print(f"== {name}: {steps} steps; cycles per step (CTA 0)")
    print(f"   PV issuer loop total      {t[6] / steps:8.0f}   wait P ready {t[4] / steps:7.0f}   wait V tile {t[5] / steps:7.0f}")
    print(f"   QK issuer: wait Q staged {t[0]:8.0f} (once)   wait PV done (buffer free) {t[1] / steps:7.0f}   wait K tile {t[2] / steps:7.0f}")
    print(f"   softmax WG0: wait S {t[8] / (steps / 2):7.0f}  ld+exp+st {t[9] / (steps / 2):7.0f} per own step;  "
          f"WG1: wait S {t[10] / (steps / 2):7.0f}  ld+exp+st {t[11] / (steps / 2):7.0f}")
    print(f"   softmax WG0 split: tcgen05.ld+wait {t[12] / (steps / 2):7.0f}  exp loop {t[13] / (steps / 2):7.0f};  "
          f"WG1: ld {t[14] / (steps / 2):7.0f}  exp {t[15] / (steps / 2):7.0f}")
