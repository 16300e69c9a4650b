"""Debug: per-tile timeline of the attention kernel (CTA 0) from clock64 stamps. Run on the GPU box."""
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
    t = buf.cpu().view(256, 16)[:128]
    t0 = t[0, 0].item()
    names = {0: "sm_ready", 1: "S_avail", 2: "S_in_regs", 3: "max_xchg", 4: "exp_done", 5: "P_pub", 7: "PVprev_done", 8: "QKnext_issued", 9: "P_avail(mma)", 10: "PV_issued"}
    print(f"== {name}: stamps relative to tile start (softmax ready), averaged over tiles 8..120")
    sel = t[8:120]
    per = (sel[1:, 0] - sel[:-1, 0]).float().mean().item()
    print(f"   period per tile: {per:.0f} clks")
    for k, nm in names.items():
        print(f"   {nm:16s} {(sel[:, k] - sel[:, 0]).float().mean().item():9.0f}")
    for j in (10, 11, 12):
        print("   tile", j, [(t[j, k].item() - t0) for k in (0, 1, 2, 3, 4, 5, 7, 8, 9, 10)])
