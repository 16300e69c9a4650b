"""Diagnostic (not a pytest file): run fmi_styled_conv_bwd_nhwc with act = 0 and read the per-sample weight gradient
G[b][t][o][i] straight out of the workspace; compare with a torch reference. Usage: python tools/debug/diag_sg2_bwd.py"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import _lib, ops  # noqa: E402

dev = "cuda"
lib = _lib.load()


def align(v, a=1024):
    return (v + a - 1) // a * a


def run(b, i, o, h, w, up, mma, pattern="rand"):
    esz = 4 if mma == _lib.MMA_TF32 else 2
    dt = torch.float32 if mma == _lib.MMA_TF32 else torch.bfloat16
    oh, ow = (2 * h, 2 * w) if up else (h, w)
    g = torch.Generator().manual_seed(0)
    if pattern == "ones":
        x = torch.ones(b, h, w, i)
        dy = torch.ones(b, oh, ow, o)
    else:
        x = torch.randn(b, h, w, i, generator=g)
        dy = torch.randn(b, oh, ow, o, generator=g)
    x = x.to(dt).to(dev)
    dy = dy.to(dt).to(dev)
    weight = torch.randn(1, o, i, 3, 3, generator=g).to(dev)
    s = (1 + 0.1 * torch.randn(b, i, generator=g)).to(dev)
    k = torch.tensor([1., 3., 3., 1.])
    k = (k[None] * k[:, None])
    k = (k / k.sum() * 4).to(dev)
    nbytes = lib.fmi_styled_conv_bwd_workspace_bytes(b, i, o, h, w, int(up), 0, mma)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dx = torch.empty_like(x)
    dw = torch.empty(o, i, 3, 3, device=dev)
    ds = torch.empty(b, i, device=dev)
    rc = lib.fmi_styled_conv_bwd_nhwc(x.data_ptr(), None, dy.data_ptr(), weight.data_ptr(), s.data_ptr(), None, 0,
                                      k.data_ptr() if up else None, dx.data_ptr(), dw.data_ptr(), ds.data_ptr(), None, None,
                                      b, i, o, h, w, int(up), 0, 1, mma, ws.data_ptr(), ws.numel(), ops._stream())
    torch.cuda.synchronize()
    if rc:
        print("  rc", rc, _lib.last_error())
        return
    off = 0
    if up:
        off += align(4 * b * (h + 1) * (w + 1) * o * esz)
    off += 2 * align(b * 9 * o * i * esz)
    G = ws[off:off + b * 9 * o * i * 4].view(torch.float32).view(b, 9, o, i).cpu()
    # reference: G[b][t][o][i]
    xf = x.float().cpu().permute(0, 3, 1, 2)    # [b,i,h,w]
    gf = dy.float().cpu().permute(0, 3, 1, 2)   # [b,o,oh,ow]
    ref = torch.zeros(b, 9, o, i)
    if up:
        kk = k.cpu()
        # gmid = blur^T g : conv_transpose of the forward blur (forward: upfirdn2d(mid, k, pad=(1,1)))
        mid = torch.zeros(b, o, 2 * h + 1, 2 * w + 1, requires_grad=True)
        from oracle import ref_ops as O
        out = O.upfirdn2d(mid, kk, pad=(1, 1))
        out.backward(gf)
        gmid = mid.grad
        for t in range(9):
            ky, kx = t // 3, t % 3
            sub = gmid[:, :, ky:ky + 2 * h:2, kx:kx + 2 * w:2]       # [b,o,h,w]
            ref[:, t] = torch.einsum('bohw,bihw->boi', sub, xf)
    else:
        xp = F.pad(xf, (1, 1, 1, 1))
        for t in range(9):
            ky, kx = t // 3, t % 3
            ref[:, t] = torch.einsum('bohw,bihw->boi', gf, xp[:, :, ky:ky + h, kx:kx + w])
    err = ((G - ref).abs().max() / ref.abs().max()).item()
    print(f"  B={b} I={i} O={o} {h}x{w} up={up} mma={'tf32' if mma == 0 else 'bf16'} {pattern}: |G|max {G.abs().max().item():.4g} "
          f"|ref|max {ref.abs().max().item():.4g} rel err {err:.3e}")
    if err > 2e-2:
        for t in range(9):
            e = ((G[:, t] - ref[:, t]).abs().max() / ref.abs().max()).item()
            print(f"     tap {t}: err {e:.3e}  G[0,t,0,:4] {G[0, t, 0, :4].tolist()} ref {ref[0, t, 0, :4].tolist()}")
        # which rows / cols are right?
        e_rows = (G - ref).abs().amax(dim=(0, 1, 3)) / ref.abs().max()
        e_cols = (G - ref).abs().amax(dim=(0, 1, 2)) / ref.abs().max()
        print("     bad rows(o):", (e_rows > 2e-2).nonzero().flatten().tolist()[:40])
        print("     bad cols(i):", (e_cols > 2e-2).nonzero().flatten().tolist()[:40])


for mma in (_lib.MMA_BF16, _lib.MMA_TF32):
    for pattern in ("ones", "rand"):
        run(1, 64, 64, 16, 16, False, mma, pattern)
    run(2, 64, 64, 16, 16, False, mma)
    run(1, 128, 128, 16, 16, False, mma)
    run(1, 32, 32, 16, 16, False, mma)
    run(1, 64, 64, 16, 16, True, mma)
    run(1, 256, 256, 8, 8, False, mma)
    run(1, 512, 512, 4, 4, False, mma)
    run(1, 64, 64, 64, 64, False, mma)
