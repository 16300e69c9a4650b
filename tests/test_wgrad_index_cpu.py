"""The index arithmetic of the weight-gradient GEMM (csrc/modconv_bwd.cu), restated on the CPU with explicit zero fill and checked
against autograd — no GPU needed:
  * plain 3x3 convolution with the kernel ROWS stacked along M (wgrad_stack): accumulator row block ky of the MMA for column tap kx is
    sum_p g[p + (1 - ky) rows, o] * x[p + (kx - 1) columns, i]  (TMA zero fill outside the image) = dW[o, i, ky, kx];
  * ConvTranspose2d(3, 2, 1, 1) from the gradient's pixel-parity planes (fmi_conv_wgrad_nhwc(transposed = 1)): tap (ky, kx) reads plane
    ((ky + 1) & 1, (kx + 1) & 1) shifted by -[ky == 0] rows, -[kx == 0] columns;
  * the data gradient of that layer as the stride-2 convolution of the same gradient with the weight as it lies (ops._ConvTShared)."""
import torch
import torch.nn.functional as F


def _shift(t, dr, dc):
    """t[..., r + dr, c + dc] with zero fill (what a TMA box at a shifted coordinate delivers); t is [B, H, W, C]."""
    b, h, w, c = t.shape
    out = torch.zeros_like(t)
    r0, r1 = max(0, -dr), min(h, h - dr)
    c0, c1 = max(0, -dc), min(w, w - dc)
    if r0 < r1 and c0 < c1:
        out[:, r0:r1, c0:c1] = t[:, r0 + dr:r1 + dr, c0 + dc:c1 + dc]
    return out


def test_stacked_rows_of_a_plain_convolution():
    g = torch.Generator().manual_seed(1)
    b, i, o, h, w = 2, 5, 3, 6, 7
    x = torch.randn(b, i, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(o, i, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(b, o, h, w, generator=g, dtype=torch.float64)
    F.conv2d(x, wt, None, padding=1).backward(gy)
    xn, gn = x.permute(0, 2, 3, 1), gy.permute(0, 2, 3, 1)               # NHWC, as the kernel sees them
    dw = torch.zeros(3, 3, o, i, dtype=torch.float64)
    for kx in range(3):                                                   # one MMA per column tap ...
        xs = _shift(xn, 0, kx - 1)
        for ky in range(3):                                               # ... whose row block ky holds the gradient tile shifted by 1 - ky rows
            gs = _shift(gn, 1 - ky, 0)
            dw[ky, kx] = torch.einsum("bhwo,bhwi->oi", gs, xs)
    assert torch.allclose(dw.permute(2, 3, 0, 1), wt.grad, atol=1e-12)


def _planes(g):
    """fmi_space_to_planes_nhwc: [B, 2H, 2W, C] -> plane 2 * (row & 1) + (col & 1) -> [4][B, H, W, C]."""
    return [g[:, py::2, px::2] for py in (0, 1) for px in (0, 1)]


def test_transposed_convolution_from_parity_planes():
    g = torch.Generator().manual_seed(2)
    b, i, o, h, w = 2, 4, 3, 5, 6
    x = torch.randn(b, i, h, w, generator=g, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(i, o, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(b, o, 2 * h, 2 * w, generator=g, dtype=torch.float64)
    F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1).backward(gy)
    xn = x.detach().permute(0, 2, 3, 1)
    planes = _planes(gy.permute(0, 2, 3, 1))
    dw = torch.zeros(3, 3, o, i, dtype=torch.float64)
    for ky in range(3):
        for kx in range(3):
            p = planes[(((ky + 1) & 1) << 1) | ((kx + 1) & 1)]
            ps = _shift(p, -1 if ky == 0 else 0, -1 if kx == 0 else 0)    # dy[2m - 1 + ky, 2n - 1 + kx]
            dw[ky, kx] = torch.einsum("bhwo,bhwi->oi", ps, xn)
    assert torch.allclose(dw.permute(3, 2, 0, 1), wt.grad, atol=1e-12)    # dwp[t][o][i] -> weight [I, O, 3, 3]
    # data gradient: Conv2d(O -> I, 3, stride 2, padding 1) of dy with W read as [out = I][in = O]
    dx = F.conv2d(gy, wt.detach(), None, stride=2, padding=1)
    assert torch.allclose(dx, x.grad, atol=1e-12)


def test_attention_column_role_term_is_a_pixel_contraction():
    """attn_bwd.cu step 4: dq[t] = sum_j (dE[t, j] + dE[j, t]) q[j]; the second term contracts over the ROW index of dE — the weight-
    gradient GEMM with 'pixels' = rows of dE, 'output channels' = columns of dE, 'input channels' = the head dimension."""
    g = torch.Generator().manual_seed(3)
    s, d = 12, 4
    de = torch.randn(s, s, generator=g, dtype=torch.float64)
    q = torch.randn(s, d, generator=g, dtype=torch.float64)
    want = (de + de.t()) @ q
    row_role = de @ q                                             # gemm_nt(dE, q^T)
    col_role = torch.einsum("po,pi->oi", de, q)                   # dwp[o][i] = sum_p dy[p, o] x[p, i]
    assert torch.allclose(row_role + col_role, want, atol=1e-12)
