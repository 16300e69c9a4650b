"""Host-side checks of the strict-fp32 contract of the conv blocks: the operand split that fmi_tf32_split3
implements, emulated with numpy bit operations (kind::tf32 reads the upper 19 bits of an fp32 operand), and the switch that
selects it (ops.tf32_split)."""
import numpy as np
import pytest
import torch


def _trunc_tf32(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _split(x):
    hi = _trunc_tf32(x)
    lo = (x.astype(np.float32) - hi).astype(np.float32)
    return hi, lo


def test_split_is_exact_and_three_terms_reach_fp32_class():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 576)).astype(np.float32)
    w = (rng.standard_normal((576, 32)) / 24).astype(np.float32)
    xh, xl = _split(x)
    wh, wl = _split(w)
    assert np.array_equal(xh + xl, x) and np.array_equal(wh + wl, w)          # lo = x - hi is exact in fp32
    assert np.all(np.abs(xl) <= np.abs(x) * 2.0 ** -10 + 1e-45)
    # what the tensor core sees: every operand truncated to tf32 again (hi unchanged, lo loses its last bits), exact products
    t = lambda a: _trunc_tf32(a).astype(np.float64)
    want = x.astype(np.float64) @ w.astype(np.float64)
    single = t(x) @ t(w)
    # channel layout of the one GEMM: activations [hi | hi | lo], weights [hi | lo | hi] along K
    a3 = np.concatenate([xh, xh, xl], axis=1)
    w3 = np.concatenate([wh, wl, wh], axis=0)
    three = t(a3) @ t(w3)
    scale = np.abs(want).max()
    e1, e3 = np.abs(single - want).max() / scale, np.abs(three - want).max() / scale
    assert e1 > 1e-4                      # single-pass TF32: ~3e-4
    assert e3 < 2e-6, e3                  # split operands: 2^-20-class
    assert e3 < e1 / 100


@pytest.mark.parametrize("env,allow,want", [("", True, False), ("", False, True), ("tf32x3", True, True), ("fp32", False, False),
                                            ("bf16", False, False), ("tf32", False, False)])
def test_switch(monkeypatch, env, allow, want):
    from face_mask_inpaint_b200 import _lib, ops
    if env:
        monkeypatch.setenv("FMI_PRECISION", env)
    else:
        monkeypatch.delenv("FMI_PRECISION", raising=False)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = allow
    try:
        assert ops.tf32_split() is want
        if env == "tf32x3":
            assert ops.mma_mode(torch.float32) == _lib.MMA_TF32
    finally:
        torch.backends.cudnn.allow_tf32 = old
