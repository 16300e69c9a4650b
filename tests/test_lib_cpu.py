"""CPU checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports every
symbol include/fmi_b200.h declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes

import pytest

from face_mask_inpaint_b200 import _lib


def test_header_and_prototypes_agree():
    assert _lib.header_symbols() == sorted(_lib.PROTOTYPES)


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    for name in _lib.header_symbols():
        assert hasattr(lib, name), name
    assert lib.fmi_version() >= 100


def test_out_size_formula():
    lib = _lib.load()
    # upfirdn2d.py:100-101  out = (in*up + pad0 + pad1 - k) // down + 1
    for (i, up, down, p0, p1, k) in [(9, 1, 1, 1, 1, 4), (4, 2, 1, 2, 1, 4), (8, 1, 2, 1, 1, 4), (5, 3, 2, 0, 0, 3)]:
        assert lib.fmi_upfirdn2d_out_size(i, up, down, p0, p1, k) == (i * up + p0 + p1 - k) // down + 1
    assert lib.fmi_upfirdn2d_out_size(1, 1, 1, 0, 0, 4) < 0


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    rc = lib.fmi_upfirdn2d(None, None, None, 1, 4, 4, 1, 64, 64, 1, 1, 1, 1, 0, 0, 0, 0, 0, None)
    assert rc == -1 and b"kernel" in lib.fmi_last_error()
    rc = lib.fmi_fused_bias_act(None, None, None, None, 7, 0, 0.2, 1.0, 16, 1, 1, 0, 0, None)
    assert rc == -1


def test_ops_refuse_cpu_tensors():
    import torch
    from face_mask_inpaint_b200 import ops
    with pytest.raises(RuntimeError):
        ops.fused_leaky_relu(torch.zeros(1, 4, 2, 2), torch.zeros(4))
    with pytest.raises(RuntimeError):
        ops.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(4, 4))


def test_reference_checker_libraries_load_when_built():
    """oracle/_ref/*.so (the reference's own CUDA ops compiled by oracle/build_ref.py) import on a CPU-only box and export the
    two pybind ops the GPU parity tests call. Skipped where they were never built (no /root/reference)."""
    import sys
    from pathlib import Path
    import pytest
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import build_ref
    have = build_ref.built()
    if len(have) < 2:
        pytest.skip("oracle/_ref not built")
    assert hasattr(build_ref.load_built("fmi_ref_fused"), "fused_bias_act")
    assert hasattr(build_ref.load_built("fmi_ref_upfirdn2d"), "upfirdn2d")
