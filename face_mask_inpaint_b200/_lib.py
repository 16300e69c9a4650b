"""ctypes binding of libfmi_b200.so — the C ABI declared in include/fmi_b200.h.

There is no CPU fallback: if the shared library is missing it is built with nvcc, and if that is
impossible or a kernel fails, the caller gets a RuntimeError carrying fmi_last_error().
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libfmi_b200.so"

F32, BF16, F16 = 0, 1, 2
MMA_TF32, MMA_BF16 = 0, 1

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol of include/fmi_b200.h
PROTOTYPES = {
    "fmi_version": (_i, []),
    "fmi_last_error": (C.c_char_p, []),
    "fmi_device_check": (_i, []),
    "fmi_kernel_launch_count": (C.c_longlong, []),
    "fmi_profile_enable": (_i, [_i]),
    "fmi_profile_collect": (_i, [_i, C.POINTER(C.c_double), C.POINTER(_i)]),
    "fmi_profile_kinds": (_i, []),
    "fmi_profile_kind_name": (C.c_char_p, [_i]),
    "fmi_profile_dump": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), _i, C.POINTER(_i)]),
    "fmi_fused_bias_act": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _f, _i64, _i64, _i64, _i, _i, _vp]),
    "fmi_bias_act_bwd": (_i, [_vp, _vp, _vp, _vp, _f, _f, _i64, _i64, _i64, _i, _vp]),
    "fmi_upfirdn2d": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_upfirdn2d_out_size": (_i, [_i, _i, _i, _i, _i, _i]),
    "fmi_scale_mask": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_composite": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_composite_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_attn_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i]),
    "fmi_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _f, _i, _vp, _i64, _vp, _i64, _vp, _vp,
                          _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "fmi_attn_materialize": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fmi_attn_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i]),
    "fmi_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _f, _i, _vp, _vp, _vp, _i64, _vp, _i64,
                          _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    "fmi_conv1x1": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_nchw_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_nhwc_to_nchw": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_style_modulation": (_i, [_vp, _i64, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "fmi_modconv_weight_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "fmi_modconv_weight_prep": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_styled_conv_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i]),
    "fmi_styled_conv_nhwc": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i64,
                                  _vp]),
    "fmi_torgb_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_torgb_weights": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "fmi_styled_conv_bwd_workspace_bytes": (_i64, [_i, _i, _i, _i, _i, _i, _i, _i]),
    "fmi_styled_conv_bwd_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i,
                                      _i, _i, _i, _i, _vp, _i64, _vp]),
    "fmi_torgb_bwd_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_styled_conv_torgb_nhwc": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                                        _vp]),
    "fmi_conv_weight_prep": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_conv_weight_prep_sn": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_conv_weight_prep_sn_batch": (_i, [_vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_nchw_to_nhwc_slice": (_i, [_vp, _vp, _i, _i, _i, _i, _i64, _i, _i, _i, _vp]),
    "fmi_instnorm_stats_nhwc": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp]),
    "fmi_norm_act_nhwc": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _f, _i, _vp]),
    "fmi_output_conv_tanh": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_avgpool2_nhwc": (_i, [_vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_tf32_split3": (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp]),
    "fmi_set_tf32_exact": (_i, [_i]),
    "fmi_get_tf32_exact": (_i, []),
    "fmi_reflect_border_nhwc": (_i, [_vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_conv3x3_nhwc": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "fmi_conv_nhwc": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _i, _vp, _f, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                           _i, _i, _vp]),
    "fmi_space_to_planes_nhwc": (_i, [_vp, _i64, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_se_gate_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_se_scale_add_nhwc": (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _i, _i, _i, _i, _i, _vp]),
    "fmi_upsample_add_nhwc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_avgpool_planes": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp]),
    "fmi_instnorm_act_bwd_nhwc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "fmi_gemm_nt": (_i, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_conv_wgrad_nhwc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fmi_spectral_norm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "fmi_spectral_norm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
}

_lib = None


def header_symbols() -> list[str]:
    """Function names declared in include/fmi_b200.h (parsed, so tests can diff against PROTOTYPES)."""
    import re
    text = (_PKG.parent / "include" / "fmi_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmi_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libfmi_b200.so. A missing library is built with nvcc under an inter-process file lock (torchrun starts N ranks on a
    fresh checkout at once) and renamed into place atomically (build.py). A library older than its sources is refused loudly
    when nvcc is here to rebuild it (FMI_ALLOW_STALE=1 loads it anyway); on a box without the sources' toolchain the shipped
    library is used as it is."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing; run `python -m face_mask_inpaint_b200.build`")
        _build.build(verbose=bool(os.environ.get("FMI_VERBOSE")))
    elif _build.stale() and os.environ.get("FMI_ALLOW_STALE") != "1":
        if build_if_missing and _build.have_nvcc():
            _build.build(verbose=bool(os.environ.get("FMI_VERBOSE")))
        else:
            raise RuntimeError(f"{LIB_PATH} is older than csrc/ (run `python -m face_mask_inpaint_b200.build`, or FMI_ALLOW_STALE=1)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().fmi_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")
