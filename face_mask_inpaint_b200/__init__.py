"""face_mask_inpaint_b200 — B200-native (sm_100a) generator hot path of syncdoth/face_mask_inpaint.

Host side mirrors the reference's operator / nn.Module interface for the hot path; all arithmetic runs in
hand-written CUDA kernels behind the C ABI of include/fmi_b200.h (libfmi_b200.so, loaded with ctypes).
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
