"""Minimal PyTorch mirror of the PICNet ResBlock + SpectralNorm that Auto_Attn owns as `model`
(modules/pluralistic_model/base_function.py:207-268, external_function.py:16-72).

OUT OF SCOPE of the CUDA hot path (plain cuDNN conv block): it exists only so that a
standalone `Auto_Attn` keeps the reference's parameter names/shapes (`model.conv1.module.weight_bar`, ...,
strict state_dict compatibility) and so the dead-in-this-repo `pre` branch stays callable. When the package
is installed over the reference (patch.py) the reference's own ResBlock is used instead.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


def _l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


def fused_spectral_norm_ok(sn, w, u, v) -> bool:
    """The fused SpectralNorm kernels (fmi_spectral_norm_fwd / _bwd) apply: one power iteration, CUDA fp32 contiguous
    parameters. FMI_SN_TORCH=1 keeps the reference's ATen formulation."""
    import os
    return (sn.power_iterations == 1 and w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and u.is_contiguous()
            and v.is_contiguous() and u.dtype == torch.float32 and os.environ.get("FMI_SN_TORCH") != "1")


class SpectralNorm(nn.Module):
    """external_function.py:16-72 — one power iteration per forward (u, v mutate even in eval, as upstream).
    u and v are updated IN PLACE (the reference re-binds `.data`; same values) so that a CUDA-graph replay of the forward
    keeps advancing them (graphs.py)."""

    def __init__(self, module, name='weight', power_iterations=1):
        super().__init__()
        self.module = module
        self.name = name
        self.power_iterations = power_iterations
        w = getattr(module, name)
        height = w.data.shape[0]
        width = w.view(height, -1).data.shape[1]
        u = nn.Parameter(_l2normalize(w.data.new(height).normal_(0, 1)), requires_grad=False)
        v = nn.Parameter(_l2normalize(w.data.new(width).normal_(0, 1)), requires_grad=False)
        w_bar = nn.Parameter(w.data)
        del module._parameters[name]
        module.register_parameter(name + "_u", u)
        module.register_parameter(name + "_v", v)
        module.register_parameter(name + "_bar", w_bar)

    def _update_u_v(self):
        u = getattr(self.module, self.name + "_u")
        v = getattr(self.module, self.name + "_v")
        w = getattr(self.module, self.name + "_bar")
        if fused_spectral_norm_ok(self, w, u, v):      # CUDA: power iteration + division as 3 kernels, backward as 2 (ops.py)
            setattr(self.module, self.name, ops.spectral_norm_weight(w, u, v))
            return
        height = w.data.shape[0]
        for _ in range(self.power_iterations):
            v.data.copy_(_l2normalize(torch.mv(torch.t(w.view(height, -1).data), u.data)))
            u.data.copy_(_l2normalize(torch.mv(w.view(height, -1).data, v.data)))
        sigma = u.dot(w.view(height, -1).mv(v))
        setattr(self.module, self.name, w / sigma.expand_as(w))

    def forward(self, *args):
        self._update_u_v()
        if len(args) == 1 and ops.conv_train_supported(self.module, args[0]):   # training: forward + backward on the kernels
            return ops.conv_train(self.module, args[0])
        return self.module.forward(*args)


def _conv(i, o, use_spect, **kw):
    c = nn.Conv2d(i, o, **kw)
    return SpectralNorm(c) if use_spect else c


class ResBlock(nn.Module):
    """base_function.py:207-268 with sample_type='none', use_coord=False (the only form Auto_Attn builds)."""

    def __init__(self, input_nc, output_nc, hidden_nc=None, norm_layer=nn.BatchNorm2d, nonlinearity=None,
                 use_spect=False):
        super().__init__()
        nonlinearity = nonlinearity if nonlinearity is not None else nn.LeakyReLU()
        hidden_nc = output_nc if hidden_nc is None else hidden_nc
        self.conv1 = _conv(input_nc, hidden_nc, use_spect, kernel_size=3, stride=1, padding=1)
        self.conv2 = _conv(hidden_nc, output_nc, use_spect, kernel_size=3, stride=1, padding=1)
        self.bypass = _conv(input_nc, output_nc, use_spect, kernel_size=1, stride=1, padding=0)
        if norm_layer is None:
            self.model = nn.Sequential(nonlinearity, self.conv1, nonlinearity, self.conv2)
        else:
            self.model = nn.Sequential(norm_layer(input_nc), nonlinearity, self.conv1, norm_layer(hidden_nc),
                                       nonlinearity, self.conv2)
        self.shortcut = nn.Sequential(self.bypass)

    def forward(self, x):
        return self.model(x) + self.shortcut(x)
