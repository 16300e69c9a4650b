"""Drop-in ExampleGuidedAttention (modules/example_guided_att.py:5-41) and Auto_Attn
(modules/pluralistic_model/base_function.py:401-448) on the fused sm_100a attention kernels (forward: fmi_attn_fwd,
backward: fmi_attn_bwd).

Parameter names/shapes are the reference's: `conv.weight [C/4,C,1,1]`, `out_conv.{weight,bias}`;
`query_conv.{weight,bias}`, `gamma`, `alpha`, `model.*`.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from .. import ops


class ExampleGuidedAttention(nn.Module):
    """modules/example_guided_att.py:5-41."""

    def __init__(self, in_channels, out_channels=None):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels // 4, 1, bias=False)  # parameter holder (:9)
        self.out_channels = out_channels
        if out_channels is not None:
            self.out_conv = nn.Conv2d(in_channels * 2, out_channels, 1)      # parameter holder (:13)

    def forward(self, src_mask, src_feature, ref_feature):
        """src_mask [N,1,H,W] (already at feature resolution), src/ref features [N,C,H,W] ->
        [N, 2C or out_channels, H, W]: cat[(1-m)*ref_att + m*ref, src_att] (:34-36) (+ out_conv :38-39)."""
        out = _EGAFunction.apply(src_mask, src_feature, ref_feature, self.conv.weight)
        if self.out_channels is not None:
            out = _Conv1x1Function.apply(out, self.out_conv.weight, self.out_conv.bias)
        return out


class _EGAFunction(torch.autograd.Function):
    """value group 0 = src (out = O_src -> channels [C,2C)), group 1 = ref (masked blend -> channels [0,C))."""

    @staticmethod
    def forward(ctx, src_mask, src_feature, ref_feature, wq):
        need = any(ctx.needs_input_grad[1:])
        if need:
            out, lse, _, o_saved = ops.attention_forward(src_feature, wq, None, src_feature, ref_feature, mask=src_mask,
                                                         b0=0.0, masked0=False, masked1=True, order=(1, 0),
                                                         need_lse=True, save_o=True)
            ctx.save_for_backward(src_mask, src_feature, ref_feature, wq, o_saved, lse)
        else:
            out, _, _ = ops.attention_forward(src_feature, wq, None, src_feature, ref_feature, mask=src_mask, b0=0.0,
                                              masked0=False, masked1=True, order=(1, 0))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        src_mask, src, ref, wq, o_saved, lse = ctx.saved_tensors
        dq, dv0, dv1, _, _ = ops.attention_backward(src, wq, None, src, ref, src_mask, None, 0.0, False, None, 0.0, True,
                                                    o_saved, lse, grad_out, order=(1, 0),
                                                    need_dv0=ctx.needs_input_grad[1], need_dv1=ctx.needs_input_grad[2])
        dw, dx, _ = ops.qconv_backward(dq, src, wq, False)
        g_src = None
        if ctx.needs_input_grad[1]:
            g_src = (dv0 + dx).to(src.dtype)
        g_ref = dv1.to(ref.dtype) if ctx.needs_input_grad[2] else None
        return None, g_src, g_ref, (dw.to(wq.dtype) if ctx.needs_input_grad[3] else None)


class _Conv1x1Function(torch.autograd.Function):
    """out_conv (example_guided_att.py:38-39): forward on fmi_conv1x1; its gradients are plain dense products."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return ops.conv1x1(x, weight, bias)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        n, co = grad_out.shape[0], grad_out.shape[1]
        ci = x.shape[1]
        g = grad_out.reshape(n, co, -1).float()
        xf = x.reshape(n, ci, -1).float()
        dw = torch.einsum('nos,ncs->oc', g, xf).reshape(weight.shape).to(weight.dtype)
        dx = torch.einsum('oc,nos->ncs', weight.reshape(co, ci).float(), g).reshape(x.shape).to(x.dtype)
        db = g.sum((0, 2)).to(weight.dtype) if ctx.has_bias else None
        return dx, dw, db


class _AutoAttnFunction(torch.autograd.Function):
    """group 0 = x (out = gamma*O + x); optional group 1 = pre (alpha*(1-m)*O_pre + m*pre)."""

    @staticmethod
    def forward(ctx, x, wq, bq, gamma, pre, mask, alpha):
        need = any(ctx.needs_input_grad)
        kw = dict(a0=gamma, b0=1.0)
        if pre is not None:
            kw.update(mask=mask, a1=alpha, masked1=True, order=(0, 1))
        res = ops.attention_forward(x, wq, bq, x, pre, need_lse=need or _materialize(), save_o=need, **kw)
        out, lse, ws = res[0], res[1], res[2]
        attn = None
        if _materialize():
            n, s = x.shape[0], x[0, 0].numel()
            attn = ops.attention_map(ws, lse, n, wq.shape[0], s, ops.mma_mode(x.dtype))
            ctx.mark_non_differentiable(attn)
        if need:
            ctx.has_pre = pre is not None
            saved = [x, wq, bq, gamma, res[3], lse]
            if pre is not None:
                saved += [pre, mask, alpha]
            ctx.save_for_backward(*saved)
        return out, attn

    @staticmethod
    def backward(ctx, grad_out, grad_attn):
        x, wq, bq, gamma, o_saved, lse = ctx.saved_tensors[:6]
        pre = mask = alpha = None
        if ctx.has_pre:
            pre, mask, alpha = ctx.saved_tensors[6:]
        dq, dv0, dv1, da0, da1 = ops.attention_backward(x, wq, bq, x, pre, mask, gamma, 1.0, False, alpha, 0.0,
                                                        ctx.has_pre, o_saved, lse, grad_out, order=(0, 1))
        dw, dx, db = ops.qconv_backward(dq, x, wq, True)
        g_x = (dv0 + dx).to(x.dtype)
        return (g_x, dw.to(wq.dtype), db.to(bq.dtype), da0.reshape(1).to(gamma.dtype),
                dv1.to(pre.dtype) if ctx.has_pre else None, None,
                da1.reshape(1).to(alpha.dtype) if ctx.has_pre else None)


def _materialize() -> bool:
    return os.environ.get("FMI_MATERIALIZE_ATTN", "0") == "1"


class Auto_Attn(nn.Module):
    """modules/pluralistic_model/base_function.py:401-448 (Short+Long attention layer).

    `forward` returns `(out, attention)` like the reference; the S x S `attention` map (1 GiB fp32 per image at
    128^2, discarded by both callers: network.py:268,365) is `None` unless FMI_MATERIALIZE_ATTN=1.
    `resblock` builds the `model` sub-block used by the `pre` branch (:413-418, :446); when the package is
    installed over the reference (patch.py) it is the reference's own ResBlock (out-of-scope conv block)."""

    resblock_factory = None  # set by patch.install(); signature (in_nc, out_nc, hidden_nc, norm_layer) -> nn.Module

    def __init__(self, input_nc, norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.input_nc = input_nc
        self.query_conv = nn.Conv2d(input_nc, input_nc // 4, kernel_size=1)  # parameter holder (:408)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.alpha = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)  # kept for attribute parity; unused
        factory = type(self).resblock_factory
        if factory is None:
            from .picnet_blocks import ResBlock
            factory = lambda i, o, h, nl: ResBlock(i, o, h, norm_layer=nl, use_spect=True)  # noqa: E731
        self.model = factory(int(input_nc * 2), input_nc, input_nc, norm_layer)

    def forward(self, x, pre=None, mask=None):
        if pre is None:
            out, attention = _AutoAttnFunction.apply(x, self.query_conv.weight, self.query_conv.bias, self.gamma,
                                                     None, None, None)
            return out, attention
        m = mask.expand(x.shape[0], 1, *x.shape[2:]) if mask.dim() == 4 else mask
        cat, attention = _AutoAttnFunction.apply(x, self.query_conv.weight, self.query_conv.bias, self.gamma, pre, m,
                                                 self.alpha)
        out = self.model(cat)  # ResBlock(cat[out, context_flow]) (:446)
        return out, attention
