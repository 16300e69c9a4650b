"""Inference path of the PICNet conv blocks — decoder, encoders, z -> f block — on the sm_100a kernels (SURVEY §8f rank 1;
csrc/conv_blocks.cu).

`ResGenerator.forward` (modules/pluralistic_model/network.py:247-268) runs, per ResBlockDecoder
(base_function.py:308-366):  IN -> LeakyReLU -> Conv2d(3,1,1) -> IN -> LeakyReLU -> ConvTranspose2d(3,2,1,1), plus a
ConvTranspose2d(3,2,1,1) shortcut of the block input, every conv wrapped in SpectralNorm; then Auto_Attn after block 1
and Output (LeakyReLU -> ReflectionPad2d(1) -> Conv2d(3) -> Tanh, :369-398) after the last block. Here the activations stay
NHWC in the tensor-core operand type from the first block to the image, each block is 2 statistics passes, 2
normalise+activate passes and 2-5 implicit-GEMM launches (conv, then the transposed conv — one launch over its 4 merged
output-parity classes when O <= 32, else one per class — with main path and shortcut concatenated along the input channels so
that their sum is the accumulator), and every GEMM writes straight into the buffer the next consumer reads (channel slice of
the next block's [a2 | x] buffer, or the interior of the reflection-padded buffer of the Output kernel, which also does the
final 4x4 average pooling). `encoder_forward` (second half of this file) does the same for ResEncoder.

Used when autograd is off, the tensors are CUDA and TF32 convolutions are allowed (torch.backends.cudnn.allow_tf32, PyTorch's
default — i.e. whenever the reference itself would run these convolutions with TF32 operands) or FMI_PRECISION=bf16; training
(autograd on) goes layer by layer through the Functions of ops.py instead (forward, data gradient, weight gradient on the same
kernels; SpectralNorm.forward and the block forwards route there).
SpectralNorm (external_function.py:44-57) keeps its one power iteration per forward, u and v advanced in place: per conv
(fmi_conv_weight_prep_sn) on a network's first forward, then for all of its convs at once (`_WeightPlan`,
fmi_conv_weight_prep_sn_batch).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from .. import _lib, ops


class _SpectralNormMeta(type):
    """isinstance(x, SpectralNorm) for this package's SpectralNorm AND the reference's own class
    (external_function.py:16-72) when the package is installed over the reference: same attributes either way."""

    def __instancecheck__(cls, obj):
        m = getattr(obj, "module", None)
        return (m is not None and hasattr(obj, "power_iterations") and getattr(obj, "name", None) == "weight"
                and hasattr(m, "weight_bar") and hasattr(m, "weight_u") and hasattr(m, "weight_v"))


class SpectralNorm(metaclass=_SpectralNormMeta):
    pass


def _p(t):
    return None if t is None else t.data_ptr()


def _effective(conv):
    """(weight fp32 contiguous, bias) of a conv that may be wrapped in SpectralNorm (advances u/v like a forward would)."""
    if isinstance(conv, SpectralNorm):
        conv._update_u_v()          # sets module.weight = w_bar / sigma (a tensor, not a Parameter)
        conv = conv.module
    w = conv.weight
    b = conv.bias
    return w.detach().float().contiguous(), (None if b is None else b.detach().float().contiguous())


def _plain(conv):
    return conv.module if isinstance(conv, SpectralNorm) else conv


def _slope(act):
    if isinstance(act, nn.LeakyReLU):
        return float(act.negative_slope)
    if isinstance(act, nn.ReLU):
        return 0.0
    return None


def _block_layout(blk):
    """(norm1, act, norm2) of a ResBlockDecoder's `model` Sequential, or None if it is not IN/none + (Leaky)ReLU."""
    mods = list(blk.model)
    if len(mods) == 6:
        n1, a1, _, n2, a2, _ = mods
    elif len(mods) == 4:
        n1 = n2 = None
        a1, _, a2, _ = mods
    else:
        return None
    for n in (n1, n2):
        if n is not None and not (isinstance(n, nn.InstanceNorm2d) and not n.track_running_stats):
            return None
    if _slope(a1) is None or _slope(a2) is None:
        return None
    return n1, a1, n2


def supported(gen, x) -> bool:
    """True when `ResGenerator.forward` can take the kernel path for this call."""
    if os.environ.get("FMI_PICNET_CUDNN") == "1" or torch.is_grad_enabled() or not x.is_cuda or x.dtype != torch.float32:
        return False
    # Operand precision follows the switch that governs the reference's own GPU numerics for these convolutions: with
    # torch.backends.cudnn.allow_tf32 (PyTorch's default) cuDNN runs them with TF32 operands, and so do the kernels here
    # (measured on the README configuration, image vs strict fp32: cuDNN-TF32 1.4e-2, these kernels 1.0e-2). With TF32
    # switched off the caller asked for strict fp32 convolutions: those run with error-compensated 3xTF32 operands
    # (ops.tf32_split, fmi_tf32_split3: fp32-class products on the same GEMM kernel, K tripled) unless FMI_PRECISION pins
    # single-pass TF32, in which case they stay on cuDNN.
    if not torch.backends.cudnn.allow_tf32 and ops.mma_mode(torch.float32) != _lib.MMA_BF16 and not ops.tf32_split():
        return False
    cached = getattr(gen, "_fmi_fast_ok", None)
    if cached is None:
        ok = True
        for i in range(gen.layers):
            blk = getattr(gen, f"decoder{i}")
            ok = ok and _block_layout(blk) is not None
            c1, c2, bp = _plain(blk.conv1), _plain(blk.conv2), _plain(blk.bypass)
            ok = ok and isinstance(c1, nn.Conv2d) and isinstance(c2, nn.ConvTranspose2d) and isinstance(bp, nn.ConvTranspose2d)
            if ok:
                cin, ch, co = c1.in_channels, c1.out_channels, c2.out_channels
                ok = cin % 32 == 0 and ch % 32 == 0 and co % 32 == 0 and max(cin, ch) <= 1024 and (co <= 256 or co % 256 == 0) \
                    and (ch <= 256 or ch % 256 == 0)
        out = getattr(gen, f"out{gen.layers - 1}", None)
        if out is None or len(list(out.model)) != 4 or _slope(out.model[0]) is None or _plain(out.conv1).out_channels > 32:
            ok = False
        for i in range(gen.layers - 1):
            if hasattr(gen, f"out{i}"):
                ok = False
        if gen.use_attn and gen.layers < 3:   # attention after block 1 must not be the last block
            ok = False
        # the z -> f blocks (network.py:213-220) run on the kernels too when they are plain ResBlocks without norm
        zb = [getattr(gen, "generator", None)] + [getattr(gen, f"generator{i}", None) for i in range(getattr(gen, "L", 0))]
        gen._fmi_z_ok = all(bk is not None and _enc_block_layout(bk) is not None and _enc_block_layout(bk)[2] is False
                            and all(isinstance(_plain(c), nn.Conv2d) for c in (bk.conv1, bk.conv2, bk.bypass))
                            and _plain(bk.conv1).in_channels % 32 == 0 and _plain(bk.conv1).out_channels % 32 == 0
                            and _plain(bk.conv2).out_channels % 32 == 0 for bk in zb)
        gen._fmi_fast_ok = cached = ok
    return cached


class _WeightPlan:
    """All SpectralNorm power iterations + weight re-layouts of one network as three launches at the start of its forward
    (fmi_conv_weight_prep_sn_batch). The first forward runs them one convolution at a time and records, in call order, what
    `_Ctx.weights` was asked for; `finalize` then allocates persistent wp / scratch buffers and the device descriptor table.
    Later forwards call `run()` once and get the buffers back in the same order."""

    def __init__(self):
        self.entries, self.ok, self.ready, self.cursor = [], True, False, 0

    def record(self, parts, o_rows, merged, i_row, ksize, shape, dtype):
        if not all(isinstance(w, SpectralNorm) and w.power_iterations == 1 for w, _ in parts):
            self.ok = False
        self.entries.append((parts, o_rows, merged, i_row, ksize, shape, dtype))

    def finalize(self, dev, mma):
        import struct
        if not self.ok or not self.entries or os.environ.get("FMI_SN_TORCH") == "1" or os.environ.get("FMI_SN_BATCH") == "0":
            self.ok = False
            return
        self.wps, rows, scratch_floats = [], [], 0
        for parts, o_rows, merged, i_row, ksize, shape, dtype in self.entries:
            wp = torch.zeros(shape, dtype=dtype, device=dev)
            self.wps.append(wp)
            off = 0
            o_real = None
            for w, tr in parts:
                m = w.module
                wb = m.weight_bar.data
                o, i = (wb.shape[1], wb.shape[0]) if tr else (wb.shape[0], wb.shape[1])
                hh, wd = wb.shape[0], wb.numel() // wb.shape[0]
                rows.append([m, wp, o, i, int(tr), o_rows, i_row, off, int(merged), ksize * ksize, hh, wd, scratch_floats])
                scratch_floats += 4 * wd + hh
                off += i
        self.scratch = torch.empty(scratch_floats, dtype=torch.float32, device=dev)
        blob = b""
        self.ptrs = []
        for m, wp, o, i, tr, o_rows, i_row, off, merged, t, hh, wd, soff in rows:
            wb, u, v = m.weight_bar.data, m.weight_u.data, m.weight_v.data
            if not (wb.is_contiguous() and u.is_contiguous() and v.is_contiguous() and wb.dtype == torch.float32
                    and wb.device == wp.device):
                self.ok = False
                return
            base = self.scratch.data_ptr() + 4 * soff
            blob += struct.pack("6Q12i", wb.data_ptr(), u.data_ptr(), v.data_ptr(), base, base + 16 * wd, wp.data_ptr(), o, i, tr,
                                o_rows, i_row, off, merged, t, hh, wd, 0, 0)
            self.ptrs.append((m, wb.data_ptr(), u.data_ptr(), v.data_ptr()))
        self.table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        self.n = len(rows)
        self.max_wd = max(r[11] for r in rows)
        self.max_hh = max(r[10] for r in rows)
        self.max_elems = max(r[2] * r[3] * r[9] for r in rows)
        self.mma = mma
        self.ready = self.max_wd <= 12288

    def valid(self):
        return all(m.weight_bar.data_ptr() == a and m.weight_u.data_ptr() == b and m.weight_v.data_ptr() == c
                   for m, a, b, c in self.ptrs)

    def run(self, lib, st):
        _lib.check(lib.fmi_conv_weight_prep_sn_batch(self.table.data_ptr(), self.n, self.max_wd, self.max_hh, self.max_elems,
                                                     self.mma, st), "fmi_conv_weight_prep_sn_batch")
        self.cursor = 0

    def next(self):
        if self.cursor >= len(self.wps):
            raise RuntimeError("fmi_b200: weight plan out of step with the forward that recorded it")
        wp = self.wps[self.cursor]
        self.cursor += 1
        return wp


class _Ctx:
    def __init__(self, dev, owner=None, sig=None):
        self.lib = _lib.load()
        self.mma = ops.mma_mode(torch.float32)
        self.dt = torch.float32 if self.mma == _lib.MMA_TF32 else torch.bfloat16
        self.dev = dev
        self.st = ops._stream()
        # strict-fp32 contract: exact fp32 activations / weights, every GEMM over the split operands (fmi_tf32_split3)
        self.x3 = self.mma == _lib.MMA_TF32 and ops.tf32_split()
        self.lib.fmi_set_tf32_exact(int(self.x3))
        self.rnd = 0 if self.x3 else 1
        # weight plan of the owning network (per operand type and device); None = prepare weights call by call
        self.plan = self.recording = None
        if owner is not None:
            from ..graphs import module_cache
            plans = module_cache(owner).setdefault("weight_plans", {})
            key = (self.mma, dev.index, sig, self.x3)   # sig: which convolutions this call pattern uses (e.g. with / without z)
            plan = plans.get(key)
            if plan is not None and plan.ready and not plan.valid():    # parameters were moved / replaced: rebuild
                plan = None
            if plan is None:
                plan = plans[key] = _WeightPlan()
                self.recording = plan
            elif plan.ready:
                plan.run(self.lib, self.st)
                self.plan = plan

    def finish(self):
        """End of the owner's forward: turn the recorded weight requests into the batched plan."""
        if self.recording is not None:
            self.recording.finalize(self.dev, self.mma)
            self.recording = None
        if self.x3:
            self.lib.fmi_set_tf32_exact(0)

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.dt, device=self.dev)

    def weights(self, parts, o_rows, merged=False, i_row=None):
        """`_weights` and, under the strict-fp32 contract, its [hi | lo | hi] split along the input channels."""
        wp = self._weights(parts, o_rows, merged, i_row)
        if not self.x3:
            return wp
        t, rows, ic = wp.shape
        wp3 = torch.empty((t, rows, 3 * ic), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.fmi_tf32_split3(wp.data_ptr(), ic, wp3.data_ptr(), t * rows, ic, 1, self.st), "fmi_tf32_split3")
        return wp3

    def _weights(self, parts, o_rows, merged=False, i_row=None):
        """parts: [(weight tensor or conv module (possibly SpectralNorm-wrapped), transposed)] concatenated along the input
        channels -> wp [9][o_rows][sum I], or with `merged` (transposed convs, fmi_conv3x3_nhwc mode 3)
        [4 input shifts][4 parity classes * O][sum I]. A SpectralNorm conv gets its power iteration here
        (fmi_conv_weight_prep_sn: u, v advanced in place, wp = w_bar / sigma)."""
        shape = lambda w: (_plain(w).weight_bar if isinstance(w, SpectralNorm) else (w if torch.is_tensor(w) else w.weight)).shape
        itot = sum((shape(w)[0] if tr else shape(w)[1]) for w, tr in parts)
        o_real = shape(parts[0][0])[1] if parts[0][1] else shape(parts[0][0])[0]
        if merged:
            o_rows = 4 * o_real
        ksize = shape(parts[0][0])[-1]
        i_row = max(itot, i_row or 0)       # the input tensor may carry zero-padded channels (3-channel images)
        if self.plan is not None:
            return self.plan.next()         # prepared by fmi_conv_weight_prep_sn_batch at the start of this forward
        wshape = (4 if merged else ksize * ksize, o_rows, i_row)
        if self.recording is not None:
            self.recording.record(parts, o_rows, merged, i_row, ksize, wshape, self.dt)
        wp = (torch.zeros if (merged or o_rows != o_real or i_row != itot) else torch.empty)(wshape, dtype=self.dt,
                                                                                            device=self.dev)
        itot = i_row
        off = 0
        for w, tr in parts:
            i = shape(w)[0] if tr else shape(w)[1]
            if isinstance(w, SpectralNorm) and w.power_iterations == 1 and os.environ.get("FMI_SN_TORCH") != "1":
                m = w.module
                wb, u, v = m.weight_bar.data, m.weight_u.data, m.weight_v.data
                if not (wb.is_contiguous() and u.is_contiguous() and v.is_contiguous() and wb.dtype == torch.float32):
                    raise RuntimeError("fmi_b200: SpectralNorm parameters must be contiguous fp32")
                scratch = torch.empty(u.numel() + v.numel(), dtype=torch.float32, device=self.dev)
                _lib.check(self.lib.fmi_conv_weight_prep_sn(_p(wb), _p(u), _p(v), _p(scratch), _p(wp), o_real, i, int(tr), o_rows,
                                                            itot, off, int(merged), ksize, self.mma, self.st), "fmi_conv_weight_prep_sn")
                off += i
                continue
            if not torch.is_tensor(w):
                w = _effective(w)[0]
            _lib.check(self.lib.fmi_conv_weight_prep(_p(w), _p(wp), o_real, i, int(tr), o_rows, itot, off, int(merged), ksize, self.mma,
                                                     self.st), "fmi_conv_weight_prep")
            off += i
        return wp

    def norm_act(self, x, x_stride, y, y_stride, norm, b, c, hw, slope):
        ss = None
        if norm is not None:
            ss = torch.empty((b, c, 2), dtype=torch.float32, device=self.dev)
            sums = torch.empty((b, c, 2), dtype=torch.float64, device=self.dev)
            g = None if norm.weight is None else norm.weight.detach().float().contiguous()
            be = None if norm.bias is None else norm.bias.detach().float().contiguous()
            _lib.check(self.lib.fmi_instnorm_stats_nhwc(x, x_stride, _p(g), _p(be), _p(ss), _p(sums), b, c, hw, float(norm.eps),
                                                        self.mma, self.st), "fmi_instnorm_stats_nhwc")
        _lib.check(self.lib.fmi_norm_act_nhwc(x, x_stride, y, y_stride, _p(ss), b, c, hw, slope, self.mma, self.st),
                   "fmi_norm_act_nhwc")
        return ss

    def conv(self, x, x_stride, wp, bias, y, y_stride, y_pad, y_nchw, nchw_c, b, i, o, h, w, mode, act, slope=0.0, round_y=1):
        if self.x3:     # activations [hi | hi | lo] against weights [hi | lo | hi]: three exact-product terms in one GEMM
            rows = b * ((h + 2) * (w + 2) if mode == 1 else h * w)
            xs = torch.empty((rows, 3 * i), dtype=torch.float32, device=self.dev)
            _lib.check(self.lib.fmi_tf32_split3(x, x_stride, xs.data_ptr(), rows, i, 0, self.st), "fmi_tf32_split3")
            x, x_stride, i, round_y = xs.data_ptr(), 3 * i, 3 * i, 0
        _lib.check(self.lib.fmi_conv3x3_nhwc(x, x_stride, _p(wp), _p(bias), y, y_stride, y_pad, _p(y_nchw), nchw_c, b, i, o, h, w,
                                             mode, act, slope, round_y, self.mma, self.st), "fmi_conv3x3_nhwc")


def decoder_forward(gen, x, f_e=None, mask=None, taps=None, pool_to=None, z=None):
    """The decoder loop of ResGenerator.forward (network.py:256-268) for `x` = encoded (+ f) [B, C, H, W] fp32 NCHW.
    Returns the image [B, 3, H * 2^layers, W * 2^layers] fp32. `taps` (diagnostics, tools/debug/diag_picnet_blocks.py): a dict that
    receives an fp32 NCHW copy of every block output. `pool_to` = (h, w): return AdaptiveAvgPool2d(pool_to) of the image
    instead (modules/model.py:111), fused into the Output kernel when it is an exact 4x4 mean. `z` [B, z_nc, H, W]: the
    latent of network.py:249-254 — f = generator(z) (+ generator{i}) is computed here and added to `x` (the residual-sum
    epilogue of f's last convolutions accumulates straight onto x in the first block's buffer)."""
    k = _Ctx(x.device, owner=gen, sig=z is not None)
    esz = 4 if k.mma == _lib.MMA_TF32 else 2
    b, c_in, h, w = x.shape
    x = x.contiguous()
    blocks = [getattr(gen, f"decoder{i}") for i in range(gen.layers)]
    ch0 = _plain(blocks[0].conv1).out_channels
    cat = k.empty(b, h, w, ch0 + c_in)            # [a2 | x] of the first block
    # block inputs / conv1 outputs feed InstanceNorm: stored exact (round_y = 0), see fmi_conv3x3_nhwc
    _lib.check(k.lib.fmi_nchw_to_nhwc_slice(_p(x), cat.data_ptr() + ch0 * esz, b, c_in, h, w, ch0 + c_in, _lib.F32, 0, k.mma, k.st),
               "fmi_nchw_to_nhwc_slice")
    if z is not None:
        zblocks = [gen.generator] + [getattr(gen, f"generator{i}") for i in range(gen.L)]
        zc = z.shape[1]
        zx = k.empty(b, h, w, zc)
        _lib.check(k.lib.fmi_nchw_to_nhwc_slice(_p(z.contiguous()), zx.data_ptr(), b, zc, h, w, zc, _lib.F32, k.rnd, k.mma, k.st),
                   "fmi_nchw_to_nhwc_slice")
        for n, blk in enumerate(zblocks):
            if n == len(zblocks) - 1:   # x += bypass(zx) + conv2(lrelu(conv1(lrelu(zx)))): both GEMMs accumulate onto x
                _res_block(k, blk, zx, zc, b, h, w, onto=(cat.data_ptr() + ch0 * esz, ch0 + c_in))
            else:
                zx, zc, _, _ = _res_block(k, blk, zx, zc, b, h, w)
    image = None
    for i, blk in enumerate(blocks):
        n1, act, n2 = _block_layout(blk)
        slope = _slope(act)
        c1, c2, cs = _plain(blk.conv1), _plain(blk.conv2), _plain(blk.bypass)
        b1, b2, bs = (None if c.bias is None else c.bias.detach().float().contiguous() for c in (c1, c2, cs))
        ch, co = c1.out_channels, c2.out_channels
        if c1.in_channels != c_in or cs.in_channels != c_in or c2.in_channels != ch or cs.out_channels != co:
            raise RuntimeError("fmi_b200: decoder block channel mismatch")
        ctot, hw = ch + c_in, h * w
        x_ptr = cat.data_ptr() + ch * esz
        # a1 = lrelu(IN(x)); h1 = conv1(a1) + b1
        a1 = k.empty(b, h, w, c_in)
        k.norm_act(x_ptr, ctot, a1.data_ptr(), c_in, n1, b, c_in, hw, slope)
        h1 = k.empty(b, h, w, ch)
        k.conv(a1.data_ptr(), c_in, k.weights([(blk.conv1, False)], ch), b1, h1.data_ptr(), ch, 0, None, 0, b, c_in, ch, h, w, 0, 2,
               round_y=0)
        del a1
        # a2 = lrelu(IN(h1)) into channels [0, ch) next to x
        k.norm_act(h1.data_ptr(), ch, cat.data_ptr(), ctot, n2, b, ch, hw, slope)
        del h1
        # y = convT(a2) + b2 + convT_shortcut(x) + bs: one GEMM over [a2 | x]
        bias = b2 if bs is None else (bs if b2 is None else b2 + bs)
        # O <= 32: one GEMM for the 4 parity classes (mode 3: 4/9 of the MMA instructions; measured 0.79 vs 0.95 ms on the last block).
        # O = 64 (N = 256, one CTA per SM) measured slower than the 4 per-class launches (0.84 vs 0.60 ms): mode 2 there
        up_mode = 3 if (co <= int(os.environ.get("FMI_CONVT_MERGE_MAX", "32"))) else 2
        wcat = k.weights([(blk.conv2, True), (blk.bypass, True)], co, merged=up_mode == 3)
        oh, ow = 2 * h, 2 * w
        last = i == gen.layers - 1
        attn = getattr(gen, f"attn{i}", None) if (i == 1 and gen.use_attn) else None
        if last:
            out_blk = getattr(gen, f"out{i}")
            if _plain(out_blk.conv1).in_channels != co:
                raise RuntimeError("fmi_b200: Output block channel mismatch")
            o_slope = _slope(out_blk.model[0])
            padded = k.empty(b, oh + 2, ow + 2, co)
            # Output's activation fused into the epilogue (the raw block output has no other reader: network.py:266-268)
            k.conv(cat.data_ptr(), ctot, wcat, bias, padded.data_ptr(), co, 1, None, 0, b, ctot, co, h, w, up_mode, 1, o_slope)
            _lib.check(k.lib.fmi_reflect_border_nhwc(padded.data_ptr(), b, co, oh, ow, k.mma, k.st), "fmi_reflect_border_nhwc")
            if taps is not None:
                taps[f"decoder{i}:lrelu"] = padded[:, 1:-1, 1:-1].float().permute(0, 3, 1, 2).contiguous()
            wo, bo = _effective(out_blk.conv1)
            n_img = wo.shape[0]
            fuse_pool = pool_to is not None and tuple(pool_to) == (oh // 4, ow // 4) and oh % 4 == 0 and ow % 4 == 0
            if co in (16, 32, 64) and n_img <= 3 and os.environ.get("FMI_OUTCONV_GEMM") != "1":
                # SIMT kernel (HBM-bound layer; fp32 accumulation), optionally with the 4x4 average pooling fused
                scratch = torch.empty(27 * co + 4, dtype=torch.float32, device=x.device)
                pooled = torch.empty((b, n_img, oh // 4, ow // 4), dtype=torch.float32, device=x.device) if fuse_pool else None
                image = None if fuse_pool else torch.empty((b, n_img, oh, ow), dtype=torch.float32, device=x.device)
                _lib.check(k.lib.fmi_output_conv_tanh(padded.data_ptr(), _p(wo), _p(bo), _p(image), _p(pooled), _p(scratch), b, co,
                                                      n_img, oh, ow, k.mma, k.st), "fmi_output_conv_tanh")
                if fuse_pool:
                    k.finish()
                    return pooled
            else:
                bo_pad = None
                if bo is not None:
                    bo_pad = torch.zeros(32, dtype=torch.float32, device=x.device)
                    bo_pad[:n_img] = bo
                image = torch.empty((b, n_img, oh, ow), dtype=torch.float32, device=x.device)
                k.conv(padded.data_ptr(), co, k.weights([(wo, False)], 32), bo_pad, None, 32, 0, image, n_img, b, co, 32, oh, ow, 1,
                       3)
        elif attn is not None:
            y = k.empty(b, oh, ow, co)
            k.conv(cat.data_ptr(), ctot, wcat, bias, y.data_ptr(), co, 0, None, 0, b, ctot, co, h, w, up_mode, 2, round_y=0)
            y_nchw = torch.empty((b, co, oh, ow), dtype=torch.float32, device=x.device)
            _lib.check(k.lib.fmi_nhwc_to_nchw(y.data_ptr(), _p(y_nchw), b, co, oh, ow, k.mma, _lib.F32, k.st), "fmi_nhwc_to_nchw")
            del y
            y_nchw, _ = attn(y_nchw, f_e, mask)
            y_nchw = y_nchw.contiguous()
            ch_next = _plain(blocks[i + 1].conv1).out_channels
            nxt = k.empty(b, oh, ow, ch_next + co)
            _lib.check(k.lib.fmi_nchw_to_nhwc_slice(_p(y_nchw), nxt.data_ptr() + ch_next * esz, b, co, oh, ow, ch_next + co,
                                                    _lib.F32, 0, k.mma, k.st), "fmi_nchw_to_nhwc_slice")
            cat = nxt
            if taps is not None:
                taps[f"decoder{i}+attn"] = y_nchw
        else:
            ch_next = _plain(blocks[i + 1].conv1).out_channels
            nxt = k.empty(b, oh, ow, ch_next + co)
            k.conv(cat.data_ptr(), ctot, wcat, bias, nxt.data_ptr() + ch_next * esz, ch_next + co, 0, None, 0, b, ctot, co, h, w,
                   up_mode, 2, round_y=0)
            cat = nxt
            if taps is not None:
                taps[f"decoder{i}"] = nxt[..., ch_next:].float().permute(0, 3, 1, 2).contiguous()
        c_in, h, w = co, oh, ow
    k.finish()
    if pool_to is not None:
        image = torch.nn.functional.adaptive_avg_pool2d(image, pool_to)
    return image


# ------------------------------------------------------------------------------------------------------------------------
# ResEncoder (network.py:73-172) on the same kernels: ResBlockEncoderOptimized / ResBlock with norm 'none'
# (base_function.py:207-305). Per block:  a2 = lrelu(conv1(a1) + b1) (activation fused into the GEMM epilogue: conv1's raw
# output has no other reader), y = bypass1x1(x) + bs, y += conv2(a2) + b2 (residual sum in the epilogue), then AvgPool2d(2)
# of the SUM for the 'down' blocks and the first block — pooling is linear and commutes with the 1x1 shortcut conv, so
# pool(main) + pool(shortcut) (:267) and pool(main) + bypass(pool(x)) (:300-303) are both pool(main + bypass(x)).
# ------------------------------------------------------------------------------------------------------------------------
def _enc_block_layout(blk):
    """(pre_activation, slope, pooled) or None. `model` is [act, conv1, act, conv2] (ResBlock, norm none) or
    [conv1, act, conv2, AvgPool2d] (ResBlockEncoderOptimized, norm none)."""
    mods = list(blk.model)
    pool = lambda m: isinstance(m, nn.AvgPool2d) and m.kernel_size in (2, (2, 2)) and m.stride in (2, (2, 2))
    if len(mods) == 4 and _slope(mods[0]) is not None and _slope(mods[2]) is not None:
        post = getattr(blk, "pool", None) if getattr(blk, "sample", False) else None
        if post is not None and not pool(post):
            return None
        if len(list(blk.shortcut)) != 1:
            return None
        return True, _slope(mods[0]), post is not None
    if len(mods) == 4 and _slope(mods[1]) is not None and pool(mods[3]):
        sc = list(blk.shortcut)
        if len(sc) != 2 or not pool(sc[0]):
            return None
        return False, _slope(mods[1]), True
    return None


def _enc_blocks(enc):
    blocks = [enc.block0] + [getattr(enc, f"encoder{i}") for i in range(enc.layers - 1)]
    if enc.ecnoder_type == 'src':
        heads = [getattr(enc, f"infer_prior{i}") for i in range(enc.L)] + [enc.prior]
    elif enc.ecnoder_type == 'ref':
        heads = [enc.posterior]
    else:
        return None, None
    return blocks, heads


def encoder_supported(enc, img) -> bool:
    if os.environ.get("FMI_PICNET_CUDNN") == "1" or torch.is_grad_enabled() or not img.is_cuda or img.dtype != torch.float32:
        return False
    if not torch.backends.cudnn.allow_tf32 and ops.mma_mode(torch.float32) != _lib.MMA_BF16 and not ops.tf32_split():
        return False                                                                             # see `supported`
    cached = getattr(enc, "_fmi_fast_ok", None)
    if cached is None:
        blocks, heads = _enc_blocks(enc)
        ok = blocks is not None
        for n, blk in enumerate((blocks or []) + (heads or [])):
            lay = _enc_block_layout(blk)
            c1, c2, bp = _plain(blk.conv1), _plain(blk.conv2), _plain(blk.bypass)
            ok = ok and lay is not None and all(isinstance(c, nn.Conv2d) for c in (c1, c2, bp))
            if ok:
                ok = (c1.kernel_size == (3, 3) and c2.kernel_size == (3, 3) and bp.kernel_size == (1, 1)
                      and (c1.in_channels % 32 == 0 or (n == 0 and c1.in_channels <= 32))
                      and all(c.out_channels % 32 == 0 and (c.out_channels <= 256 or c.out_channels % 256 == 0) for c in (c1, c2)))
        enc._fmi_fast_ok = cached = bool(ok)
    return cached


def _res_block(k, blk, x, cbuf, b, h, w, exact_out=False, onto=None):
    """One encoder-style residual block on x [B,h,w,cbuf] (operand type; channels beyond the conv's in_channels are zero).
    Returns (y [B,h',w',co], co, h', w'). `onto` = (pointer, pixel stride) of an NHWC tensor (slice) that already holds
    values: the block's result is ADDED to it (exact fp32 in TF32 mode) instead of being written to a new tensor."""
    pre_act, slope, pooled = _enc_block_layout(blk)
    c1, c2, bp = _plain(blk.conv1), _plain(blk.conv2), _plain(blk.bypass)
    b1, b2, bs = (None if c.bias is None else c.bias.detach().float().contiguous() for c in (c1, c2, bp))
    ch, co = c1.out_channels, c2.out_channels
    if c1.in_channels > cbuf or bp.in_channels != c1.in_channels or c2.in_channels != ch or bp.out_channels != co or \
            (c1.in_channels != cbuf and cbuf != 32):
        raise RuntimeError("fmi_b200: residual block channel mismatch")
    a1 = x
    if pre_act:
        a1 = k.empty(b, h, w, cbuf)
        k.norm_act(x.data_ptr(), cbuf, a1.data_ptr(), cbuf, None, b, cbuf, h * w, slope)
    a2 = k.empty(b, h, w, ch)
    k.conv(a1.data_ptr(), cbuf, k.weights([(blk.conv1, False)], ch, i_row=cbuf), b1, a2.data_ptr(), ch, 0, None, 0, b, cbuf, ch, h, w,
           0, 1, slope)
    if onto is not None:
        if pooled:
            raise RuntimeError("fmi_b200: a pooled block cannot accumulate onto an existing tensor")
        k.conv(x.data_ptr(), cbuf, k.weights([(blk.bypass, False)], co, i_row=cbuf), bs, onto[0], onto[1], 0, None, 0, b, cbuf, co,
               h, w, 4, 12, round_y=0)
        k.conv(a2.data_ptr(), ch, k.weights([(blk.conv2, False)], co), b2, onto[0], onto[1], 0, None, 0, b, ch, co, h, w, 0, 12,
               round_y=0)
        return None, co, h, w
    y = k.empty(b, h, w, co)
    k.conv(x.data_ptr(), cbuf, k.weights([(blk.bypass, False)], co, i_row=cbuf), bs, y.data_ptr(), co, 0, None, 0, b, cbuf, co, h, w,
           4, 2, round_y=0)
    k.conv(a2.data_ptr(), ch, k.weights([(blk.conv2, False)], co), b2, y.data_ptr(), co, 0, None, 0, b, ch, co, h, w, 0, 12,
           round_y=0 if (pooled or exact_out) else 1)
    if pooled:
        yp = k.empty(b, h // 2, w // 2, co)
        _lib.check(k.lib.fmi_avgpool2_nhwc(y.data_ptr(), co, yp.data_ptr(), co, b, co, h, w, 0 if (exact_out or k.x3) else 1, k.mma, k.st),
                   "fmi_avgpool2_nhwc")
        return yp, co, h // 2, w // 2
    return y, co, h, w


def encoder_forward(enc, img):
    """ResEncoder.forward (network.py:133-172): returns ([mu, softplus(std)], features) with NCHW fp32 tensors."""
    k = _Ctx(img.device, owner=enc)
    b, c_img, h, w = img.shape
    if h % (2 ** ((enc.layers + 1) // 2)) or w % (2 ** ((enc.layers + 1) // 2)):
        raise RuntimeError("fmi_b200: encoder input size must be divisible by its total down-sampling factor")
    blocks, heads = _enc_blocks(enc)
    cbuf = 32                                        # the image's channels zero-padded to one 128-byte fp32 row
    x = torch.zeros((b, h, w, cbuf), dtype=k.dt, device=img.device)
    _lib.check(k.lib.fmi_nchw_to_nhwc_slice(_p(img.contiguous()), x.data_ptr(), b, c_img, h, w, cbuf, _lib.F32, k.rnd, k.mma, k.st),
               "fmi_nchw_to_nhwc_slice")
    for n, blk in enumerate(blocks):
        x, cbuf, h, w = _res_block(k, blk, x, cbuf, b, h, w, exact_out=n == len(blocks) - 1)
    feats = torch.empty((b, cbuf, h, w), dtype=torch.float32, device=img.device)
    _lib.check(k.lib.fmi_nhwc_to_nchw(x.data_ptr(), _p(feats), b, cbuf, h, w, k.mma, _lib.F32, k.st), "fmi_nhwc_to_nchw")
    o = x
    for n, blk in enumerate(heads):
        o, co, _, _ = _res_block(k, blk, o, cbuf if n == 0 else co, b, h, w, exact_out=n == len(heads) - 1)
    dist = torch.empty((b, co, h, w), dtype=torch.float32, device=img.device)
    _lib.check(k.lib.fmi_nhwc_to_nchw(o.data_ptr(), _p(dist), b, co, h, w, k.mma, _lib.F32, k.st), "fmi_nhwc_to_nchw")
    k.finish()
    mu, std = torch.split(dist, enc.z_nc, dim=1)
    return [mu, torch.nn.functional.softplus(std)], feats
