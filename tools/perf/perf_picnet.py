"""PICNet-ref generator forward, 256x256 inputs (BASELINE config 1) — NOT a pytest file; run on the GPU box:

    python tools/perf/perf_picnet.py > gpurun_out/perf_picnet.txt

`ours`   : modules/picnet.py::ReferenceFill — ExampleGuidedAttention @32^2, Auto_Attn @128^2, mask scaling and (when TF32
           convolutions are allowed or FMI_PRECISION=bf16) the encoder / decoder conv blocks on the sm_100a kernels; with
           cudnn_tf32=0 and the fp32 contract the conv blocks are strict-fp32 cuDNN as in the reference.
`ref-GPU`: the SAME network and weights with the two attention modules computed the reference's way (oracle functions =
           the reference's bmm / softmax / bmm formulation with the S x S map materialised, modules/example_guided_att.py:15-41,
           base_function.py:420-448) and every conv block on cuDNN (FMI_PICNET_CUDNN=1) — what the unmodified reference
           executes on this GPU.
`ref-CPU`: that formulation on the host cores (the reference's CPU path), batch 4, one forward.
"""
import copy
import ctypes
import os
import sys
import time
import types
from pathlib import Path

import torch
from torch import nn

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from face_mask_inpaint_b200 import _lib  # noqa: E402
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref  # noqa: E402
from golden_util import fill_by_name, mean_z, picnet_inputs  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


class TorchEGA(nn.Module):
    def __init__(self, mod):
        super().__init__()
        self.conv = mod.conv

    def forward(self, mask, src, ref):
        return O.example_guided_attention(mask, src, ref, self.conv.weight)


class TorchAutoAttn(nn.Module):
    def __init__(self, mod):
        super().__init__()
        self.query_conv, self.gamma = mod.query_conv, mod.gamma

    def forward(self, x, pre=None, mask=None):
        return O.auto_attn(x, self.query_conv.weight, self.query_conv.bias, self.gamma)[0], None


def time_cuda(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    lib = _lib.load()
    base = fill_by_name(build_picnet_ref()).eval()
    base.decoder.get_z = types.MethodType(mean_z, base.decoder)
    ours = copy.deepcopy(base).cuda()
    ours.decoder.get_z = types.MethodType(mean_z, ours.decoder)
    refm = copy.deepcopy(base)
    refm.attention = TorchEGA(refm.attention)
    refm.decoder.attn1 = TorchAutoAttn(refm.decoder.attn1)
    refm.decoder.get_z = types.MethodType(mean_z, refm.decoder)
    refm = refm.cuda()
    gflop_img = 296.3  # 123.9 conv + 171.8 Auto_Attn + 0.6 EGA
    print(f"{'case':58s} {'ms/batch':>10s} {'img/s':>9s} {'TFLOP/s':>8s}")
    with torch.no_grad():
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False  # the reference's bmm default
            for batch in (1, 4, 8):
                src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
                for name, model, prec in (("ours fp32 contract", ours, "fp32"), ("ours bf16 attention operands", ours, "bf16"),
                                          ("ref-GPU (bmm+softmax, S x S materialised)", refm, None)):
                    if prec:
                        os.environ["FMI_PRECISION"] = prec
                    if model is refm and batch * 16384 * 16384 * 4 * 2 > 100e9:
                        print(f"{name} B={batch}: skipped (S x S maps would need {batch * 2} GiB+)")
                        continue
                    def fn(model=model):
                        # the reference arm keeps its conv blocks on cuDNN (the kernel path would otherwise serve it too)
                        os.environ["FMI_PICNET_CUDNN"] = "1" if model is refm else "0"
                        return model(src, ref, mask)
                    t = time_cuda(fn)
                    extra = ""
                    if model is ours:
                        lib.fmi_profile_enable(1)
                        fn()
                        torch.cuda.synchronize()
                        lib.fmi_profile_enable(0)
                        tot, n = ctypes.c_double(0), ctypes.c_int(0)
                        lib.fmi_profile_collect(0, ctypes.byref(tot), ctypes.byref(n))
                        extra = f"  | attention main kernels {tot.value:6.2f} ms in {n.value} launches"
                    print(f"{name + f' B={batch} cudnn_tf32={int(tf32)}':58s} {t:10.2f} {batch / t * 1e3:9.1f} "
                          f"{gflop_img * batch / t:8.1f}{extra}", flush=True)
                os.environ.pop("FMI_PRECISION", None)
        # bf16 autocast + channels_last for the cuDNN conv blocks (library tuning of the out-of-scope part), bf16 attention
        torch.backends.cudnn.allow_tf32 = True
        ours_cl = copy.deepcopy(ours)   # (SpectralNorm views its 4-D weight as a matrix: the parameters stay NCHW)
        ours_cl.decoder.get_z = types.MethodType(mean_z, ours_cl.decoder)
        ref_cl = copy.deepcopy(refm)
        ref_cl.decoder.get_z = types.MethodType(mean_z, ref_cl.decoder)
        for batch in (4, 8):
            src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
            for name, model in (("ours, bf16 autocast, channels_last activations", ours_cl),
                                ("ref-GPU, bf16 autocast, channels_last activations", ref_cl)):
                def fn(model=model):
                    os.environ["FMI_PICNET_CUDNN"] = "1" if model is ref_cl else "0"
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return model(src.contiguous(memory_format=torch.channels_last),
                                     ref.contiguous(memory_format=torch.channels_last), mask)
                try:
                    t = time_cuda(fn)
                    print(f"{name + f' B={batch}':58s} {t:10.2f} {batch / t * 1e3:9.1f} {gflop_img * batch / t:8.1f}", flush=True)
                except Exception as ex:  # noqa: BLE001
                    print(f"{name} B={batch}: failed: {type(ex).__name__}: {str(ex)[:200]}")
        # parity of the two GPU paths on the same weights and the same (fresh) SpectralNorm state (B=1)
        torch.backends.cudnn.allow_tf32 = False

        def fresh(torch_attention):
            m = copy.deepcopy(base)
            if torch_attention:
                m.attention = TorchEGA(m.attention)
                m.decoder.attn1 = TorchAutoAttn(m.decoder.attn1)
            m.decoder.get_z = types.MethodType(mean_z, m.decoder)
            return m

        src, ref, mask = (t.cuda() for t in picnet_inputs(1))
        os.environ["FMI_PICNET_CUDNN"] = "1"      # strict fp32 on both sides: only the attention formulation differs
        a = fresh(False).cuda()(src, ref, mask)
        b = fresh(True).cuda()(src, ref, mask)
        print(f"ours vs ref-GPU output image (same weights, fresh SpectralNorm state, B=1): rel err "
              f"{((a - b).abs().max() / b.abs().max()).item():.2e}")
        # the reference's CPU path on this box's host cores (mask scaling the reference's way: F.interpolate)
        torch.set_num_threads(os.cpu_count())
        cpu = fresh(True)

        def cpu_forward(src, ref, mask):
            sd, sf = cpu.src_encoder(src)
            rd, rf = cpu.ref_encoder(ref)
            enc = cpu.attention(O.scale_img(mask.unsqueeze(1), sf.shape[-2:]), sf, rf)
            return cpu.pool(cpu.decoder(enc, z=cpu.decoder.get_z(sd, rd)))

        src, ref, mask = picnet_inputs(4)
        cpu_forward(src, ref, mask)
        t0 = time.perf_counter()
        cpu_forward(src, ref, mask)
        dt = time.perf_counter() - t0
        print(f"ref-CPU B=4, {os.cpu_count()} host threads: {dt * 1e3:.0f} ms/batch = {4 / dt:.2f} img/s")


if __name__ == "__main__":
    main()
