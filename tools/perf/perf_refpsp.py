"""RefpSp inference, 1024x1024 output (BASELINE config 3: pSp encoder + StyleGAN2 decoder with attention, batch 8) — NOT a
pytest file; run on the GPU box:  python tools/perf/perf_refpsp.py > gpurun_out/perf_refpsp.txt

`ours`   : modules/psp.py::pSp — decoder, both attention modules and the masked blends on the sm_100a kernels (bf16 operands
           = the configuration's precision, and the fp32 contract); IR-SE50 trunk + map2style heads on cuDNN.
`ref-GPU`: the same network and weights computed the reference's way on this GPU: StyleGAN2 decoder through the oracle's
           Generator.forward restatement (per-sample weights, grouped cuDNN convs, upfirdn2d / fused_leaky_relu as ATen ops),
           attention as bmm + softmax + bmm, blends as elementwise ops. Batch 2 (per-sample weights are memory hungry).
"""
import copy
import ctypes
import os
import sys
from pathlib import Path

import torch
from torch import nn

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from face_mask_inpaint_b200 import _lib  # noqa: E402
from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts  # noqa: E402
from golden_util import refpsp_inputs  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


class TorchEGA(nn.Module):
    def __init__(self, mod):
        super().__init__()
        self.conv, self.out_conv = mod.conv, getattr(mod, 'out_conv', None)

    def forward(self, mask, src, ref):
        oc = self.out_conv
        return O.example_guided_attention(mask, src, ref, self.conv.weight, oc.weight if oc is not None else None,
                                          oc.bias if oc is not None else None)


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
    torch.manual_seed(0)
    net = pSp(refpsp_opts(output_size=1024)).eval().cuda()
    torch.backends.cudnn.allow_tf32 = True
    print(f"{'case':64s} {'ms/batch':>10s} {'img/s':>9s}")
    with torch.no_grad():
        for batch in (2, 8):
            x, ref, mask = (t.cuda() for t in refpsp_inputs(batch))
            for prec in ("fp32", "bf16"):
                os.environ["FMI_PRECISION"] = prec
                fn = lambda: net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
                t = time_cuda(fn)
                enc = time_cuda(lambda: net.encoder(x, ref=ref, mask=mask))
                lib.fmi_profile_enable(1)
                fn()
                torch.cuda.synchronize()
                lib.fmi_profile_enable(0)
                tot, n = ctypes.c_double(0), ctypes.c_int(0)
                lib.fmi_profile_collect(1, ctypes.byref(tot), ctypes.byref(n))
                print(f"{f'ours {prec} operands B={batch}':64s} {t:10.2f} {batch / t * 1e3:9.1f}  | encoder (cuDNN trunk + our "
                      f"attention/blends) {enc:6.2f} ms, decoder {t - enc:6.2f} ms (implicit GEMMs {tot.value:5.2f} ms)", flush=True)
            os.environ.pop("FMI_PRECISION")
        # library tuning of the out-of-scope trunk: bf16 autocast (+ channels_last inputs) around the IR-SE50 encoder only
        os.environ["FMI_PRECISION"] = "bf16"
        for batch in (2, 8):
            x, ref, mask = (t.cuda() for t in refpsp_inputs(batch))
            for cl in (False, True):
                xx = x.contiguous(memory_format=torch.channels_last) if cl else x
                rr = ref.contiguous(memory_format=torch.channels_last) if cl else ref

                def fn():
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        codes = net.encoder(xx, ref=rr, mask=mask)
                    codes = codes.float() + net.latent_avg.cuda()
                    return net.face_pool(net.decoder([codes], input_is_latent=True, randomize_noise=False)[0])

                try:
                    t = time_cuda(fn)
                    print(f"{f'ours bf16 operands + bf16-autocast trunk (channels_last={int(cl)}) B={batch}':64s} {t:10.2f} "
                          f"{batch / t * 1e3:9.1f}", flush=True)
                except Exception as ex:  # noqa: BLE001
                    print(f"autocast trunk cl={cl} B={batch} failed: {type(ex).__name__}: {str(ex)[:200]}")
        os.environ.pop("FMI_PRECISION")
        # the reference formulation on this GPU, batch 2
        x, ref, mask = (t.cuda() for t in refpsp_inputs(2))
        refnet = copy.deepcopy(net)
        refnet.encoder.attention1 = TorchEGA(refnet.encoder.attention1)
        refnet.encoder.attention2 = TorchEGA(refnet.encoder.attention2)
        sd = {k: v.detach() for k, v in refnet.decoder.state_dict().items()}

        def ref_forward():
            enc = refnet.encoder
            (c1, c2, c3), (r1, r2, r3) = enc._trunk(x), enc._trunk(ref)     # two trunk passes, as the reference does
            m = mask.unsqueeze(1)
            m3, m2, m1 = (O.scale_img(m, t.shape[-2:]) for t in (r3, r2, r1))
            c3 = enc.attention1(m3, c3, r3)
            c2 = enc.attention2(m2, c2, r2)
            c1 = m1 * r1 + (1 - m1) * c1
            lat = [enc.styles[j](c3) for j in range(3)]
            p2 = enc._upsample_add(c3, enc.latlayer1(c2))
            lat += [enc.styles[j](p2) for j in range(3, 7)]
            p1 = enc._upsample_add(p2, enc.latlayer2(c1))
            lat += [enc.styles[j](p1) for j in range(7, enc.style_count)]
            codes = torch.stack(lat, dim=1) + refnet.latent_avg.cuda()
            return refnet.face_pool(O.generator_synthesis(sd, codes))

        want = ref_forward()
        t_ref = time_cuda(ref_forward, 1, 3)
        print(f"{'ref-GPU (reference formulation, cuDNN TF32 allowed) B=2':64s} {t_ref:10.2f} {2 / t_ref * 1e3:9.1f}")
        got = net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
        print(f"ours (fp32 contract) vs ref-GPU image, B=2: rel err {((got - want).abs().max() / want.abs().max()).item():.2e}")


if __name__ == "__main__":
    main()
