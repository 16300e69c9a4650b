"""Whole-model forwards, eager launch vs one CUDA-graph replay (face_mask_inpaint_b200/graphs.py) — NOT a pytest file:

    python tools/perf/perf_graphs.py > gpurun_out/perf_graphs.txt

PICNet-ref 256^2 (BASELINE config 1; fp32 contract, cuDNN TF32 allowed) at batch 1 / 4 / 8 and RefpSp 1024^2 (config 3; bf16
operands) at batch 2 / 8, plus the bf16-autocast variants of the cuDNN trunks. Same kernels in both columns: the difference is
host launch time.
"""
import os
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from face_mask_inpaint_b200.graphs import CapturedForward  # noqa: E402
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref  # noqa: E402
from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts  # noqa: E402
from golden_util import fill_by_name, mean_z, picnet_inputs, refpsp_inputs  # noqa: E402


def time_cuda(fn, warmup=3, iters=10):
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


def row(name, batch, t_eager, t_graph):
    print(f"{name + f' B={batch}':62s} {t_eager:9.2f} {batch / t_eager * 1e3:9.1f} {t_graph:9.2f} {batch / t_graph * 1e3:9.1f}",
          flush=True)


def main():
    torch.backends.cudnn.allow_tf32 = True
    print(f"{'case':62s} {'eager ms':>9s} {'img/s':>9s} {'graph ms':>9s} {'img/s':>9s}")
    with torch.no_grad():
        net = fill_by_name(build_picnet_ref()).eval().cuda()
        for sampled in (False, True):
            if not sampled:
                net.decoder.get_z = types.MethodType(mean_z, net.decoder)
            else:
                del net.decoder.get_z      # back to the class method: rsample() inside the graph
            for batch in (1, 4, 8):
                src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
                for prec in ("fp32", "bf16"):
                    os.environ["FMI_PRECISION"] = prec
                    te = time_cuda(lambda: net(src, ref, mask))
                    fwd = CapturedForward(net, src, ref, mask)
                    tg = time_cuda(lambda: fwd(src, ref, mask))
                    row(f"PICNet-ref 256^2 {'z~N(mu,sigma)' if sampled else 'z=mu'}, attention operands {prec}", batch, te, tg)
                    del fwd
            os.environ.pop("FMI_PRECISION", None)

        def autocast_fwd(src, ref, mask):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return net(src.contiguous(memory_format=torch.channels_last),
                           ref.contiguous(memory_format=torch.channels_last), mask)

        os.environ["FMI_PRECISION"] = "bf16"
        for batch in (4, 8):
            src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
            te = time_cuda(lambda: autocast_fwd(src, ref, mask))
            fwd = CapturedForward(autocast_fwd, src, ref, mask)
            tg = time_cuda(lambda: fwd(src, ref, mask))
            row("PICNet-ref 256^2 bf16 autocast + channels_last conv blocks", batch, te, tg)
            del fwd
        del net
        torch.cuda.empty_cache()

        torch.manual_seed(0)
        net = pSp(refpsp_opts(output_size=1024)).eval().cuda()
        for batch in (2, 8):
            x, ref, mask = (t.cuda() for t in refpsp_inputs(batch))
            for prec in ("fp32", "bf16"):
                os.environ["FMI_PRECISION"] = prec
                te = time_cuda(lambda: net(x, ref=ref, src_mask=mask, resize=True, randomize_noise=False))
                fwd = CapturedForward(net, x, ref=ref, src_mask=mask, resize=True, randomize_noise=False)
                tg = time_cuda(lambda: fwd(x, ref=ref, src_mask=mask))
                row(f"RefpSp 1024^2 operands {prec}", batch, te, tg)
                del fwd
        os.environ["FMI_PRECISION"] = "bf16"

        def autocast_psp(x, ref, mask):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                codes = net.encoder(x.contiguous(memory_format=torch.channels_last),
                                    ref=ref.contiguous(memory_format=torch.channels_last), mask=mask)
            codes = codes.float() + net.latent_avg
            return net.face_pool(net.decoder([codes], input_is_latent=True, randomize_noise=False)[0])

        for batch in (2, 8):
            x, ref, mask = (t.cuda() for t in refpsp_inputs(batch))
            te = time_cuda(lambda: autocast_psp(x, ref, mask))
            fwd = CapturedForward(autocast_psp, x, ref, mask)
            tg = time_cuda(lambda: fwd(x, ref, mask))
            row("RefpSp 1024^2 bf16 operands + bf16-autocast channels_last trunk", batch, te, tg)
            del fwd


if __name__ == "__main__":
    main()
