"""Where a PICNet-ref forward spends its time — NOT a pytest file:  python tools/perf/perf_picnet_breakdown.py
First two lines: the forward split into CUDA graphs (encoders + EGA / decoder) = GPU times. Then per sub-module CUDA events in
eager mode (host-launch bound; the decoder blocks show 0 there because the kernel path replaces `ResGenerator.forward` as a
whole and does not call the block modules — set FMI_PICNET_CUDNN=1 to see the cuDNN blocks)."""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref  # noqa: E402
from golden_util import fill_by_name, mean_z, picnet_inputs  # noqa: E402


def main():
    torch.backends.cudnn.allow_tf32 = True
    net = fill_by_name(build_picnet_ref()).eval().cuda()
    net.decoder.get_z = types.MethodType(mean_z, net.decoder)
    names = ["src_encoder", "ref_encoder", "attention", "decoder.generator"] + [f"decoder.decoder{i}" for i in range(5)] + \
            ["decoder.attn1", "decoder.out4", "pool"]
    mods = dict(net.named_modules())
    events = {n: [] for n in names}
    for n in names:
        def pre(m, a, n=n):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events[n].append([e, None])

        def post(m, a, o, n=n):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            events[n][-1][1] = e
        mods[n].register_forward_pre_hook(pre)
        mods[n].register_forward_hook(post)
    for batch in (8, 4, 1):
        src, ref, mask = (t.cuda() for t in picnet_inputs(batch))
        with torch.no_grad():
            for _ in range(3):
                net(src, ref, mask)
            torch.cuda.synchronize()
            for n in names:
                events[n].clear()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                net(src, ref, mask)
            e1.record()
            torch.cuda.synchronize()
        print(f"batch {batch}: {e0.elapsed_time(e1) / 5:.2f} ms per forward")
        for n in names:
            ms = sum(a.elapsed_time(b) for a, b in events[n]) / 5
            print(f"  {n:22s} {ms:8.3f} ms")


def graphed_pieces():
    """The same forward split into three CUDA graphs (encoders + ExampleGuidedAttention, z -> f ResBlock, decoder): eager
    per-module times above are host-launch bound, these are the GPU times."""
    from face_mask_inpaint_b200 import ops
    from face_mask_inpaint_b200.graphs import CapturedForward
    torch.backends.cudnn.allow_tf32 = True
    net = fill_by_name(build_picnet_ref()).eval().cuda()
    net.decoder.get_z = types.MethodType(mean_z, net.decoder)

    def timed(fn, *a):
        g = CapturedForward(fn, *a)
        for _ in range(3):
            g(*a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g(*a)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10

    for batch in (8, 4):
        src, ref, mask = (t.cuda() for t in picnet_inputs(batch))

        def encoders(src, ref, mask):
            sd, sf = net.src_encoder(src)
            rd, rf = net.ref_encoder(ref)
            enc = net.attention(ops.scale_img(mask.unsqueeze(1), sf.shape[-2:]), sf, rf)
            return enc, sd[0], rd[0]

        with torch.no_grad():
            enc, smu, rmu = (t.clone() for t in encoders(src, ref, mask))
            z = torch.cat([smu, rmu], dim=1)
            f = net.decoder.generator(z).clone()
            x = (enc + f).clone()
        t_enc = timed(encoders, src, ref, mask)
        t_gen = timed(lambda z: net.decoder.generator(z), z)
        t_dec = timed(lambda x: net.decoder(x, pool_to=(256, 256)), x)
        t_all = timed(lambda a, b, c: net(a, b, c), src, ref, mask)
        print(f"graphed, batch {batch}: whole forward {t_all:.2f} ms | encoders + EGA {t_enc:.2f} | z->f ResBlock as a stand-alone cuDNN module {t_gen:.2f} (inside the forward it is part of the decoder graph, on the kernels) | "
              f"decoder blocks + Auto_Attn + Output + pool {t_dec:.2f}")


if __name__ == "__main__":
    graphed_pieces()
    main()
