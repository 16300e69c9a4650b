"""Per-block comparison of the PICNet decoder kernel path (picnet_fast.decoder_forward) with the cuDNN path (fp32, TF32 off)
on the real shapes — NOT a pytest file:  python tools/debug/diag_picnet_blocks.py [fp32|bf16]
Each block of the kernel path is also re-run from the cuDNN path's input of that block, so the per-block error is separated
from the accumulated one."""
import copy
import os
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from face_mask_inpaint_b200.modules import picnet_fast as PF  # noqa: E402
from face_mask_inpaint_b200.modules.picnet import build_picnet_ref  # noqa: E402
from golden_util import fill_by_name, mean_z, picnet_inputs  # noqa: E402


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def main():
    os.environ["FMI_PRECISION"] = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    torch.backends.cudnn.allow_tf32 = False
    base = fill_by_name(build_picnet_ref()).eval()
    src, ref, mask = (t.cuda() for t in picnet_inputs(2))
    def run_cudnn(tf32):
        m = copy.deepcopy(base).cuda()
        m.decoder.get_z = types.MethodType(mean_z, m.decoder)
        ins, outs = {}, {}
        for i in range(5):
            def hook(mod, a, o, i=i):
                ins[i], outs[i] = a[0].clone(), o.clone()
            getattr(m.decoder, f"decoder{i}").register_forward_hook(hook)

        def attn_hook(mod, a, o):
            outs["attn"] = o[0].clone()
        m.decoder.attn1.register_forward_hook(attn_hook)
        os.environ["FMI_PICNET_CUDNN"] = "1"
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad():
            img = m(src, ref, mask, resize=False)
        torch.backends.cudnn.allow_tf32 = False
        os.environ["FMI_PICNET_CUDNN"] = "0"
        return img, ins, outs

    want, ins, outs = run_cudnn(False)
    # what the reference's own GPU execution does under PyTorch defaults: cuDNN convolutions with TF32 operands
    got_tf32, _, outs_tf32 = run_cudnn(True)
    print(f"cuDNN TF32 (torch default) vs cuDNN fp32: image rel err {rel(got_tf32, want):.3e}")
    for i in range(5):
        print(f"  cuDNN TF32 accumulated decoder{i} rel err {rel(outs_tf32[i], outs[i]):.3e}")
    print(f"  cuDNN TF32 accumulated decoder1+attn rel err {rel(outs_tf32['attn'], outs['attn']):.3e}")
    ours = copy.deepcopy(base).cuda()
    ours.decoder.get_z = types.MethodType(mean_z, ours.decoder)
    taps = {}
    orig = PF.decoder_forward
    PF.decoder_forward = lambda gen, x, f_e=None, mask=None, pool_to=None, z=None: orig(gen, x, f_e, mask, taps=taps, pool_to=pool_to, z=z)
    torch.backends.cudnn.allow_tf32 = True     # the kernel path is taken when TF32 convolutions are allowed
    with torch.no_grad():
        got = ours(src, ref, mask, resize=False)
    torch.backends.cudnn.allow_tf32 = False
    PF.decoder_forward = orig
    print(f"image: rel err {rel(got, want):.3e}")
    for key, t in taps.items():
        i = int(key[7])
        w = outs["attn"] if key.endswith("+attn") else outs[i]
        if key.endswith(":lrelu"):
            w = torch.nn.functional.leaky_relu(w, 0.1)
        print(f"accumulated  {key:18s} {tuple(t.shape)}  rel err {rel(t, w):.3e}   max|ref| {w.abs().max().item():.3f} "
              f"std {w.std().item():.3f}")
    # isolated: each block alone from the cuDNN path's input (fresh SpectralNorm state: one more power iteration than the
    # reference run -> compare with a cuDNN re-run of the same module state)
    for i in range(4):
        blk_k = copy.deepcopy(getattr(base.decoder, f"decoder{i}")).cuda()
        blk_c = copy.deepcopy(getattr(base.decoder, f"decoder{i}")).cuda()
        nxt = copy.deepcopy(getattr(base.decoder, f"decoder{i + 1}")).cuda()
        from face_mask_inpaint_b200.modules import picnet as P
        nxt_co = PF._plain(nxt.conv2).out_channels
        gen = types.SimpleNamespace(layers=2, use_attn=False, decoder0=blk_k, decoder1=nxt,
                                    out1=P.Output(nxt_co, 3, 3, None, torch.nn.LeakyReLU(0.1), True, False).cuda())
        t2 = {}
        with torch.no_grad():
            w = blk_c(ins[i])
            blk_t = copy.deepcopy(getattr(base.decoder, f"decoder{i}")).cuda()
            torch.backends.cudnn.allow_tf32 = True
            wt = blk_t(ins[i])
            torch.backends.cudnn.allow_tf32 = False
            print(f"isolated     decoder{i}  cuDNN TF32 rel err {rel(wt, w):.3e}")
            try:
                orig(gen, ins[i], taps=t2)
                print(f"isolated     decoder{i}  rel err {rel(t2['decoder0'], w):.3e}")
            except Exception as ex:  # noqa: BLE001
                print(f"isolated decoder{i}: {type(ex).__name__}: {str(ex)[:200]}")


if __name__ == "__main__":
    main()
