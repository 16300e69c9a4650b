"""StyleGAN2-1024 decoder forward+backward (BASELINE config 5's decoder part: train_psp.py with train_decoder) — NOT a
pytest file; run on the GPU box:  python tools/perf/perf_stylegan2_train.py > gpurun_out/perf_stylegan2_train.txt

Times Generator.forward([codes], input_is_latent=True, randomize_noise=False) + image.backward() on our kernels (fp32
contract and bf16) and on the reference formulation (oracle functions under autograd on the same GPU: cuDNN grouped
conv / conv_transpose, ATen elementwise), per-GPU batch 2 (train_psp.sh's batch size) and 4.
"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import _lib  # noqa: E402
from face_mask_inpaint_b200.modules import stylegan2 as SG  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


def time_cuda(fn, warmup=2, iters=4):
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
    dev = "cuda"
    size = int(os.environ.get("SG_SIZE", "1024"))
    torch.manual_seed(0)
    gen = SG.Generator(size, 512, 8).to(dev).train()
    lib = _lib.load()
    for batch in (2, 4):
        latent = torch.randn(batch, gen.n_latent, 512, device=dev, requires_grad=True)
        gout = torch.randn(batch, 3, size, size, device=dev)
        grads = {}
        for mode in ("fp32", "bf16"):
            os.environ["FMI_PRECISION"] = mode

            def step():
                gen.zero_grad(set_to_none=True)
                latent.grad = None
                img, _ = gen([latent], input_is_latent=True, randomize_noise=False)
                img.backward(gout)

            n0 = lib.fmi_kernel_launch_count()
            step()
            launches = lib.fmi_kernel_launch_count() - n0
            t = time_cuda(step)
            grads[mode] = latent.grad.detach().clone()
            # 3x the forward FLOPs: dgrad and wgrad each cost one forward
            print(f"ours {mode}: {size}x{size} B={batch}  fwd+bwd {t:8.2f} ms  {batch / t * 1e3:7.1f} img/s  "
                  f"{3 * 148.5 * batch / t:6.1f} TFLOP/s (3x modconv fwd flops)  {launches} kernel launches/step",
                  flush=True)
        os.environ.pop("FMI_PRECISION")
        sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.startswith('noises.')
                                                  and not k.endswith('.kernel'))
              for k, v in gen.state_dict().items()}
        lat_r = latent.detach().clone().requires_grad_(True)

        def step_ref():
            for v in sd.values():
                v.grad = None
            lat_r.grad = None
            O.generator_synthesis(sd, lat_r).backward(gout)

        try:
            t_ref = time_cuda(step_ref, 1, 2)
            e32 = ((grads['fp32'] - lat_r.grad).abs().max() / lat_r.grad.abs().max()).item()
            e16 = ((grads['bf16'] - lat_r.grad).abs().max() / lat_r.grad.abs().max()).item()
            print(f"reference formulation (ATen/cuDNN autograd, TF32 convs allowed) B={batch}: fwd+bwd {t_ref:8.2f} ms  "
                  f"{batch / t_ref * 1e3:7.1f} img/s ; d(latent) ours vs it: fp32-contract {e32:.2e}, bf16 {e16:.2e} "
                  f"(includes leaky-ReLU sign flips)", flush=True)
        except torch.OutOfMemoryError as ex:
            print(f"reference formulation B={batch}: out of memory ({ex})")
        del sd, lat_r
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
