"""StyleGAN2-1024 decoder benchmark (BASELINE config 3, decoder only) — NOT a pytest file; run on the GPU box:

    python tools/perf/perf_stylegan2.py > gpurun_out/perf_stylegan2.txt

Times Generator.forward([codes], input_is_latent=True, randomize_noise=False) at batch 8 on our kernels (fp32 contract
and bf16), the per-kernel-class split from the library's CUDA-event hooks, and — beside it — the reference formulation
(oracle functions executed on the same GPU: cuDNN grouped conv / conv_transpose + ATen elementwise), batch 2.
"""
import ctypes
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from face_mask_inpaint_b200 import _lib  # noqa: E402
from face_mask_inpaint_b200.modules import stylegan2 as SG  # noqa: E402
from oracle import ref_ops as O  # noqa: E402


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
    dev = "cuda"
    size = int(os.environ.get("SG_SIZE", "1024"))
    batch = int(os.environ.get("SG_BATCH", "8"))
    torch.manual_seed(0)
    gen = SG.Generator(size, 512, 8).to(dev).eval()
    latent = torch.randn(batch, gen.n_latent, 512, device=dev)
    lib = _lib.load()
    gflop_img = 148.5 if size == 1024 else float("nan")
    with torch.no_grad():
        modes = ("fp32", "bf16") if not os.environ.get("SG_ONLY") else (os.environ["SG_ONLY"],)
        for mode in modes:
            os.environ["FMI_PRECISION"] = mode
            fn = lambda: gen([latent], input_is_latent=True, randomize_noise=False)[0]
            img = fn()
            t = time_cuda(fn)
            n0 = lib.fmi_kernel_launch_count()
            lib.fmi_profile_enable(1)
            fn()
            torch.cuda.synchronize()
            lib.fmi_profile_enable(0)
            tot, n = ctypes.c_double(0), ctypes.c_int(0)
            lib.fmi_profile_collect(1, ctypes.byref(tot), ctypes.byref(n))
            launches = lib.fmi_kernel_launch_count() - n0
            print(f"ours {mode}: {size}x{size} B={batch}  {t:8.2f} ms/forward  {batch / t * 1e3:8.1f} img/s  "
                  f"{gflop_img * batch / t:7.1f} TFLOP/s(modconv algorithmic)  | implicit-GEMM kernels {tot.value:7.2f} ms in "
                  f"{n.value} launches, {launches} launches total", flush=True)
            if mode == "fp32":
                img32 = img
        os.environ.pop("FMI_PRECISION")
        if os.environ.get("SG_ONLY"):
            return
        print(f"bf16 vs fp32-contract image: rel diff {((img.float() - img32).abs().max() / img32.abs().max()).item():.3e}")
        # reference formulation on the same GPU (smaller batch: per-sample weights + grouped conv are memory hungry)
        rb = 2
        sd = {k: v.detach() for k, v in gen.state_dict().items()}
        lat = latent[:rb]
        fn_ref = lambda: O.generator_synthesis(sd, lat)
        want = fn_ref()
        t_ref = time_cuda(fn_ref, 1, 3)
        got = gen([lat], input_is_latent=True, randomize_noise=False)[0]
        err = ((got - want).abs().max() / want.abs().max()).item()
        print(f"reference formulation on this GPU (cuDNN grouped convs, TF32 allowed): B={rb} {t_ref:8.2f} ms  "
              f"{rb / t_ref * 1e3:7.1f} img/s ; ours(fp32 contract) vs it: rel err {err:.3e}")


if __name__ == "__main__":
    main()
