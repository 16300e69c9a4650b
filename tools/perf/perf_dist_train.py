"""RefpSp decoder training step at BASELINE config 5's scale, batch-sharded over N GPUs — NOT a pytest file:

    python tools/perf/perf_dist_train.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 \
        tools/perf/perf_dist_train.py                                      # N GPUs, NCCL

Every rank: StyleGAN2-1024 Generator (train_decoder: all 30.4 M parameters trainable) forward + backward on its own batch of 2
(train_psp.sh's per-GPU batch) through the modulated-conv / upfirdn2d / bias-act kernels and their backward kernels, the bucketed
NCCL gradient all-reduce launched from the gradient hooks (dist.GradientAllReducer, overlapped with the rest of backward), Adam
step. Timed with CUDA events between barriers, max over ranks; rank 0 prints one JSON line. The pSp encoder and the losses of
train_psp.py are the reference's own PyTorch (out of scope) and are not part of the step.
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
os.environ.setdefault("FMI_PRECISION", "bf16")
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
from face_mask_inpaint_b200 import dist as fdist  # noqa: E402
from face_mask_inpaint_b200.modules import stylegan2 as SG  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        rank, local_rank, world = fdist.init_from_env()
    else:
        rank = local_rank = 0
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    gen = SG.Generator(1024, 512, 8).to(dev).train()
    if world > 1:
        fdist.broadcast_module_state(gen)
    per_rank = int(os.environ.get("FMI_PER_RANK_BATCH", "2"))
    g = torch.Generator().manual_seed(100 + rank)
    latent = torch.randn(per_rank, gen.n_latent, 512, generator=g).to(dev)
    target = torch.randn(per_rank, 3, 1024, 1024, generator=g).to(dev)
    params = [p for p in gen.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-5)
    reducer = fdist.GradientAllReducer(params, bucket_bytes=32 << 20).attach(opt) if world > 1 else None

    def step():
        opt.zero_grad(set_to_none=True)
        img, _ = gen([latent], input_is_latent=True, randomize_noise=False)
        loss = (img - target).square().mean()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    for _ in range(3):
        step()
    iters = 10
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        nparam = sum(p.numel() for p in params)
        print(json.dumps({"what": "StyleGAN2-1024 decoder train step (fwd + bwd + gradient all-reduce + Adam), bf16 operands",
                          "n_gpus": world, "per_gpu_batch": per_rank, "ms_per_step": ms.item(),
                          "img_per_s": world * per_rank / (ms.item() * 1e-3), "trainable_params": nparam,
                          "allreduce_bytes_per_step": nparam * 4 if world > 1 else 0,
                          "buckets": len(reducer.buckets) if reducer else 0, "loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
