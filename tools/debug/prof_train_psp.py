"""torch.profiler breakdown of one RefpSp train step over the installed drop-ins (bench.py's train_psp step): CPU vs GPU time and the
GPU kernels by time:  python tools/debug/prof_train_psp.py"""
import os
import sys
from argparse import Namespace
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
os.environ["FMI_PRECISION"] = "bf16"
import bench  # noqa: E402

R = bench._patched_reference()
from modules.psp.criteria import pSpLoss  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(31)
G = R.psp(output_size=1024, use_attention=1, train_decoder=1).to(dev)
G.latent_avg = G.latent_avg.to(dev)
params = list(G.encoder.parameters()) + list(G.decoder.parameters())
opt = torch.optim.Adam(params, lr=1e-5)
largs = Namespace(id_lambda=0, lpips_lambda=0.8, l2_lambda=1.0, style_lambda=250.0, lpips_lambda_ref=0, l2_lambda_ref=0, cx_lambda=1.0,
                  w_norm_lambda=0, start_from_latent_avg=1)
loss_fn = pSpLoss(largs).to(dev)
G.train()
src, ref, gt, mask = (t.to(dev) for t in bench.make_inputs("train_psp", 2, 4000))


def step():
    gen, latent = G(src, ref=ref, src_mask=mask, return_latents=True, randomize_noise=1)
    loss, _, _ = loss_fn(src, gt, gen, latent, latent_avg=G.latent_avg, ref=ref, mask=mask)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
from torch.autograd import DeviceType  # noqa: E402
ka = prof.key_averages()
kern = sorted((e for e in ka if e.device_type == DeviceType.CUDA), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in kern)
cpu = sum(e.self_cpu_time_total for e in ka)
print(f"--- CPU self time {cpu / 1e3:.1f} ms; GPU kernels: {tot / 1e3:.2f} ms in {sum(e.count for e in kern)} launches")
for e in kern[:45]:
    print(f"{e.self_device_time_total / 1e3:9.3f} ms {e.count:5d}  {e.key[:120]}")
