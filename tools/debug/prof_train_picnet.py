"""torch.profiler breakdown of one PICNet GAN train step over the installed drop-ins (bench.py's train_picnet step):
    python tools/debug/prof_train_picnet.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

R = bench._patched_reference()
from modules.loss import GANOptimizer  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(21)
G, D = R.reference_fill().to(dev), R.discriminator().to(dev)
optG, optD = torch.optim.Adam(G.parameters(), lr=1e-5), torch.optim.Adam(D.parameters(), lr=1e-5)
gan = GANOptimizer(optD, optG).to(dev)
G.train(); D.train()
src, ref, gt, mask = (t.to(dev) for t in bench.make_inputs("train_picnet", 4, 3000))


def step():
    gen = G(src, ref, src_mask=mask)
    return gan(D, src, gt, ref, gen, mask)


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU],
                            record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
# GPU kernels only, all of them that matter: name, total us, launches
from torch.autograd import DeviceType  # noqa: E402
kern = sorted((e for e in prof.key_averages() if e.device_type == DeviceType.CUDA), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in kern)
print(f"--- GPU kernels: {tot / 1e3:.2f} ms in {sum(e.count for e in kern)} launches")
for e in kern[:70]:
    print(f"{e.self_device_time_total / 1e3:9.3f} ms {e.count:5d}  {e.key[:110]}")
# where the layout copies come from: aten::copy_ / contiguous / clone by input shape
print("--- copies by input shape (device time of the op incl. children)")
byshape = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ("aten::copy_", "aten::contiguous", "aten::clone")]
for e in sorted(byshape, key=lambda e: -e.device_time_total)[:25]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {e.count:5d}  {e.key:18s} {str(e.input_shapes)[:120]}")
