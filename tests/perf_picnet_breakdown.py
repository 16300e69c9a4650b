"""Where a PICNet-ref forward spends its time, per block (CUDA events around each sub-module, eager launch, batch 8 and 4) —
NOT a pytest file:  python tests/perf_picnet_breakdown.py > gpurun_out/perf_picnet_breakdown.txt"""
import sys
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
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


if __name__ == "__main__":
    main()
