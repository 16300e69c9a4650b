"""Air-gapped runs of the reference's scripts (`FMI_OFFLINE=1 python -m face_mask_inpaint_b200.run <script> ...`).

The loss networks of the training scripts ask torchvision for pretrained weights (modules/loss.py:21 `vgg16(pretrained=True)`,
modules/psp/criteria/lpips/networks.py) which is a download. Without a network the same architectures are built with seeded
random weights instead — throughput runs and the script-level tests need the arithmetic, not the weights. Nothing here is on
the hot path."""
from __future__ import annotations

import sys
import zlib


def stub_pretrained() -> None:
    import torch
    import torchvision
    if getattr(torchvision.models, "_fmi_offline", False):
        return
    for name in ("vgg16", "alexnet", "squeezenet1_1", "inception_v3", "resnet50"):
        real = getattr(torchvision.models, name, None)
        if real is None:
            continue

        def make(real=real, name=name):
            def ctor(pretrained=False, progress=True, weights=None, **kw):   # `models.alexnet(True)` passes it positionally
                with torch.random.fork_rng(devices=[]):
                    torch.manual_seed(zlib.crc32(name.encode()) & 0xFFFF)
                    if name == "inception_v3":
                        kw.setdefault("init_weights", True)
                    return real(weights=None, **kw)
            return ctor
        setattr(torchvision.models, name, make())
    torchvision.models._fmi_offline = True
    # LPIPS linear heads (modules/psp/criteria/lpips/utils.py:11-20 downloads them): non-negative seeded stand-ins
    real_hub = torch.hub.load_state_dict_from_url
    lpips_channels = {"alex": (64, 192, 384, 256, 256), "vgg": (64, 128, 256, 512, 512), "squeeze": (64, 128, 256, 384, 384, 512, 512)}

    def load_state_dict_from_url(url, *a, **kw):
        for net, chans in lpips_channels.items():
            if "PerceptualSimilarity" in url and url.endswith(f"/{net}.pth"):
                g = torch.Generator().manual_seed(zlib.crc32(url.encode()) & 0xFFFF)
                return {f"lin{i}.model.1.weight": torch.rand(1, c, 1, 1, generator=g) / c for i, c in enumerate(chans)}
        return real_hub(url, *a, **kw)
    torch.hub.load_state_dict_from_url = load_state_dict_from_url
    sys.stderr.write("[fmi_b200] FMI_OFFLINE: torchvision pretrained weights replaced by seeded random weights\n")
