"""BENCH / TEST INFRASTRUCTURE ONLY — the UNMODIFIED reference (syncdoth/face_mask_inpaint) as the other arm.

`baseline/_ref/` is a verbatim copy of the reference checkout made in the build container (`ensure_copy()`, called by
`__graft_entry__.build()`; git-ignored, NOT gpurun-ignored, so it travels to the GPU box like the built .so files). Nothing
under `face_mask_inpaint_b200/` imports this module. Users:
  * `bench.py --impl reference`     the reference's own modules on the host cores (ReferenceFill runs on a CPU as it stands);
  * `bench.py` key `gpu_reference`  the same unmodified modules on the same B200 (its own formulation on cuBLAS / cuDNN, its
                                    own two CUDA ops from oracle/_ref/*.so = the reference's four op sources compiled as
                                    they lie, oracle/build_ref.py);
  * `tests/test_scripts_gpu.py`     the four entry scripts under `python -m face_mask_inpaint_b200.run` on a fabricated
                                    dataset + random-init checkpoints (`fabricate_*` below).
Stubs, all outside the arithmetic being compared (there is no network and no pretrained file in the container):
`pytorch_msssim` (metric; the published algorithm, face_mask_inpaint_b200.patch), `torchvision.models.vgg16/alexnet
(pretrained=True)` -> the same architecture with seeded random weights, `lpips` weights likewise, the reference's import-time
JIT `torch.utils.cpp_extension.load` -> the prebuilt oracle/_ref modules (GPU) or the reference's own `upfirdn2d_native` plus
the leaky-ReLU restatement (CPU; SURVEY 8c — the reference has no CPU path for these two ops).
"""
from __future__ import annotations

import ast
import os
import shutil
import sys
import types
from argparse import Namespace
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
SRC = Path("/root/reference")

# SURVEY 8d, cfg 1 (README.md:58-70 + the argparse defaults of PICNet_inference.py:38-58)
PICNET_ENCODER = dict(type='pluralistic', ngf=32, z_nc=128, img_f=128, layers=5, norm='none', activation='LeakyReLU',
                      init_type='orthogonal')
PICNET_DECODER = dict(ngf=32, z_nc=256, img_f=256, L=0, layers=5, norm='instance', activation='LeakyReLU',
                      init_type='orthogonal')
PICNET_DISC = dict(ndf=32, layers=5, model_type='ResDis', init_type='orthogonal', img_f=128)


def ensure_copy() -> Path | None:
    """Copy /root/reference to baseline/_ref when the source tree is present (build container); the GPU box only has the
    copy. Returns the copy's path, or None when there is neither."""
    if SRC.exists():
        stamp = REF / ".copied_from"
        if not stamp.exists():
            if REF.exists():
                shutil.rmtree(REF)
            shutil.copytree(SRC, REF, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
            stamp.write_text(str(SRC))
    return REF if (REF / "modules" / "model.py").exists() else None


def available() -> bool:
    return (REF / "modules" / "model.py").exists()


def stub_pretrained() -> None:
    """torchvision's `pretrained=True` / `weights=...` would download: same architectures, seeded random weights."""
    sys.path.insert(0, str(ROOT))
    from face_mask_inpaint_b200.offline import stub_pretrained as stub
    stub()


def _native_upfirdn2d():
    """The reference's own CPU restatement (op/upfirdn2d.py:150-184) taken from its file by AST (importing the file would
    JIT-build the CUDA extension) with `F` injected — the file forgets to import it."""
    import torch
    import torch.nn.functional as F
    src = (REF / "modules/psp/stylegan2/op/upfirdn2d.py").read_text()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "upfirdn2d_native"][0]
    ns = {"F": F, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "upfirdn2d_native", "exec"), ns)
    return ns["upfirdn2d_native"]


def _install_cpu_ops() -> None:
    import torch
    import torch.nn.functional as F
    from torch import nn
    native = _native_upfirdn2d()

    def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
        n, c, h, w = input.shape
        out = native(input.reshape(-1, h, w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        return out.reshape(n, c, out.shape[1], out.shape[2])

    def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):   # op/fused_bias_act_kernel.cu:26-47
        return scale * F.leaky_relu(input + bias.view(1, -1, *([1] * (input.ndim - 2))), negative_slope)

    class FusedLeakyReLU(nn.Module):
        def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
            super().__init__()
            self.bias = nn.Parameter(torch.zeros(channel))
            self.negative_slope, self.scale = negative_slope, scale

        def forward(self, input):
            return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)

    pkg = types.ModuleType("modules.psp.stylegan2.op")
    pkg.upfirdn2d, pkg.fused_leaky_relu, pkg.FusedLeakyReLU = upfirdn2d, fused_leaky_relu, FusedLeakyReLU
    pkg.__path__ = []
    sys.modules["modules.psp.stylegan2.op"] = pkg


def _prebuilt_jit() -> None:
    """The reference builds its two extensions at import time (`load('fused', ...)`, `load('upfirdn2d', ...)`): hand it the
    same sources compiled ahead of time (oracle/_ref/*.so) instead of running ninja on the GPU box."""
    import torch.utils.cpp_extension as ce
    sys.path.insert(0, str(ROOT))
    from oracle import build_ref
    real = ce.load
    names = {"fused": "fmi_ref_fused", "upfirdn2d": "fmi_ref_upfirdn2d"}

    def load(name, *a, **kw):
        if name in names:
            return build_ref.load_built(names[name])
        return real(name, *a, **kw)
    ce.load = load


def import_unpatched(ops: str = "cpu"):
    """Put the reference copy first on sys.path with nothing of this package installed over it. `ops`: 'cpu' (restated
    StyleGAN2 ops, no GPU needed), 'cuda' (the reference's own compiled ops), 'none' (PICNet side only)."""
    if not available():
        raise FileNotFoundError(f"{REF} is missing: run __graft_entry__.build() in the build container")
    if "face_mask_inpaint_b200.patch" in sys.modules and sys.modules["face_mask_inpaint_b200.patch"]._INSTALLED:
        raise RuntimeError("the drop-ins are installed in this process: the unmodified reference needs its own process")
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    os.environ.setdefault("WANDB_MODE", "disabled")
    sys.path.insert(0, str(ROOT))
    from face_mask_inpaint_b200.patch import _ensure_msssim_shim
    _ensure_msssim_shim()
    stub_pretrained()
    if ops == "cpu":
        _install_cpu_ops()
    elif ops == "cuda":
        _prebuilt_jit()


def reference_fill(use_att=1):
    """The reference's ReferenceFill at BASELINE config 1 / SURVEY 8d."""
    from modules.model import ReferenceFill
    return ReferenceFill(None, dict(PICNET_ENCODER), dict(PICNET_DECODER), use_att=use_att)


def discriminator():
    from modules.pluralistic_model import network
    return network.define_d(**PICNET_DISC)


def psp_opts(output_size=1024, use_attention=1, train_decoder=0):
    return Namespace(output_size=output_size, encoder_type='GradualStyleEncoder', use_attention=use_attention,
                     train_decoder=train_decoder, start_from_latent_avg=1, learn_in_w=0, pt_ckpt_path=None,
                     stylegan_weights=None)


def psp(output_size=1024, use_attention=1, train_decoder=0):
    """The reference's pSp with random-init weights (its load_weights needs checkpoint files that are not shipped: bypassed,
    `latent_avg` = zeros — SURVEY 8d cfg 3)."""
    import torch
    from modules.psp import psp as psp_mod
    real = psp_mod.pSp.load_weights
    psp_mod.pSp.load_weights = lambda self: None
    try:
        net = psp_mod.pSp(psp_opts(output_size, use_attention, train_decoder))
    finally:
        psp_mod.pSp.load_weights = real
    net.latent_avg = torch.zeros(net.opts.n_styles, 512)
    return net


# ---------------------------------------------------------------------------------------------------------------------------
# fabricated dataset + checkpoints for the script-level tests (dataloader.py:122-266 layout)
# ---------------------------------------------------------------------------------------------------------------------------
def fabricate_dataset(root: Path, n_ids: int = 4, per_id: int = 2, size: int = 256, full: int = 1024, seed: int = 0) -> dict:
    """<root>/src/<id>_surgical.jpg, <root>/ref/<id>.jpg, <root>/mask/<id>.npy (uint8 0/1, same size as the source image),
    <root>/identity.txt ("<id>.jpg <identity>"). Images are `full` x `full` so that the scripts' own `img_scale` resizes them
    (PICNet: --img_scale 0.25 of 1024 -> 256; psp_inference hard-codes scale=0.25)."""
    import numpy as np
    from PIL import Image
    rng = np.random.default_rng(seed)
    root = Path(root)
    for d in ("src", "ref", "mask"):
        (root / d).mkdir(parents=True, exist_ok=True)
    lines = []
    k = 0
    for ident in range(n_ids):
        base = rng.integers(0, 255, size=(8, 8, 3), dtype=np.uint8)
        for _ in range(per_id):
            k += 1
            img_id = f"{k:05d}"
            face = np.asarray(Image.fromarray(base).resize((full, full), Image.BICUBIC)).astype(np.int16)
            face = np.clip(face + rng.integers(-20, 20, size=face.shape), 0, 255).astype(np.uint8)
            mask = np.zeros((full, full), dtype=np.uint8)
            mask[full // 2:int(full * 0.9), int(full * 0.2):int(full * 0.8)] = 1
            src = face.copy()
            src[mask > 0] = 255
            Image.fromarray(face).save(root / "ref" / f"{img_id}.jpg", quality=95)
            Image.fromarray(src).save(root / "src" / f"{img_id}_surgical.jpg", quality=95)
            np.save(root / "mask" / f"{img_id}.npy", mask)
            lines.append(f"{img_id}.jpg {ident}")
    (root / "identity.txt").write_text("\n".join(lines) + "\n")
    # `use_ssim=True` (psp_inference.py:153, train_psp.py:143) loads <source_dir>/../best_reference_map.pkl when it exists;
    # the code that would build it is broken as shipped (dataloader.py:200-201 calls `.copy()` on a tensor)
    import pickle
    by_ident = {}
    for ln in lines:
        name, ident = ln.split(" ")
        by_ident.setdefault(ident, []).append(name.split(".")[0])
    best = {n: next(o for o in names if o != n) for names in by_ident.values() for n in names}
    with open(root / "best_reference_map.pkl", "wb") as f:
        pickle.dump(best, f)
    return {"data_root": str(root), "src_img_path": "src", "ref_img_path": "ref", "mask_path": "mask",
            "identity_file_path": "identity.txt", "n": k}
