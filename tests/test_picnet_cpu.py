"""modules/picnet.py is a drop-in for the reference's ReferenceFill + PICNet networks: identical state_dict layout
(keys pinned in tests/golden/picnet_ref.npz, recorded from the reference itself) and parameter count. No GPU needed."""
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).resolve().parent / "golden" / "picnet_ref.npz"


def test_state_dict_layout_matches_reference():
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    g = np.load(GOLD)
    model = build_picnet_ref()
    mine = sorted(model.state_dict().keys())
    want = [str(k) for k in g["keys"]]
    assert mine == want, f"missing {sorted(set(want) - set(mine))[:5]} extra {sorted(set(mine) - set(want))[:5]}"
    assert sum(p.numel() for p in model.parameters()) == int(g["n_params"]) == 12145892


def test_fill_by_name_is_deterministic():
    from golden_util import fill_by_name
    from face_mask_inpaint_b200.modules.picnet import build_picnet_ref
    a, b = fill_by_name(build_picnet_ref()), fill_by_name(build_picnet_ref())
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    assert float(a.decoder.attn1.gamma) == 1.0


def test_runtime_caches_do_not_travel_with_the_module():
    """Weight plans, side streams and captured graphs live in a per-module cache that deepcopy / pickle replace by an empty one
    (a model that has run on the GPU must stay deep-copyable: EMA copies, torch.save(model))."""
    import copy
    import pickle
    import torch
    from face_mask_inpaint_b200.graphs import NoCopyCache, module_cache
    m = torch.nn.Linear(2, 2)
    module_cache(m)["side_stream"] = object()          # stands for a torch.cuda.Stream (not picklable)
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert isinstance(clone.__dict__["_fmi_cache"], NoCopyCache) and len(clone.__dict__["_fmi_cache"]) == 0
    assert "side_stream" in module_cache(m)


def test_weight_plan_descriptor_table_layout():
    """_WeightPlan.finalize packs one 96-byte FmiSnPrepDesc per convolution ("6Q12i", include/fmi_b200.h): pointers to the
    module's own w_bar / u / v, scratch carved as [4*Wd | Hh] floats per conv, the shared wp buffer of concatenated parts with
    increasing i_off. Built on CPU tensors here (no launch): only the host-side packing is under test."""
    import struct
    import torch
    from torch import nn
    from face_mask_inpaint_b200 import _lib
    from face_mask_inpaint_b200.modules import picnet_fast as PF
    from face_mask_inpaint_b200.modules.picnet_blocks import SpectralNorm
    conv2 = SpectralNorm(nn.ConvTranspose2d(32, 64, 3, 2, 1, 1))      # main path  [I=32, O=64, 3, 3]
    byp = SpectralNorm(nn.ConvTranspose2d(96, 64, 3, 2, 1, 1))        # shortcut   [I=96, O=64, 3, 3]
    conv1 = SpectralNorm(nn.Conv2d(96, 32, 3, 1, 1))
    plan = PF._WeightPlan()
    plan.record([(conv1, False)], 32, False, 96, 3, (9, 32, 96), torch.float32)
    plan.record([(conv2, True), (byp, True)], 64, False, 128, 3, (9, 64, 128), torch.float32)
    plan.finalize(torch.device("cpu"), _lib.MMA_TF32)
    assert plan.ready and plan.n == 3 and len(plan.wps) == 2
    raw = bytes(plan.table.numpy().tobytes())
    assert len(raw) == 3 * 96
    rows = [struct.unpack("6Q12i", raw[k * 96:(k + 1) * 96]) for k in range(3)]
    mods = [conv1.module, conv2.module, byp.module]
    for r, m in zip(rows, mods):
        assert r[0] == m.weight_bar.data_ptr() and r[1] == m.weight_u.data_ptr() and r[2] == m.weight_v.data_ptr()
        hh, wd = r[14], r[15]
        assert hh == m.weight_bar.shape[0] and wd == m.weight_bar.numel() // hh
        assert r[4] - r[3] == 16 * wd                               # u_raw follows the four partial W^T u vectors
    # (O, I, transposed, O_rows, I_row, i_off, merged, T)
    assert rows[0][6:14] == (32, 96, 0, 32, 96, 0, 0, 9)
    assert rows[1][6:14] == (64, 32, 1, 64, 128, 0, 0, 9)
    assert rows[2][6:14] == (64, 96, 1, 64, 128, 32, 0, 9)
    assert rows[1][5] == rows[2][5] == plan.wps[1].data_ptr() and rows[0][5] == plan.wps[0].data_ptr()
    assert plan.max_wd == 96 * 9 and plan.max_hh == 96 and plan.max_elems == 64 * 96 * 9 and plan.valid()
    conv1.module.weight_u.data = conv1.module.weight_u.data.clone()   # a re-bound parameter invalidates the plan
    assert not plan.valid()
