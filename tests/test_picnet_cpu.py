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
