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
