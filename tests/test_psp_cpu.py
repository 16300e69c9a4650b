"""modules/psp.py is a drop-in for the reference's pSp + GradualStyleEncoder: identical state_dict layout (keys pinned in
tests/golden/refpsp256.npz, recorded from the reference itself). No GPU needed."""
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / "golden" / "refpsp256.npz"


def test_state_dict_layout_matches_reference():
    from face_mask_inpaint_b200.modules.psp import pSp, refpsp_opts
    g = np.load(GOLD)
    net = pSp(refpsp_opts(output_size=256))
    mine = sorted(net.state_dict().keys())
    want = [str(k) for k in g["keys"]]
    assert mine == want, f"missing {sorted(set(want) - set(mine))[:5]} extra {sorted(set(mine) - set(want))[:5]}"
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"])
    assert net.latent_avg.shape == (14, 512) and not any(p.requires_grad for p in net.decoder.parameters())
