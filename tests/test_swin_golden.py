"""Pin the Swin CPU oracle against the reference's own SwinTransformer outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import swin_oracle as WO
from vnext_b200.swin_spec import swin_state_spec
from vnext_b200.synth import fill_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    "w12": (dict(embed_dim=96, depths=(2, 2, 2, 2), num_heads=(3, 6, 12, 24), window_size=12), 2, 160, 224, 0),
    "w7": (dict(embed_dim=64, depths=(2, 2, 2, 2), num_heads=(2, 4, 8, 16), window_size=7), 1, 96, 130, 1),
}


@pytest.mark.parametrize("name", list(CASES))
def test_swin_oracle_matches_reference(name):
    golden = np.load(os.path.join(ROOT, "tests", "golden", "swin_golden.npz"))
    cfg, B, H, W, seed = CASES[name]
    sd = fill_state_dict(swin_state_spec("", **cfg), seed)
    x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed + 50))
    with torch.no_grad():
        out = WO.swin(sd, x, "", **cfg)
    for k in ("res2", "res3", "res4", "res5"):
        g = torch.from_numpy(golden[f"{name}_{k}"])
        assert torch.allclose(out[k][:, ::4], g, rtol=1e-3, atol=1e-4), k
