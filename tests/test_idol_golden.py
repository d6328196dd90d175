"""Pin the IDOL CPU oracle against outputs of the reference's own IDOL modules; state-dict spec vs the reference tree."""
import os

import numpy as np
import pytest
import torch

from oracle import idol_oracle as IO
from vnext_b200.idol.state_spec import idol_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"nopad": (3, 96, 128, 0), "pad": (2, 80, 112, 1)}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "idol_golden.npz"))


def test_state_spec_matches_reference_tree(golden):
    assert len(idol_r50_state_spec()) == int(golden["num_state_keys"][0])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_modules(golden, name):
    T, H, W, seed = CASES[name]
    sd = fill_state_dict(idol_r50_state_spec(), seed)
    with torch.no_grad():
        out = IO.inference_forward(sd, synthetic_clip(T, H, W, seed))
    g = lambda k: torch.from_numpy(golden[f"{name}_{k}"])
    assert torch.allclose(out["pred_logits"], g("pred_logits"), rtol=1e-3, atol=1e-3)
    assert torch.allclose(out["pred_boxes"], g("pred_boxes"), rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["pred_inst_embed"][:, ::5], g("pred_inst_embed"), rtol=1e-3, atol=1e-3)
    assert torch.allclose(out["pred_masks"][:, ::10], g("pred_masks"), rtol=1e-3, atol=2e-3)
