"""Pin the CPU model oracle (oracle/seqformer_oracle.py) against outputs of the reference's own
SeqFormer modules (tests/golden/make_seqformer_golden.py), and the state-dict spec against the
reference module tree."""
import os

import numpy as np
import pytest
import torch

from oracle import seqformer_oracle as SO
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"nopad": (3, 96, 128, 0), "pad": (2, 80, 112, 1)}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "seqformer_golden.npz"))


def test_state_spec_matches_reference_tree(golden):
    spec = seqformer_r50_state_spec()
    assert len(spec) == int(golden["num_state_keys"][0])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_modules(golden, name):
    T, H, W, seed = CASES[name]
    sd = fill_state_dict(seqformer_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    with torch.no_grad():
        out = SO.inference(sd, frames, return_intermediates=True)
    g = lambda k: torch.from_numpy(golden[f"{name}_{k}"])
    assert torch.allclose(out["feats"]["res3"].flatten()[::613], g("res3_sample"), rtol=1e-4, atol=1e-4)
    assert torch.allclose(out["feats"]["res5"][:, ::16], g("res5"), rtol=1e-4, atol=1e-3)
    assert torch.allclose(out["pred_logits"], g("pred_logits"), rtol=1e-3, atol=1e-3)
    assert torch.allclose(out["pred_boxes"], g("pred_boxes"), rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["pred_masks"][:, ::10], g("pred_masks"), rtol=1e-3, atol=2e-3)
