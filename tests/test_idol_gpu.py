"""GPU parity of the IDOL per-frame engine against the CPU oracle and the reference-module goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import idol_oracle as IO
from vnext_b200.idol.state_spec import idol_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"nopad": (3, 96, 128, 0), "pad": (2, 80, 112, 1)}


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("tf32", 2e-2)])
def test_idol_engine(name, precision, tol):
    from vnext_b200.idol.engine import IDOLEngine
    golden = np.load(os.path.join(ROOT, "tests", "golden", "idol_golden.npz"))
    T, H, W, seed = CASES[name]
    sd = fill_state_dict(idol_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    out = IDOLEngine(sd, device="cuda", precision=precision).inference_forward(torch.stack(frames).cuda())
    with torch.no_grad():
        ref = IO.inference_forward(sd, frames)
    errs = {k: _rel(out[k].cpu(), ref[k]) for k in ("pred_logits", "pred_boxes", "pred_inst_embed", "pred_masks")}
    print(precision, name, errs)
    assert max(errs.values()) < tol
    if precision == "fp32":
        g = lambda k: torch.from_numpy(golden[f"{name}_{k}"])
        assert _rel(out["pred_logits"].cpu(), g("pred_logits")) < 1e-3
        assert _rel(out["pred_inst_embed"].cpu()[:, ::5], g("pred_inst_embed")) < 1e-3
        assert _rel(out["pred_masks"].cpu()[:, ::10], g("pred_masks")) < 1e-3


def test_idol_facade_chunks_long_videos():
    from vnext_b200.idol.meta_arch import IDOL
    sd = fill_state_dict(idol_r50_state_spec(), 0)
    frames = synthetic_clip(5, 96, 128, 3)
    m = IDOL(state_dict=sd, device="cuda", precision="fp32").eval()
    m.batch_infer_len = 2                           # 5 frames -> chunks of 2, 2, 1
    out = m.dense_forward([{"image": frames, "height": 96, "width": 128}])
    assert out["pred_logits"].shape == (5, 300, 40) and out["pred_masks"].shape == (5, 300, 1, 24, 32)
    m.batch_infer_len = 10
    whole = m.dense_forward([{"image": frames, "height": 96, "width": 128}])
    assert _rel(out["pred_logits"], whole["pred_logits"]) < 1e-4      # frames are independent
    res = m([{"image": frames, "height": 72, "width": 100}])        # full eval forward: dense path + online tracker
    assert set(res) == {"image_size", "pred_scores", "pred_labels", "pred_masks"} and res["image_size"] == (72, 100)
    assert len(res["pred_scores"]) == len(res["pred_labels"]) == len(res["pred_masks"])
    for per_frame in res["pred_masks"]:
        assert len(per_frame) == 5 and all(x is None or tuple(x.shape) == (72, 100) for x in per_frame)
    # the same video with the masks in the evaluator's wire format (device RLE): every string decodes to the boolean mask
    from oracle import rle_oracle as R
    m.mask_format = "rle"
    rle = m([{"image": frames, "height": 72, "width": 100}])
    m.mask_format = "bool"
    assert rle["pred_scores"] == res["pred_scores"] and rle["pred_labels"] == res["pred_labels"]
    n_masks = 0
    for a, b in zip(rle["pred_masks"], res["pred_masks"]):
        for ra, mb in zip(a, b):
            assert (ra is None) == (mb is None)
            if ra is not None:
                assert ra["size"] == [72, 100] and (R.decode(R.from_string(ra["counts"]), 72, 100) == mb.numpy()).all()
                n_masks += 1
    assert n_masks > 0 or len(res["pred_masks"]) == 0
    seen = []
    m.tracker_fn = lambda out, ori, img: seen.append((ori, img, out["pred_logits"].shape)) or "mine"
    assert m([{"image": frames, "height": 72, "width": 100}]) == "mine" and seen == [((72, 100), (96, 128), (5, 300, 40))]


@pytest.mark.parametrize("name", ["default", "seed1", "nofw"])
def test_tracker_on_device_matches_reference(name):
    from test_idol_tracker_golden import check_case
    check_case(name, "cuda")
