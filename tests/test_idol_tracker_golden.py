"""Online association + video post-processing (vnext_b200/idol/tracker.py) against the REFERENCE's own
`IDOL.inference` + `IDOL_Tracker` outputs on scripted videos (tests/golden/make_idol_tracker_golden.py)."""
import os

import numpy as np
import pytest
import torch

from vnext_b200.idol.tracker import IDOLTracker, greedy_mask_nms, pairwise_mask_iou, video_inference
from vnext_b200.synth import synthetic_track_outputs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "idol_tracker_golden.npz")
CASES = {"default": dict(seed=0), "seed1": dict(seed=1), "nofw": dict(seed=2, fw=False, tw=False)}
ORI, IMG = (72, 100), (90, 120)


def make_tracker(fw=True, tw=True):
    return IDOLTracker(init_score_thr=0.2, obj_score_thr=0.1, nms_thr_pre=0.5, nms_thr_post=0.05, addnew_score_thr=0.2,
                       memo_tracklet_frames=10, memo_momentum=0.8, long_match=tw, frame_weight=(tw | fw),
                       temporal_weight=tw, memory_len=3)                                  # idol.py:276-290


def check_case(name, device):
    gold = np.load(GOLD)
    kw = dict(CASES[name])
    out = {k: v.to(device) for k, v in synthetic_track_outputs(seed=kw.pop("seed")).items()}
    res = video_inference(out, make_tracker(**kw), ORI, IMG)
    assert res["image_size"] == ORI
    assert res["pred_labels"] == gold[f"{name}_labels"].tolist()
    np.testing.assert_allclose(np.array(res["pred_scores"], dtype=np.float32), gold[f"{name}_scores"], rtol=1e-5, atol=1e-6)
    n, T = len(res["pred_scores"]), 14
    present = np.array([[m is not None for m in per_frame] for per_frame in res["pred_masks"]])
    assert np.array_equal(present, gold[f"{name}_present"])
    masks = np.zeros((n, T) + ORI, dtype=bool)
    for i, per_frame in enumerate(res["pred_masks"]):
        for t, m in enumerate(per_frame):
            if m is not None:
                assert m.dtype == torch.bool and tuple(m.shape) == ORI and m.device.type == "cpu"
                masks[i, t] = m.numpy()
    got = np.packbits(masks.reshape(n, -1), axis=1)
    mismatch = np.unpackbits(got ^ gold[f"{name}_masks"]).sum()
    assert mismatch <= (0 if device == "cpu" else 8), mismatch          # GPU bilinear may flip a borderline pixel


@pytest.mark.parametrize("name", list(CASES))
def test_video_inference_matches_reference(name):
    check_case(name, "cpu")


def test_pairwise_mask_iou_and_nms():
    g = torch.Generator().manual_seed(0)
    m = torch.rand(6, 50, generator=g) > 0.5
    m[3] = m[0]                                                   # a duplicate: IoU 1 -> suppressed
    iou = pairwise_mask_iou(m)
    i, j = m[1].float(), m[4].float()
    want = ((i * j).sum() + 1e-6) / ((i + j - i * j).sum() + 1e-6)
    assert torch.equal(iou[1, 4], want) and torch.allclose(iou.diag(), torch.ones(6))
    keep = greedy_mask_nms(iou, 0.9)
    assert keep == [True, True, True, False, True, True]
    assert greedy_mask_nms(iou[:0, :0], 0.5) == []


def test_empty_video_and_no_confident_frame():
    out = synthetic_track_outputs(num_frames=3, seed=5)
    out["pred_logits"].fill_(-8.0)                               # nothing above any threshold: top-1 fallback, no tracks
    res = video_inference(out, make_tracker(), ORI, IMG)
    assert res["pred_scores"] == [] and res["pred_labels"] == [] and res["pred_masks"] == []
