"""SeqFormer CLIP_MATCHING inference (vnext_b200/seqformer/clip_matching.py) against the REFERENCE's own
`inference_clip` + `Videos`/`Clips` + `clip_matching_postprocess` (tests/golden/make_seqformer_clipmatch_golden.py)."""
import os

import numpy as np
import pytest
import torch

from vnext_b200.seqformer.clip_matching import (ClipTracks, clip_matching_postprocess, clip_windows,
                                                select_clip_instances)
from vnext_b200.synth import synthetic_clip_outputs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "seqformer_clipmatch_golden.npz")
CASES = {"stride1": dict(video_length=9, clip_length=5, stride=1, seed=0),
         "stride3": dict(video_length=11, clip_length=5, stride=3, seed=1),
         "short": dict(video_length=4, clip_length=5, stride=1, seed=2)}
ORI, IMG = (40, 60), (44, 62)


def run_case(name, device):
    gold = np.load(GOLD)
    kw = CASES[name]
    windows = synthetic_clip_outputs(**kw)
    assert [f for f, _ in windows] == clip_windows(kw["video_length"], kw["clip_length"], kw["stride"])
    tracks = None
    for frame_idx, out in windows:
        logits, masks = out["pred_logits"].to(device), out["pred_masks"].to(device)
        if tracks is None:
            tracks = ClipTracks(kw["clip_length"], kw["video_length"], 40, tuple(masks.shape[-2:]), device)
        cls, ml = select_clip_instances(logits, masks)
        tracks.update(frame_idx, cls, ml)
    pred_cls, pred_logits = tracks.get_result()
    assert pred_cls.shape[0] == int(gold[f"{name}_num_inst"])
    rtol = 1e-5 if device == "cpu" else 5e-4                     # CUDA sigmoid / reduction order differ in the last bits
    np.testing.assert_allclose(pred_cls.cpu().numpy(), gold[f"{name}_cls"], rtol=rtol, atol=1e-7)
    res = clip_matching_postprocess(pred_cls, pred_logits, ORI, IMG)
    assert res["pred_labels"] == gold[f"{name}_labels"].tolist()
    np.testing.assert_allclose(np.array(res["pred_scores"], dtype=np.float32), gold[f"{name}_scores"], rtol=rtol)
    n = len(res["pred_scores"])
    got = np.packbits(torch.stack(res["pred_masks"]).numpy().reshape(n, -1), axis=1)
    mismatch = int(np.unpackbits(got ^ gold[f"{name}_masks"]).sum())
    assert mismatch <= (0 if device == "cpu" else 16), mismatch
    return tracks


@pytest.mark.parametrize("name", list(CASES))
def test_clip_matching_matches_reference(name):
    tracks = run_case(name, "cpu")
    # the state is the running sums + a short window, not the reference's [clips, 120, L, h, w] tensors
    assert tracks.logit_sum.shape[0] == tracks.num_inst and len(tracks.window) <= CASES[name]["clip_length"]


def test_clip_windows():
    assert clip_windows(4, 5, 1) == [[0, 1, 2, 3]]
    assert clip_windows(7, 5, 1) == [[0, 1, 2, 3, 4], [1, 2, 3, 4, 5], [2, 3, 4, 5, 6]]
    assert clip_windows(11, 5, 3) == [[0, 1, 2, 3, 4], [3, 4, 5, 6, 7], [6, 7, 8, 9, 10]]
