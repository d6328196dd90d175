"""The RLE oracle (oracle/rle_oracle.py, restating pycocotools maskApi.c) against the RLE strings the reference's own tests
hold (tests/golden/rle_golden.json <- tests/data/test_coco_evaluation.py:24, tests/test_visualizer.py:54), and the host-side
pieces of vnext_b200/rle.py / seqformer/ytvis_eval.py that need no GPU."""
import json
import os

import numpy as np
import pytest

from oracle import rle_oracle as R

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rle_golden.json")))


@pytest.mark.parametrize("case", GOLDEN, ids=[f"{c['source']}#{i}" for i, c in enumerate(GOLDEN)])
def test_golden_strings_decode_and_reencode_identically(case):
    h, w = case["size"]
    cnts = R.from_string(case["counts"])
    assert sum(cnts) == h * w and all(c >= 0 for c in cnts)
    mask = R.decode(cnts, h, w)
    assert R.encode(mask)["counts"].decode("ascii") == case["counts"]
    assert R.to_string_scalar(R.counts_scalar(mask)).decode("ascii") == case["counts"]
    assert list(R.counts(mask)) == cnts == R.counts_scalar(mask)


def test_documented_examples_of_the_format():
    # maskApi.h: M = [0 0 1 1 1 0 1] -> counts [2 3 1 1];  M = [1 1 1 1 1 1 0] -> counts [0 6 1] (odd counts are zeros runs)
    assert R.counts_scalar(np.array([[0, 0, 1, 1, 1, 0, 1]]).T) == [2, 3, 1, 1]
    assert R.counts_scalar(np.array([[1, 1, 1, 1, 1, 1, 0]]).T) == [0, 6, 1]
    assert list(R.counts(np.array([[1, 1, 1, 1, 1, 1, 0]]).T)) == [0, 6, 1]


@pytest.mark.parametrize("seed", range(6))
def test_vectorised_equals_scalar_and_round_trips(seed):
    g = np.random.default_rng(seed)
    h, w = int(g.integers(1, 40)), int(g.integers(1, 40))
    p = [0.02, 0.5, 0.98, 0.3, 0.0, 1.0][seed]
    mask = (g.random((h, w)) < p).astype(np.uint8)
    cnts = R.counts_scalar(mask)
    assert list(R.counts(mask)) == cnts
    s = R.to_string_scalar(cnts)
    assert R.to_string(cnts) == s
    assert R.from_string(s) == cnts
    assert (R.decode(cnts, h, w) == mask).all()


def test_long_runs_need_several_characters_and_negative_deltas():
    cnts = [5_000_000, 1, 3, 900_000, 2, 1]                # deltas: +, +, +, 900000-1, 2-3 = -1, 1-900000
    s = R.to_string_scalar(cnts)
    assert R.to_string(cnts) == s and R.from_string(s) == cnts


def test_split_strings_and_empty_rle_helpers():
    import torch
    from vnext_b200.rle import split_strings
    from vnext_b200.seqformer.ytvis_eval import _empty_rle
    a, b = R.encode(np.eye(7, 9, dtype=np.uint8))["counts"], R.encode(np.zeros((7, 9), np.uint8))["counts"]
    chars = torch.frombuffer(bytearray(a + b + b"junk"), dtype=torch.uint8)
    out = split_strings(torch.tensor([0, len(a), len(a) + len(b)]), chars, 7, 9)
    assert out == [{"size": [7, 9], "counts": a.decode()}, {"size": [7, 9], "counts": b.decode()}]
    for h, w in [(7, 9), (720, 1280), (1, 1), (4000, 6000)]:
        assert _empty_rle(h, w) == {"size": [h, w], "counts": R.to_string_scalar([h * w]).decode()}


def test_evaluator_record_from_already_encoded_masks_needs_no_device():
    """instances_to_coco_json_video on mask_format="rle" outputs: pure assembly (ytvis_eval.py:196-211); a frame where the
    instance is absent (None) gets the all-zero mask's RLE like the reference."""
    from vnext_b200.seqformer.ytvis_eval import instances_to_coco_json_video
    h, w = 6, 8
    m = np.zeros((h, w), np.uint8)
    m[2:4, 3:6] = 1
    rle = {"size": [h, w], "counts": R.encode(m)["counts"].decode()}
    inputs = [{"video_id": 3, "length": 2, "height": h, "width": w}]
    outputs = {"pred_scores": [0.9, 0.4], "pred_labels": [5, 1], "pred_masks": [[rle, None], [None, rle]]}
    rec = instances_to_coco_json_video(inputs, outputs)
    zero = {"size": [h, w], "counts": R.encode(np.zeros((h, w), np.uint8))["counts"].decode()}
    assert rec == [{"video_id": 3, "score": 0.9, "category_id": 5, "segmentations": [rle, zero]},
                   {"video_id": 3, "score": 0.4, "category_id": 1, "segmentations": [zero, rle]}]
    assert (R.decode(R.from_string(rec[0]["segmentations"][0]["counts"]), h, w) == m).all()
    with pytest.raises(AssertionError):
        instances_to_coco_json_video(inputs * 2, outputs)
