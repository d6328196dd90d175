"""Host-side training logic (vnext_b200/idol/criterion.py) against what the REFERENCE's own matcher / criterion /
select_pos_neg produced (tests/golden/make_idol_train_golden.py): assignments bit-exact, loss terms to 1e-5."""
import os
import random

import numpy as np
import pytest
import torch

from vnext_b200.idol import criterion as crit
from vnext_b200.synth import synthetic_targets

GOLD = os.path.join(os.path.dirname(__file__), "golden", "idol_train_golden.npz")
B, H, W, SEED = 2, 96, 128, 3


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


@pytest.fixture(scope="module")
def targets():
    return synthetic_targets(B, H, W, SEED)


def _indices(gold, lvl):
    return [(torch.from_numpy(gold[f"lvl{lvl}_b{b}_selected"]), torch.from_numpy(gold[f"lvl{lvl}_b{b}_gt"]))
            for b in range(B)]


def test_simota_assignment_bit_exact(gold, targets):
    det, _ = targets
    for lvl in range(6):
        logits = torch.from_numpy(gold[f"lvl{lvl}_pred_logits"])
        boxes = torch.from_numpy(gold[f"lvl{lvl}_pred_boxes"])
        indices, matched = crit.simota_match(logits, boxes, det)
        for b in range(B):
            assert torch.equal(indices[b][0], torch.from_numpy(gold[f"lvl{lvl}_b{b}_selected"]))
            assert torch.equal(indices[b][1], torch.from_numpy(gold[f"lvl{lvl}_b{b}_gt"]))
            assert torch.equal(matched[b], torch.from_numpy(gold[f"lvl{lvl}_b{b}_matched"]))
        assert sum(int(i[0].sum()) for i in indices) >= 3 * B


def test_simota_empty_targets():
    logits, boxes = torch.randn(1, 20, 4), torch.rand(1, 20, 4)
    t = [{"labels": torch.zeros(0, dtype=torch.long), "boxes": torch.zeros(0, 4)}]
    indices, matched = crit.simota_match(logits, boxes, t)
    assert indices[0][0].sum() == 0 and len(indices[0][1]) == 0 and len(matched[0]) == 0


def test_label_and_box_losses(gold, targets):
    det, ref = targets
    c = crit.SetCriterion(40, crit.default_weight_dict())
    for lvl in range(6):
        out = {"pred_logits": torch.from_numpy(gold[f"lvl{lvl}_pred_logits"]),
               "pred_boxes": torch.from_numpy(gold[f"lvl{lvl}_pred_boxes"])}
        sfx = "" if lvl == 5 else f"_{lvl}"
        got = {**c.loss_labels(out, det, _indices(gold, lvl)), **c.loss_boxes(out, det, _indices(gold, lvl))}
        for k, v in got.items():
            np.testing.assert_allclose(float(v), float(gold["loss/" + k + sfx]), rtol=1e-5, atol=1e-6)


def test_mask_losses(gold, targets):
    det, ref = targets
    c = crit.SetCriterion(40, crit.default_weight_dict())
    out = {"pred_masks": [torch.from_numpy(gold[f"final_pred_masks_b{b}"]) for b in range(B)]}
    got = c.loss_masks(out, det, ref, _indices(gold, 5))
    np.testing.assert_allclose(float(got["loss_mask"]), float(gold["loss/loss_mask"]), rtol=1e-5)
    np.testing.assert_allclose(float(got["loss_dice"]), float(gold["loss/loss_dice"]), rtol=1e-5)


def test_select_pos_neg_and_reid_loss(gold, targets):
    det, ref = targets
    matched = [torch.from_numpy(gold[f"lvl5_b{b}_matched"]) for b in range(B)]
    random.seed(0)
    items = crit.select_pos_neg(torch.from_numpy(gold["reid_ref_box"]), matched, ref, det,
                                torch.from_numpy(gold["reid_key_embed"]), torch.from_numpy(gold["reid_ref_embed"]),
                                torch.from_numpy(gold["reid_ref_cls"]))
    assert len(items) == int(gold["num_contrast_items"]) == 5        # one instance left the second video's ref frame
    for i, it in enumerate(items):
        assert np.array_equal(it["label"].numpy(), gold[f"qd{i}_label"])
        assert np.array_equal(it["aux_label"].numpy(), gold[f"qd{i}_aux_label"])
        np.testing.assert_allclose(it["contrast"].numpy(), gold[f"qd{i}_contrast"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(it["aux_consin"].numpy(), gold[f"qd{i}_aux_cos"], rtol=1e-5, atol=1e-6)
    c = crit.SetCriterion(40, crit.default_weight_dict())
    got = c.loss_reid({"pred_qd": items, "pred_logits": torch.zeros(1)})
    np.testing.assert_allclose(float(got["loss_reid"]), float(gold["loss/loss_reid"]), rtol=1e-5)
    np.testing.assert_allclose(float(got["loss_reid_aux"]), float(gold["loss/loss_reid_aux"]), rtol=1e-5)


def test_paired_giou_matches_pairwise_diagonal():
    g = torch.Generator().manual_seed(0)
    a = crit.cxcywh_to_xyxy(torch.rand(16, 4, generator=g) * 0.5 + 0.2)
    b = crit.cxcywh_to_xyxy(torch.rand(16, 4, generator=g) * 0.5 + 0.2)
    np.testing.assert_allclose((1 - crit.pairwise_giou(a, b).diag()).numpy(), crit.paired_giou_loss(a, b).numpy(),
                               atol=1e-5)
