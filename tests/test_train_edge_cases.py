"""Training-path edge cases the reference handles explicitly: a batch with NO ground truth at all (idol/models/matcher.py:
59-65, deformable_detr.py `len(gt_multi_idx)==0` branches, pos_neg_select.py valid filtering) and a video whose only
object is absent from the reference frame.  CPU, oracle MSDA (host logic under test, not the kernels)."""
import random

import pytest
import torch

from oracle.msda_oracle import forward_torch as msda_core
from vnext_b200.idol import train as IT
from vnext_b200.idol.state_spec import idol_r50_state_spec
from vnext_b200.seqformer import train as ST
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_clip_targets, synthetic_targets

H, W = 64, 96


class OracleMSDA:
    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return msda_core(value, shapes, loc, attn)


def _empty(t):
    return {k: (v[:0] if torch.is_tensor(v) and v.dim() > 0 and k != "size" else v) for k, v in t.items()}


@pytest.mark.parametrize("batched", [False, True])
def test_idol_no_ground_truth(monkeypatch, batched):
    monkeypatch.setattr(IT, "MSDeformAttnFunction", OracleMSDA)
    model = IT.IDOLTrainModel(fill_state_dict(idol_r50_state_spec(), 0), device="cpu", dropout=0.0, batched=batched).train()
    det, ref = synthetic_targets(2, H, W, 0)
    det, ref = [_empty(t) for t in det], [_empty(t) for t in ref]
    random.seed(0)
    loss, _ = model(torch.stack(synthetic_clip(4, H, W, 0)), det, ref)
    assert len(loss) == 32 and all(torch.isfinite(v) for v in loss.values())
    for k, v in loss.items():                          # nothing matched: every term but the classification one is zero
        assert float(v) == 0 or k.startswith("loss_ce"), k
    sum(model.weighted(loss).values()).backward()
    assert torch.isfinite(model.flat.g).all()


@pytest.mark.parametrize("batched", [False, True])
def test_idol_object_missing_from_reference_frame(monkeypatch, batched):
    monkeypatch.setattr(IT, "MSDeformAttnFunction", OracleMSDA)
    model = IT.IDOLTrainModel(fill_state_dict(idol_r50_state_spec(), 0), device="cpu", dropout=0.0, batched=batched).train()
    det, ref = synthetic_targets(1, H, W, 0, insts=1)
    ref[0]["valid"][:] = False                          # the only instance left the reference frame: no contrastive item
    ref[0]["inst_id"][:] = -1
    random.seed(0)
    loss, out = model(torch.stack(synthetic_clip(2, H, W, 0)), det, ref)
    assert float(loss["loss_reid"]) == 0 and float(loss["loss_reid_aux"]) == 0 and float(loss["loss_mask"]) > 0
    sum(model.weighted(loss).values()).backward()
    assert torch.isfinite(model.flat.g).all()


@pytest.mark.parametrize("batched", [False, True])
def test_seqformer_no_ground_truth(monkeypatch, batched):
    monkeypatch.setattr(IT, "MSDeformAttnFunction", OracleMSDA)
    model = ST.SeqFormerTrainModel(fill_state_dict(seqformer_r50_state_spec(), 0), 2, device="cpu", dropout=0.0,
                                   batched=batched).train()
    targets = [_empty(t) for t in synthetic_clip_targets(2, 2, H, W, 0)]
    loss, _ = model(torch.stack(synthetic_clip(4, H, W, 0)), targets)
    assert all(torch.isfinite(v) for v in loss.values())
    assert float(loss["loss_bbox"]) == 0 and float(loss["loss_mask"]) == 0 and float(loss["loss_ce"]) > 0
    sum(model.weighted(loss).values()).backward()
    assert torch.isfinite(model.flat.g).all()
