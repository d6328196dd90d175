"""The IDOL training forward/backward (vnext_b200/idol/train.py) on CPU against the reference's own losses and
gradients (tests/golden/make_idol_train_golden.py).  The product path has no CPU implementation of the sampling op, so
THIS TEST substitutes the oracle's (oracle.msda_oracle.forward_torch = the reference's ms_deform_attn_core_pytorch
restated) for the autograd Function: what is pinned here is everything around the kernels — module wiring, key/ref
batching, box-refinement detach points, the once-per-step mask branch, flat parameter/gradient views.  The same
comparison with the real sm_100a forward/backward kernels is tests/test_idol_train_gpu.py."""
import os
import random

import numpy as np
import pytest
import torch

from oracle.msda_oracle import forward_torch as msda_core
from vnext_b200.idol import train as T
from vnext_b200.idol.state_spec import idol_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_targets

GOLD = os.path.join(os.path.dirname(__file__), "golden", "idol_train_golden.npz")
B, H, W, SEED = 2, 96, 128, 3
ALIAS_FROM, ALIAS_TO = "detr.detr.transformer.decoder.bbox_embed.", "detr.detr.bbox_embed."


class OracleMSDA:
    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return msda_core(value, shapes, loc, attn)


def compare_with_golden(model, gold, loss_dict, rtol_loss, rtol_norm, tol_grad):
    for k, v in loss_dict.items():
        np.testing.assert_allclose(float(v), float(gold["loss/" + k]), rtol=rtol_loss, atol=1e-5, err_msg=k)
    keys = [k[len("gradnorm/"):] for k in gold if k.startswith("gradnorm/")]
    assert len(keys) == 360
    worst = 0.0
    for k in keys:
        mine = k.replace(ALIAS_FROM, ALIAS_TO)
        g = model.flat.params[mine].grad
        want = float(gold["gradnorm/" + k])
        rel = abs(float(g.norm()) - want) / max(want, 1e-6)
        worst = max(worst, rel)
        assert rel < rtol_norm, (k, float(g.norm()), want)
    for k in gold:
        if k.startswith("grad/"):
            g = model.flat.params[k[5:].replace(ALIAS_FROM, ALIAS_TO)].grad.cpu().numpy()
            assert np.abs(g - gold[k]).max() <= tol_grad * max(np.abs(gold[k]).max(), 1e-6), k
        if k.startswith("gradstrided/"):
            g = model.flat.params[k[12:]].grad.flatten()[::97].cpu().numpy()
            assert np.abs(g - gold[k]).max() <= tol_grad * max(np.abs(gold[k]).max(), 1e-6), k
    return worst


def check_assignment(gold, outputs, batched):
    """simOTA picked exactly the reference's queries / ground truths at every decoder level."""
    for lvl in range(6):
        for b in range(B):
            want_sel, want_gt = gold[f"lvl{lvl}_b{b}_selected"], gold[f"lvl{lvl}_b{b}_gt"]
            if batched:
                a = outputs["assign"]
                sel = a["sel"][lvl, b].cpu().numpy()
                assert np.array_equal(sel, want_sel)
                assert np.array_equal(a["gt_of"][lvl, b].cpu().numpy()[sel], want_gt)
                assert np.array_equal(a["best_q"][lvl, b].cpu().numpy()[:len(gold[f"lvl{lvl}_b{b}_matched"])],
                                      gold[f"lvl{lvl}_b{b}_matched"])
            elif lvl == 5:
                assert outputs["pred_masks"][b].shape[1] == int(want_sel.sum())


@pytest.mark.timeout(900)
@pytest.mark.parametrize("batched", [False, True])
def test_train_forward_backward_matches_reference(monkeypatch, batched):
    gold = dict(np.load(GOLD))
    monkeypatch.setattr(T, "MSDeformAttnFunction", OracleMSDA)
    sd = fill_state_dict(idol_r50_state_spec(), SEED)
    model = T.IDOLTrainModel(sd, device="cpu", dropout=0.0, batched=batched).train()
    assert len(model.flat.params) == 360 and all(p.grad.data_ptr() >= model.flat.g.data_ptr() for p in model.parameters())
    frames = torch.stack(synthetic_clip(2 * B, H, W, SEED))
    det, ref = synthetic_targets(B, H, W, SEED)
    random.seed(0)
    loss_dict, outputs = model(frames, det, ref)
    assert set(loss_dict) == {k[5:] for k in gold if k.startswith("loss/")}
    check_assignment(gold, outputs, batched)
    total = sum(model.weighted(loss_dict).values())
    np.testing.assert_allclose(float(total.detach()), float(gold["total"]), rtol=1e-4)
    total.backward()
    worst = compare_with_golden(model, gold, loss_dict, rtol_loss=1e-4, rtol_norm=2e-3, tol_grad=2e-3)
    print("worst grad-norm rel err", worst)
    # gradients landed in the flat buffer (views), frozen tensors got none
    assert float(model.flat.g.abs().sum()) > 0
    assert all(not v.requires_grad for v in model.frozen.values())


def test_batched_matcher_repair_loop_falls_back_to_sequential():
    """A ground truth no query wants (the reference's `while` loop, matcher.py:151-162) goes through the sequential code:
    two identical GT boxes compete for the same single best query."""
    from vnext_b200.idol import criterion as seq
    from vnext_b200.idol import criterion_batched as cb
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(1, 40, 5, generator=g) - 3
    boxes = torch.rand(1, 40, 4, generator=g) * 0.2 + 0.4
    boxes[0, :, 2:] = 0.02                                                   # tiny boxes: IoU ~ 0 -> dynamic k = 1
    tgt = [{"labels": torch.tensor([1, 1]), "boxes": torch.tensor([[0.5, 0.5, 0.3, 0.3], [0.5, 0.5, 0.3, 0.3]])}]
    ind, mq = seq.simota_match(logits.clone(), boxes.clone(), tgt)
    gtb, gtl, ex = cb.pad_targets(tgt, "cpu")
    sel, gt_of, best_q = cb.simota_batched(logits, boxes, gtb, gtl, ex)
    assert torch.equal(sel[0], ind[0][0]) and torch.equal(gt_of[0][sel[0]], ind[0][1]) and torch.equal(best_q[0], mq[0])
    assert int(sel.sum()) == 2


def test_reference_state_dict_roundtrip():
    spec = idol_r50_state_spec()
    sd = fill_state_dict(spec, 1)
    model = T.IDOLTrainModel(sd, device="cpu")
    out = model.reference_state_dict()
    assert set(out) == set(spec)
    for k in spec:
        src = sd[k.replace(ALIAS_FROM, ALIAS_TO)] if k.startswith(ALIAS_FROM) else sd[k]
        assert torch.equal(out[k], src), k


def test_adamw_refuses_cpu():
    z = torch.zeros(4)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        T.adamw_(z, z, z, z, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, torch.ones(1))


class _Boxes:
    def __init__(self, t):
        self.tensor = t


class FakeInstances:
    """Duck type of detectron2.structures.Instances as the YTVIS dataset mapper fills it."""

    def __init__(self, t, h, w):
        from vnext_b200.idol.criterion import cxcywh_to_xyxy
        self.image_size = (h, w)
        self.gt_classes = t["labels"]
        self.gt_boxes = _Boxes(cxcywh_to_xyxy(t["boxes"]) * torch.tensor([w, h, w, h], dtype=torch.float))
        self.gt_masks = _Boxes(t["masks"])
        self.gt_ids = t["inst_id"]


def fake_batched_inputs(seed=SEED):
    frames = synthetic_clip(2 * B, H, W, seed)
    det, ref = synthetic_targets(B, H, W, seed)
    return [{"image": [frames[2 * v], frames[2 * v + 1]],
             "instances": [FakeInstances(det[v], H, W), FakeInstances(ref[v], H, W)]} for v in range(B)], det, ref


def test_prepare_targets_matches_reference_format():
    from vnext_b200.idol.meta_arch import IDOL
    batched, det, ref = fake_batched_inputs()
    m = IDOL(device="cpu")
    d2, r2 = m.prepare_targets([i for v in batched for i in v["instances"]])
    for a, b in zip(det + ref, d2 + r2):
        assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["valid"], b["valid"])
        assert torch.allclose(a["boxes"], b["boxes"], atol=1e-6) and torch.equal(a["masks"], b["masks"])
    # an object missing from the KEY frame is dropped from both frames (idol.py:313-321)
    insts = [i for v in batched for i in v["instances"]]
    insts[0].gt_ids = insts[0].gt_ids.clone()
    insts[0].gt_ids[1] = -1
    d3, r3 = m.prepare_targets(insts)
    assert len(d3[0]["labels"]) == len(r3[0]["labels"]) == 2 and len(d3[1]["labels"]) == 3


def test_trainer_checkpoint_roundtrip():
    """state_dict(): weights in the reference key layout + flat Adam moments + step; load_state_dict() restores them
    into a freshly built model (resume)."""
    spec = idol_r50_state_spec()
    a = T.IDOLTrainer(T.IDOLTrainModel(fill_state_dict(spec, 1), device="cpu"))
    with torch.no_grad():
        a.flat.w.add_(0.01)                       # "trained" weights and non-trivial optimiser state
        a.flat.m.uniform_(-1, 1)
        a.flat.v.uniform_(0, 1)
    a.step_count = 17
    state = a.state_dict()
    assert set(state["model"]) == set(spec) and state["step"] == 17
    b = T.IDOLTrainer(T.IDOLTrainModel(fill_state_dict(spec, 2), device="cpu"))
    b.load_state_dict(state)
    assert all(torch.equal(a.flat.params[k], b.flat.params[k]) for k in a.flat.params)    # (alignment gaps are not state)
    assert torch.equal(a.flat.m, b.flat.m) and torch.equal(a.flat.v, b.flat.v)
    assert b.step_count == 17 and all(torch.equal(a.model.frozen[k], b.model.frozen[k]) for k in a.model.frozen)
    state["layout"] = {}
    with pytest.raises(ValueError, match="layout"):
        b.load_state_dict(state)
