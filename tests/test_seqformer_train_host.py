"""SeqFormer training forward/backward (vnext_b200/seqformer/train.py + criterion.py) on CPU against the reference's
own losses, Hungarian assignments and gradients (tests/golden/make_seqformer_train_golden.py); the sampling op is the
oracle's here (see tests/test_idol_train_host.py for why), the real kernels are exercised by the GPU twin."""
import os

import numpy as np
import pytest
import torch

from oracle.msda_oracle import forward_torch as msda_core
from vnext_b200.idol import train as IT
from vnext_b200.seqformer import criterion as crit
from vnext_b200.seqformer import train as ST
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_clip_targets

GOLD = os.path.join(os.path.dirname(__file__), "golden", "seqformer_train_golden.npz")
B, NF, H, W, SEED = 2, 2, 96, 128, 4
ALIAS_FROM, ALIAS_TO = "detr.detr.transformer.decoder.bbox_embed.", "detr.detr.bbox_embed."


class OracleMSDA:
    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return msda_core(value, shapes, loc, attn)


def compare(model, gold, loss_dict, outputs, indices_of, rtol_loss, rtol_norm, tol_grad):
    for lvl in range(6):
        for b in range(B):
            src, tgt = indices_of(lvl)[b]
            assert np.array_equal(src.cpu().numpy(), gold[f"lvl{lvl}_b{b}_src"])
            assert np.array_equal(tgt.cpu().numpy(), gold[f"lvl{lvl}_b{b}_tgt"])
    assert set(loss_dict) == {k[5:] for k in gold if k.startswith("loss/")}
    for k, v in loss_dict.items():
        np.testing.assert_allclose(float(v.detach()), float(gold["loss/" + k]), rtol=rtol_loss, atol=1e-5, err_msg=k)
    keys = [k[len("gradnorm/"):] for k in gold if k.startswith("gradnorm/")]
    assert len(keys) == 462
    worst = 0.0
    for k in keys:
        g = model.flat.params[k.replace(ALIAS_FROM, ALIAS_TO)].grad
        want = float(gold["gradnorm/" + k])
        rel = abs(float(g.norm()) - want) / max(want, 1e-4)     # time_attention_weights.bias: analytically zero gradient
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


def run(device, rtol_loss, rtol_norm, tol_grad, precision="fp32", batched=False):
    gold = dict(np.load(GOLD))
    sd = fill_state_dict(seqformer_r50_state_spec(), SEED)
    model = ST.SeqFormerTrainModel(sd, NF, device=device, dropout=0.0, precision=precision, batched=batched).train()
    frames = torch.stack(synthetic_clip(B * NF, H, W, SEED)).to(device)
    targets = [{k: v.to(device) for k, v in t.items()} for t in synthetic_clip_targets(B, NF, H, W, SEED)]
    seen = []
    orig = model.criterion.matcher
    model.criterion.matcher = lambda out, tg: seen.append(orig(out, tg)) or seen[-1]
    loss_dict, outputs = model(frames, targets)
    total = sum(model.weighted(loss_dict).values())
    np.testing.assert_allclose(float(total.detach()), float(gold["total"]), rtol=rtol_loss)
    total.backward()
    indices_of = (lambda lvl: outputs["indices_list"][lvl]) if batched else (lambda lvl: seen[lvl])
    return compare(model, gold, loss_dict, outputs, indices_of, rtol_loss, rtol_norm, tol_grad), model


@pytest.mark.timeout(900)
@pytest.mark.parametrize("batched", [False, True])
def test_train_forward_backward_matches_reference(monkeypatch, batched):
    monkeypatch.setattr(IT, "MSDeformAttnFunction", OracleMSDA)
    worst, model = run("cpu", 1e-4, 2e-3, 2e-3, batched=batched)
    print("worst grad-norm rel err", worst)
    # encoder layers carry an `output_proj_box` the encode path never uses (ms_deform_attn.py:89-124): zero gradient
    # (the scalar bias of `time_attention_weights` is softmax-shift-invariant: its gradient is 0 up to rounding and
    # may come out exactly 0 depending on the host's BLAS summation order)
    unused = [k for k, p in model.flat.params.items() if float(p.grad.abs().sum()) == 0
              and not k.endswith("time_attention_weights.bias")]
    assert unused and all("encoder" in k and "output_proj_box" in k for k in unused)


def test_hungarian_and_losses_on_reference_predictions():
    gold = dict(np.load(GOLD))
    targets = synthetic_clip_targets(B, NF, H, W, SEED)
    c = crit.SetCriterion(40, crit.default_weight_dict(), NF)
    for lvl in range(6):
        out = {"pred_logits": torch.from_numpy(gold[f"lvl{lvl}_pred_logits"]),
               "pred_boxes": torch.from_numpy(gold[f"lvl{lvl}_pred_boxes"])}
        ind = c.matcher(out, targets)
        for b in range(B):
            assert np.array_equal(ind[b][0].numpy(), gold[f"lvl{lvl}_b{b}_src"])
            assert np.array_equal(ind[b][1].numpy(), gold[f"lvl{lvl}_b{b}_tgt"])
        sfx = "" if lvl == 5 else f"_{lvl}"
        got = {**c.loss_labels(out, targets, ind, 6.0, lvl == 5), **c.loss_boxes(out, targets, ind, 6.0)}
        for k, v in got.items():
            np.testing.assert_allclose(float(v), float(gold["loss/" + k + sfx]), rtol=1e-5, atol=1e-6, err_msg=k)
    out = {"pred_masks": [torch.from_numpy(gold[f"final_pred_masks_b{b}"]) for b in range(B)]}
    ind = [(torch.from_numpy(gold[f"lvl5_b{b}_src"]), torch.from_numpy(gold[f"lvl5_b{b}_tgt"])) for b in range(B)]
    got = c.loss_masks(out, targets, ind, 6.0)
    np.testing.assert_allclose(float(got["loss_mask"]), float(gold["loss/loss_mask"]), rtol=1e-5)
    np.testing.assert_allclose(float(got["loss_dice"]), float(gold["loss/loss_dice"]), rtol=1e-5)


class _T:
    def __init__(self, t):
        self.tensor = t


class FakeInstances:
    """Duck type of detectron2.structures.Instances of one frame, as the YTVIS dataset mapper fills it."""

    def __init__(self, target, f, h, w, ids):
        from vnext_b200.idol.criterion import cxcywh_to_xyxy
        self.image_size = (h, w)
        self.gt_classes = target["labels"]
        self.gt_boxes = _T(cxcywh_to_xyxy(target["boxes"][:, f]) * torch.tensor([w, h, w, h], dtype=torch.float))
        self.gt_masks = _T(target["masks"][:, f])
        self.gt_ids = ids


def fake_batched_inputs():
    frames = synthetic_clip(B * NF, H, W, SEED)
    targets = synthetic_clip_targets(B, NF, H, W, SEED)
    ids = torch.arange(3)
    return [{"image": frames[v * NF:(v + 1) * NF],
             "instances": [FakeInstances(targets[v], f, H, W, ids) for f in range(NF)]} for v in range(B)], targets


def test_prepare_targets_matches_reference_format():
    from vnext_b200.seqformer.meta_arch import SeqFormer
    batched, targets = fake_batched_inputs()
    got = SeqFormer(device="cpu").prepare_targets(batched)
    for a, b in zip(targets, got):
        assert torch.equal(a["labels"], b["labels"]) and torch.allclose(a["boxes"], b["boxes"], atol=1e-6)
        assert torch.equal(a["masks"], b["masks"]) and b["size"].tolist() == [H, W]
    # an instance absent from every frame (gt_ids -1 everywhere) gets label 0 = class x valid (seqformer.py:291-296)
    for inst in batched[0]["instances"]:
        inst.gt_ids = torch.tensor([0, -1, 2])
    assert SeqFormer(device="cpu").prepare_targets(batched)[0]["labels"][1] == 0
