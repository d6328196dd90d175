"""Golden vectors of IDOL's TRAINING forward/backward from the REFERENCE's own modules (build container only).

    python tests/golden/make_idol_train_golden.py

Runs `CondInst_segm.forward(samples, det_targets, ref_targets, criterion, train=True)`
(projects/IDOL/idol/models/segmentation_condInst.py:77-231) with the reference's own `SetCriterion`
(models/deformable_detr.py:225-498), simOTA `HungarianMatcher` (models/matcher.py:19-174) and `select_pos_neg`
(models/pos_neg_select.py:13-199) on CPU, then back-propagates the weighted loss sum (idol.py:160-176, :205-209).
Mounting and stand-ins are those of make_idol_golden.py, plus `fvcore.nn.giou_loss`: fvcore is a third-party
dependency, absent offline and unpinned by the reference (requirements.txt); its published per-pair GIoU loss
(fvcore/nn/giou_loss.py, eps=1e-7) is restated below.  Dropout is 0 (the stochastic part cannot be pinned across
implementations); `random.seed(0)` pins select_pos_neg's negative sampling (pos_neg_select.py:54).

Stored: every (unweighted) loss, the simOTA assignment of every decoder level, the intermediate predictions the
criterion consumed (so the host-side criterion restatement can be pinned on CPU), the gradient L2 norm of EVERY
trainable parameter and full / strided gradients of a few."""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_idol_golden as mig  # noqa: E402
from make_seqformer_golden import PIXEL_MEAN, PIXEL_STD, Clip, tv_to_d2_key  # noqa: E402
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_targets  # noqa: E402

B, H, W, SEED = 2, 96, 128, 3          # 2 videos x (key, ref) frames
WEIGHTS = {"loss_ce": 2.0, "loss_bbox": 5.0, "loss_giou": 2.0, "loss_reid": 2.0, "loss_reid_aux": 3.0,
           "loss_mask": 2.0, "loss_dice": 5.0}                       # config.py:23-28, idol.py:160-166
FULL_GRADS = ["detr.detr.class_embed.5.bias", "detr.detr.transformer.level_embed", "detr.controller.layers.2.bias",
              "detr.detr.transformer.reference_points.weight", "detr.detr.bbox_embed.0.layers.2.weight",
              "detr.reid_embed_head.layers.2.bias", "detr.detr.transformer.decoder.layers.5.norm1.weight",
              "detr.detr.input_proj.3.1.weight", "detr.mask_head.lay2.bias"]
STRIDED_GRADS = ["detr.detr.transformer.encoder.layers.0.self_attn.sampling_offsets.weight",
                 "detr.detr.transformer.encoder.layers.5.self_attn.value_proj.weight",
                 "detr.detr.transformer.decoder.layers.2.cross_attn.attention_weights.weight",
                 "detr.detr.backbone.0.backbone.res5.2.conv3.weight",
                 "detr.detr.backbone.0.backbone.res3.0.conv1.weight", "detr.detr.query_embed.weight",
                 "detr.reid_embed_head.layers.0.weight"]


def giou_loss(boxes1, boxes2, reduction="none", eps=1e-7):
    x1, y1, x2, y2 = boxes1.unbind(-1)
    x1g, y1g, x2g, y2g = boxes2.unbind(-1)
    xk1, yk1 = torch.max(x1, x1g), torch.max(y1, y1g)
    xk2, yk2 = torch.min(x2, x2g), torch.min(y2, y2g)
    inter = torch.zeros_like(x1)
    m = (yk2 > yk1) & (xk2 > xk1)
    inter[m] = (xk2[m] - xk1[m]) * (yk2[m] - yk1[m])
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter
    iou = inter / (union + eps)
    xc1, yc1 = torch.min(x1, x1g), torch.min(y1, y1g)
    xc2, yc2 = torch.max(x2, x2g), torch.max(y2, y2g)
    area_c = (xc2 - xc1) * (yc2 - yc1)
    loss = 1 - (iou - (area_c - union) / (area_c + eps))
    return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss


def main():
    mig.mount()
    sys.modules["fvcore.nn"].giou_loss = giou_loss
    import importlib
    dd = importlib.import_module("refidol.models.deformable_detr")
    dd.giou_loss = giou_loss
    from refidol.models.matcher import HungarianMatcher

    model = mig.build()
    tr = model.detr.detr.transformer
    for m in tr.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    model.train()
    spec, back = {}, {}
    for k, v in model.state_dict().items():
        pre = "detr.detr.backbone.0.body."
        k2 = "detr.detr.backbone.0.backbone." + tv_to_d2_key(k[len(pre):]) if k.startswith(pre) else k
        spec[k2], back[k2] = v, k
    sd = fill_state_dict(spec, SEED)
    model.load_state_dict({back[k]: v for k, v in sd.items()}, strict=True)
    fwd = {v: k for k, v in back.items()}
    for k, p in model.named_parameters():                     # FREEZE_AT=2: stem + res2 (resnet.py:613-650)
        d2 = fwd[k]
        if ".backbone.stem." in d2 or ".backbone.res2." in d2:
            p.requires_grad_(False)

    frames = synthetic_clip(2 * B, H, W, SEED)
    mean, std = torch.tensor(PIXEL_MEAN).view(3, 1, 1), torch.tensor(PIXEL_STD).view(3, 1, 1)
    clip = Clip([(f - mean) / std for f in frames])
    clip.image_sizes = [(H, W)] * (2 * B)
    det_targets, ref_targets = synthetic_targets(B, H, W, SEED)

    weight_dict = dict(WEIGHTS)
    for i in range(5):
        weight_dict.update({f"{k}_{i}": v for k, v in WEIGHTS.items()})
    matcher = HungarianMatcher(multi_frame=True, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0)
    criterion = dd.SetCriterion(40, matcher, weight_dict, ["labels", "boxes", "masks", "reid"], mask_out_stride=4,
                                focal_alpha=0.25, num_frames=1)

    captured = {}
    orig_matcher_forward = matcher.forward
    level = [0]

    def spy(outputs, targets, nf=1):
        captured[f"lvl{level[0]}_pred_logits"] = outputs["pred_logits"].detach().numpy().copy()
        captured[f"lvl{level[0]}_pred_boxes"] = outputs["pred_boxes"].detach().numpy().copy()
        indices, matched = orig_matcher_forward(outputs, targets, nf)
        for b, ((sel, gt), mq) in enumerate(zip(indices, matched)):
            captured[f"lvl{level[0]}_b{b}_selected"] = sel.numpy().copy()
            captured[f"lvl{level[0]}_b{b}_gt"] = gt.numpy().copy()
            captured[f"lvl{level[0]}_b{b}_matched"] = mq.numpy().copy()
        level[0] += 1
        return indices, matched
    matcher.forward = spy

    segm = importlib.import_module("refidol.models.segmentation_condInst")
    orig_select = segm.select_pos_neg

    def spy_select(ref_box, all_indices, targets, dets, embed_head, hs_key, hs_ref, ref_cls):
        captured["reid_ref_box"] = ref_box.detach().numpy().copy()
        captured["reid_ref_cls"] = ref_cls.detach().numpy().copy()
        captured["reid_key_embed"] = embed_head(hs_key).detach().numpy().copy()
        captured["reid_ref_embed"] = embed_head(hs_ref).detach().numpy().copy()
        return orig_select(ref_box, all_indices, targets, dets, embed_head, hs_key, hs_ref, ref_cls)
    segm.select_pos_neg = spy_select

    random.seed(0)
    outputs, loss_dict = model.detr(clip, det_targets, ref_targets, criterion, train=True)
    total = sum(loss_dict[k] * weight_dict[k] for k in loss_dict if k in weight_dict)
    total.backward()

    out = dict(captured)
    for k, v in loss_dict.items():
        out["loss/" + k] = np.array(float(v))
    out["total"] = np.array(float(total))
    out["num_contrast_items"] = np.array(len(outputs["pred_qd"]))
    for i, m in enumerate(outputs["pred_masks"]):
        out[f"final_pred_masks_b{i}"] = m.detach().numpy()
    for i, q in enumerate(outputs["pred_qd"]):
        out[f"qd{i}_contrast"] = q["contrast"].detach().numpy()
        out[f"qd{i}_label"] = q["label"].numpy()
        out[f"qd{i}_aux_cos"] = q["aux_consin"].detach().numpy()
        out[f"qd{i}_aux_label"] = q["aux_label"].numpy()
    n_tr = 0
    for k, p in model.named_parameters():
        d2 = fwd[k]
        if p.grad is None:
            continue
        n_tr += 1
        out["gradnorm/" + d2] = np.array(float(p.grad.norm()))
        if d2 in FULL_GRADS:
            out["grad/" + d2] = p.grad.numpy()
        if d2 in STRIDED_GRADS:
            out["gradstrided/" + d2] = p.grad.flatten()[::97].numpy()
    print("trainable with grad:", n_tr, " total loss", float(total))
    print({k: round(float(v), 5) for k, v in loss_dict.items() if "_" not in k[-2:]})
    path = os.path.join(HERE, "idol_train_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
