"""Golden vectors of SeqFormer's TRAINING forward/backward from the REFERENCE's own modules (build container only).

`CondInst_segm.forward(samples, gt_targets, criterion, train=True)` (projects/SeqFormer/seqformer/models/
segmentation_condInst.py:69-204) with the reference's `SetCriterion` (models/deformable_detr.py:231-440) and
`HungarianMatcher` (models/matcher.py:24-91, scipy linear_sum_assignment) on CPU, weighted loss sum back-propagated
(seqformer.py:116-130, :217-225).  Mounting / stand-ins as in make_seqformer_golden.py; `fvcore.nn.giou_loss` restated
from its published formula (see make_idol_train_golden.py).  Dropout 0."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_seqformer_golden as msg  # noqa: E402
from make_idol_train_golden import giou_loss  # noqa: E402
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_clip_targets  # noqa: E402

B, NF, H, W, SEED = 2, 2, 96, 128, 4
WEIGHTS = {"loss_ce": 2.0, "loss_bbox": 5.0, "loss_giou": 2.0, "loss_mask": 2.0, "loss_dice": 5.0}   # S/config.py:21-25
FULL_GRADS = ["detr.detr.class_embed.5.bias", "detr.detr.transformer.level_embed", "detr.controller.layers.2.bias",
              "detr.detr.transformer.reference_points.weight", "detr.detr.bbox_embed.0.layers.2.weight",
              "detr.detr.transformer.decoder.layers.5.norm1_box.weight",
              "detr.detr.transformer.decoder.layers.3.time_attention_weights.weight", "detr.mask_head.lay2.bias"]
STRIDED_GRADS = ["detr.detr.transformer.encoder.layers.0.self_attn.sampling_offsets.weight",
                 "detr.detr.transformer.decoder.layers.2.cross_attn.output_proj_box.weight",
                 "detr.detr.transformer.decoder.layers.4.self_attn_box.in_proj_weight",
                 "detr.detr.backbone.0.backbone.res5.2.conv3.weight", "detr.detr.query_embed.weight"]


def main():
    import importlib
    msg.mount_reference()
    sys.modules["fvcore.nn"].giou_loss = giou_loss
    dd = importlib.import_module("refseq.models.deformable_detr")
    mm = importlib.import_module("refseq.models.matcher")
    dd.giou_loss = mm.giou_loss = giou_loss

    model = msg.build_reference_model(300, NF)
    for m in model.detr.detr.transformer.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    model.train()
    spec, back = msg.reference_state_spec(model)
    sd = fill_state_dict(spec, SEED)
    model.load_state_dict({back[k]: v for k, v in sd.items()}, strict=True)
    fwd = {v: k for k, v in back.items()}
    for k, p in model.named_parameters():
        d2 = fwd[k]
        if ".backbone.stem." in d2 or ".backbone.res2." in d2:
            p.requires_grad_(False)

    frames = synthetic_clip(B * NF, H, W, SEED)
    mean, std = torch.tensor(msg.PIXEL_MEAN).view(3, 1, 1), torch.tensor(msg.PIXEL_STD).view(3, 1, 1)
    clip = msg.Clip([(f - mean) / std for f in frames])
    targets = synthetic_clip_targets(B, NF, H, W, SEED)

    weight_dict = dict(WEIGHTS)
    for i in range(5):
        weight_dict.update({f"{k}_{i}": v for k, v in WEIGHTS.items()})
    matcher = mm.HungarianMatcher(multi_frame=True, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0)
    criterion = dd.SetCriterion(40, matcher, weight_dict, ["labels", "boxes", "masks"], mask_out_stride=4,
                                focal_alpha=0.25, num_frames=NF)
    captured, level = {}, [0]
    orig = matcher.forward

    def spy(outputs, tg, nf, vr):
        ind = orig(outputs, tg, nf, vr)
        captured[f"lvl{level[0]}_pred_logits"] = outputs["pred_logits"].detach().numpy().copy()
        captured[f"lvl{level[0]}_pred_boxes"] = outputs["pred_boxes"].detach().numpy().copy()
        for b, (i, j) in enumerate(ind):
            captured[f"lvl{level[0]}_b{b}_src"] = i.numpy().copy()
            captured[f"lvl{level[0]}_b{b}_tgt"] = j.numpy().copy()
        level[0] += 1
        return ind
    matcher.forward = spy

    outputs, loss_dict = model.detr(clip, targets, criterion, train=True)
    total = sum(loss_dict[k] * weight_dict[k] for k in loss_dict if k in weight_dict)
    total.backward()
    out = dict(captured)
    for k, v in loss_dict.items():
        out["loss/" + k] = np.array(float(v.detach()))
    out["total"] = np.array(float(total.detach()))
    for i, m in enumerate(outputs["pred_masks"]):
        out[f"final_pred_masks_b{i}"] = m.detach().numpy()
    n = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        n += 1
        d2 = fwd[k]
        out["gradnorm/" + d2] = np.array(float(p.grad.norm()))
        if d2 in FULL_GRADS:
            out["grad/" + d2] = p.grad.numpy()
        if d2 in STRIDED_GRADS:
            out["gradstrided/" + d2] = p.grad.flatten()[::97].numpy()
    print("params with grad:", n, "total", float(total.detach()))
    print({k: round(float(v), 5) for k, v in loss_dict.items() if not k[-1].isdigit()})
    path = os.path.join(HERE, "seqformer_train_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
