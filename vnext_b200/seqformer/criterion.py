"""Host-side training logic of SeqFormer: clip-level Hungarian assignment and the loss terms (Python/PyTorch, as the
north star prescribes for matcher and losses).  Follows, under /root/reference/projects/SeqFormer/seqformer:

  hungarian_match     models/matcher.py:53-91     (scipy.optimize.linear_sum_assignment, like the reference)
  SetCriterion        models/deformable_detr.py:231-440
  focal / dice        models/segmentation_condInst.py (same code as IDOL's, restated in vnext_b200/idol/criterion.py)
  paired GIoU loss    fvcore.nn.giou_loss (third-party, unpinned; published formula)
"""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from ..idol.criterion import (cxcywh_to_xyxy, dice_loss, pad_masks, paired_giou_loss, pairwise_giou,
                              sigmoid_focal_loss)


def hungarian_match(pred_logits, pred_boxes, targets, nf, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0):
    """pred_logits [bs,nq,K], pred_boxes [bs,nf,nq,4]; targets: per video {"labels" [n], "boxes" [n,nf,4]}.
    -> per video (query indices, target indices), int64 CPU tensors."""
    with torch.no_grad():
        bs, nq = pred_logits.shape[:2]
        prob = pred_logits.flatten(0, 1).sigmoid()
        tgt_ids = torch.cat([t["labels"] for t in targets])
        n = len(tgt_ids)
        tgt_bbox = torch.cat([t["boxes"] for t in targets]).reshape(n, nf, 4)
        out_bbox = pred_boxes.permute(0, 2, 1, 3).flatten(0, 1)                              # [bs*nq, nf, 4]
        c_bbox = torch.cdist(out_bbox.flatten(1, 2), tgt_bbox.flatten(1, 2))
        tgt_bbox = torch.clip(tgt_bbox, min=1e-7, max=1)
        c_giou = 0
        for i in range(nf):
            c_giou = c_giou - pairwise_giou(cxcywh_to_xyxy(out_bbox[:, i]), cxcywh_to_xyxy(tgt_bbox[:, i]))
        c_giou = c_giou / nf
        alpha, gamma = 0.25, 2.0
        neg = (1 - alpha) * (prob ** gamma) * (-(1 - prob + 1e-8).log())
        pos = alpha * ((1 - prob) ** gamma) * (-(prob + 1e-8).log())
        c_class = pos[:, tgt_ids] - neg[:, tgt_ids]
        C = (cost_bbox * c_bbox + cost_class * c_class + cost_giou * c_giou).view(bs, nq, -1).cpu()
        sizes = [len(t["labels"]) for t in targets]
        out = []
        for i, c in enumerate(C.split(sizes, -1)):
            r, col = linear_sum_assignment(c[i])
            out.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(col, dtype=torch.int64)))
        return out


class SetCriterion:
    def __init__(self, num_classes, weight_dict, num_frames, focal_alpha=0.25, mask_out_stride=4, cost_class=2.0,
                 cost_bbox=5.0, cost_giou=2.0, world_size_fn=None, all_reduce_fn=None):
        self.num_classes, self.weight_dict, self.nf = num_classes, weight_dict, num_frames
        self.focal_alpha, self.mask_out_stride = focal_alpha, mask_out_stride
        self.costs = (cost_class, cost_bbox, cost_giou)
        self.world_size_fn = world_size_fn or (lambda: 1)
        self.all_reduce_fn = all_reduce_fn

    def matcher(self, outputs, targets):
        return hungarian_match(outputs["pred_logits"], outputs["pred_boxes"], targets, self.nf, *self.costs)

    @staticmethod
    def _src_idx(indices):
        return (torch.cat([torch.full_like(s, i) for i, (s, _) in enumerate(indices)]),
                torch.cat([s for s, _ in indices]))

    @staticmethod
    def _tgt_idx(indices):
        return (torch.cat([torch.full_like(t, i) for i, (_, t) in enumerate(indices)]),
                torch.cat([t for _, t in indices]))

    def loss_labels(self, outputs, targets, indices, num_boxes, log):
        logits = outputs["pred_logits"]
        idx = self._src_idx(indices)
        classes_o = torch.cat([t["labels"][j.to(t["labels"].device)] for t, (_, j) in zip(targets, indices)])
        classes = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        classes[idx[0].to(logits.device), idx[1].to(logits.device)] = classes_o
        onehot = torch.zeros(logits.shape[0], logits.shape[1], logits.shape[2] + 1, dtype=logits.dtype, device=logits.device)
        onehot.scatter_(2, classes.unsqueeze(-1), 1)
        out = {"loss_ce": sigmoid_focal_loss(logits, onehot[:, :, :-1], num_boxes, alpha=self.focal_alpha) * logits.shape[1]}
        if log:        # top-1 error of the matched queries (util/misc.py accuracy), logged with the last level only
            sel = logits[idx[0].to(logits.device), idx[1].to(logits.device)]
            if classes_o.numel() == 0:
                out["class_error"] = 100 - torch.zeros([], device=logits.device)
            else:
                acc = (sel.argmax(-1) == classes_o).float().sum() * (100.0 / classes_o.numel())
                out["class_error"] = 100 - acc
        return out

    def loss_boxes(self, outputs, targets, indices, num_boxes):
        dev = outputs["pred_boxes"].device
        idx = self._src_idx(indices)
        src = outputs["pred_boxes"].permute(0, 2, 1, 3)[idx[0].to(dev), idx[1].to(dev)]      # [n, nf, 4]
        n, nf = src.shape[:2]
        tgt = torch.cat([t["boxes"].reshape(-1, nf, 4)[j.to(dev)] for t, (_, j) in zip(targets, indices)], 0)
        l1 = F.l1_loss(src.flatten(1, 2), tgt.flatten(1, 2), reduction="none") / nf
        giou = paired_giou_loss(cxcywh_to_xyxy(src.flatten(0, 1)), cxcywh_to_xyxy(tgt.flatten(0, 1))) / nf
        return {"loss_bbox": l1.sum() / num_boxes, "loss_giou": giou.sum() / num_boxes}

    def loss_masks(self, outputs, targets, indices, num_boxes):
        src = outputs["pred_masks"]
        if isinstance(src, list):
            src = torch.cat(src, dim=1)[0]                                                   # [n, nf, h, w]
        bs = len(targets)
        tm = pad_masks([t["masks"].flatten(0, 1) for t in targets]).to(src)                  # [bs, n_max*nf, Hp, Wp]
        s = self.mask_out_stride
        full_h, full_w = tm.shape[-2:]
        tm = tm[:, :, int(s // 2)::s, int(s // 2)::s]
        assert tm.size(2) * s == full_h and tm.size(3) * s == full_w
        nf = src.shape[1]
        tm = tm.reshape(bs, -1, nf, tm.shape[-2], tm.shape[-1])
        ti = self._tgt_idx(indices)
        tgt = tm[ti[0].to(src.device), ti[1].to(src.device)].flatten(1)
        src = src.flatten(1)
        if len(tgt) == 0:
            z = (src * 0).sum()
            return {"loss_mask": z, "loss_dice": z}
        return {"loss_mask": sigmoid_focal_loss(src, tgt, num_boxes), "loss_dice": dice_loss(src, tgt, num_boxes)}

    def __call__(self, outputs, targets, indices_list):
        dev = outputs["pred_logits"].device
        nb = torch.as_tensor([float(sum(len(t["labels"]) for t in targets))], device=dev)
        if self.all_reduce_fn is not None:
            self.all_reduce_fn(nb)                                                           # deformable_detr.py:417-419
        num_boxes = torch.clamp(nb / self.world_size_fn(), min=1).item()

        def level(out, indices, log):
            d = self.loss_labels(out, targets, indices, num_boxes, log)
            d.update(self.loss_boxes(out, targets, indices, num_boxes))
            d.update(self.loss_masks(out, targets, indices, num_boxes))
            return d
        losses = level(outputs, indices_list[-1], True)
        for i, aux in enumerate(outputs.get("aux_outputs", [])):
            for k, v in level(aux, indices_list[i], False).items():
                losses[f"{k}_{i}"] = v
        return losses


def default_weight_dict(class_w=2.0, l1_w=5.0, giou_w=2.0, mask_w=2.0, dice_w=5.0, dec_layers=6, deep_supervision=True):
    """seqformer.py:116-130."""
    w = {"loss_ce": class_w, "loss_bbox": l1_w, "loss_giou": giou_w, "loss_mask": mask_w, "loss_dice": dice_w}
    if deep_supervision:
        base = dict(w)
        for i in range(dec_layers - 1):
            w.update({f"{k}_{i}": v for k, v in base.items()})
    return w
