"""Host-side training logic of IDOL — assignment, contrastive sample selection and the loss terms.  The north star
keeps these in Python/PyTorch ("host code stays Python/PyTorch for the model wrapper and loss/matcher"); they run on
whatever device the predictions live on and launch no kernels of ours.  What each piece follows (paths under
/root/reference/projects/IDOL/idol):

  box helpers / GIoU      util/box_ops.py:18-87
  in_boxes_info           models/matcher.py:97-127        (same code again at models/pos_neg_select.py:131-160)
  dynamic_k               models/matcher.py:129-173 and models/pos_neg_select.py:162-199
  simota_match            models/matcher.py:47-95
  select_pos_neg          models/pos_neg_select.py:13-129
  focal / dice            models/segmentation_condInst.py:657-700
  SetCriterion            models/deformable_detr.py:225-498
  paired GIoU loss        fvcore.nn.giou_loss (third-party, unpinned by the reference; published formula, eps 1e-7)

Quirks that parity depends on are kept and marked `# ref quirk`.
"""
import random

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- boxes
def cxcywh_to_xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def xyxy_to_cxcywh(b):
    x0, y0, x1, y1 = b.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], -1)


def _area(b):
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def pairwise_iou(a, b):
    """[n,4] x [m,4] (xyxy) -> iou [n,m], union [n,m]."""
    wh = (torch.min(a[:, None, 2:], b[:, 2:]) - torch.max(a[:, None, :2], b[:, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = _area(a)[:, None] + _area(b) - inter
    return inter / union, union


def pairwise_giou(a, b):
    iou, union = pairwise_iou(a, b)
    wh = (torch.max(a[:, None, 2:], b[:, 2:]) - torch.min(a[:, None, :2], b[:, :2])).clamp(min=0)
    hull = wh[..., 0] * wh[..., 1]
    return iou - (hull - union) / (hull + 1e-7)


def paired_giou_loss(a, b, eps=1e-7):
    """1 - GIoU of box a[i] with box b[i] (xyxy)."""
    x1, y1, x2, y2 = a.unbind(-1)
    x1g, y1g, x2g, y2g = b.unbind(-1)
    ix1, iy1, ix2, iy2 = torch.max(x1, x1g), torch.max(y1, y1g), torch.min(x2, x2g), torch.min(y2, y2g)
    overlap = (iy2 > iy1) & (ix2 > ix1)
    inter = torch.where(overlap, (ix2 - ix1) * (iy2 - iy1), torch.zeros_like(x1))
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter
    hull = (torch.max(x2, x2g) - torch.min(x1, x1g)) * (torch.max(y2, y2g) - torch.min(y1, y1g))
    return 1 - (inter / (union + eps) - (hull - union) / (hull + eps))


# ----------------------------------------------------------------------------- simOTA
def in_boxes_info(boxes, gts, stride=32, radius=2.5):
    """boxes [nq,4], gts [n,4], both cxcywh normalised.  -> (query is a candidate for any gt [nq],
    query centre inside gt box AND inside gt centre window [nq,n])."""
    g = cxcywh_to_xyxy(gts)
    cx, cy = boxes[:, 0:1], boxes[:, 1:2]
    inside = (cx > g[:, 0]) & (cx < g[:, 2]) & (cy > g[:, 1]) & (cy < g[:, 3])
    r = radius / stride
    near = (cx > gts[:, 0] - r) & (cx < gts[:, 0] + r) & (cy > gts[:, 1] - r) & (cy < gts[:, 1] + r)
    return inside.any(1) | near.any(1), inside & near


def _ota_cost(prob, boxes, labels, gts):
    fg, both = in_boxes_info(boxes, gts)
    bx, gx = cxcywh_to_xyxy(boxes), cxcywh_to_xyxy(gts)
    iou, _ = pairwise_iou(bx, gx)
    alpha, gamma = 0.25, 2.0
    neg = (1 - alpha) * (prob ** gamma) * (-(1 - prob + 1e-8).log())
    pos = alpha * ((1 - prob) ** gamma) * (-(prob + 1e-8).log())
    cost = (pos[:, labels] - neg[:, labels]) + 3.0 * (-pairwise_giou(bx, gx)) + 100.0 * (~both)
    cost[~fg] = cost[~fg] + 10000.0
    return cost, iou


def _dynamic_k_matrix(cost, iou, num_gt, n_candidate):
    """The assignment matrix [nq, num_gt] of simOTA's dynamic-k step.  `cost` is modified in place like the
    reference's (the callers rely on that)."""
    match = torch.zeros_like(cost)
    ks = torch.clamp(torch.topk(iou, n_candidate, dim=0)[0].sum(0).int(), min=1).tolist()
    for g in range(num_gt):
        match[torch.topk(cost[:, g], k=ks[g], largest=False)[1], g] = 1.0
    multi = match.sum(1) > 1                         # ref quirk: this mask is never recomputed inside the loop below
    if multi.any():
        best = cost[multi].argmin(1)
        match[multi] = 0
        match[multi, best] = 1
    while (match.sum(0) == 0).any():
        cost[match.sum(1) > 0] += 100000.0
        for g in torch.nonzero(match.sum(0) == 0).flatten().tolist():
            match[cost[:, g].argmin(), g] = 1.0
        if (match.sum(1) > 1).any():
            best = cost[multi].argmin(1)
            match[multi] = 0
            match[multi, best] = 1
    return match


def simota_match(pred_logits, pred_boxes, targets):
    """-> (indices, matched): per image ((selected_query bool [nq], gt index of every selected query), best query
    of every gt)."""
    indices, matched = [], []
    with torch.no_grad():
        prob = pred_logits.sigmoid()
        for b, t in enumerate(targets):
            n = len(t["labels"])
            if n == 0:
                empty = torch.arange(0, 0).to(prob)
                indices.append((torch.zeros(prob.shape[1], device=prob.device) > 0, empty))
                matched.append(empty)
                continue
            gts = t["boxes"].reshape(n, 4)
            cost, iou = _ota_cost(prob[b], pred_boxes[b], t["labels"], gts)
            match = _dynamic_k_matrix(cost, iou, n, 10)
            sel = match.sum(1) > 0
            gt_of = match[sel].max(1)[1]
            cost[match == 0] = cost[match == 0] + float("inf")
            indices.append((sel, gt_of))
            matched.append(cost.min(0)[1])
    return indices, matched


def _ref_pos_neg(ref_boxes, ref_prob, gts, labels, valid):
    """Per key-frame instance: mask of positive reference queries (k=10) and of NON-negative ones (k=100)."""
    with torch.no_grad():
        if not bool(valid.all()):
            gts, labels = gts[valid], labels[valid]
        nv = int(valid.sum())
        if nv == 0:
            return [None] * len(valid), [None] * len(valid)
        cost, iou = _ota_cost(ref_prob, ref_boxes, labels, gts)
        # ref quirk: the second call sees the cost the first one already penalised in place
        pos = _dynamic_k_matrix(cost, iou, nv, 10)
        neg = _dynamic_k_matrix(cost, iou, nv, 100)
        pos_l, neg_l, j = [], [], 0
        for ok in valid.tolist():
            pos_l.append(pos[:, j] > 0 if ok else None)
            neg_l.append(neg[:, j] > 0 if ok else None)
            j += int(ok)
        return pos_l, neg_l


def select_pos_neg(ref_boxes, matched, ref_targets, det_targets, key_embed, ref_embed, ref_cls):
    """Contrastive items of every (valid) instance: dot products of its key embedding with positive and negative
    reference-frame embeddings (+ the cosine variant on a random subset of the negatives)."""
    items = []
    for b, (rt, mq) in enumerate(zip(ref_targets, matched)):
        n = len(rt["labels"])
        valid = rt["valid"]
        pos_l, neg_l = _ref_pos_neg(ref_boxes[b], ref_cls[b], rt["boxes"].reshape(n, 4), rt["labels"], valid)
        for i, ok in enumerate(valid.tolist()):
            if not ok:
                continue
            key = key_embed[b, mq[i]].unsqueeze(0)
            pos = ref_embed[b][pos_l[i]]
            neg = ref_embed[b][~neg_l[i]]
            one, zero = key.new_ones(()), key.new_zeros(())
            both = torch.cat([pos, neg], 0)
            label = torch.cat([one.repeat(len(pos)), zero.repeat(len(neg))], 0)
            contrast = both @ key.t()
            if len(pos) == 0:
                k = 10
            elif len(pos) * 10 >= len(neg):
                k = len(neg)
            else:
                k = len(pos) * 10
            ids = random.sample(list(range(0, len(neg))), k)
            aux = F.normalize(torch.cat([pos, neg[ids]], 0).float(), dim=1)
            aux_label = torch.cat([one.repeat(len(pos)), zero.repeat(k)], 0)
            cosine = aux @ F.normalize(key.float(), dim=1).t()
            items.append({"contrast": contrast, "label": label, "aux_consin": cosine, "aux_label": aux_label})
    return items


# ----------------------------------------------------------------------------- loss terms
def sigmoid_focal_loss(inputs, targets, num_boxes, alpha=0.25, gamma=2.0):
    prob = inputs.sigmoid()
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.mean(1).sum() / num_boxes


def dice_loss(inputs, targets, num_boxes):
    inputs = inputs.sigmoid().flatten(1)
    num = 2 * (inputs * targets).sum(1)
    den = inputs.sum(-1) + targets.sum(-1)
    return (1 - (num + 1) / (den + 1)).sum() / num_boxes


def pad_masks(mask_list, div=32):
    """[n_i,H_i,W_i] masks -> [len, max n, Hp, Wp] zero padded (util/misc.py:288-316 with split=False)."""
    n = max(m.shape[0] for m in mask_list)
    H = max(m.shape[1] for m in mask_list)
    W = max(m.shape[2] for m in mask_list)
    Hp, Wp = (H + div - 1) // div * div, (W + div - 1) // div * div
    out = mask_list[0].new_zeros((len(mask_list), n, Hp, Wp))
    for i, m in enumerate(mask_list):
        out[i, :m.shape[0], :m.shape[1], :m.shape[2]] = m
    return out


class SetCriterion:
    """loss_labels / loss_boxes / loss_masks / loss_reid and the deep-supervision loop."""

    def __init__(self, num_classes, weight_dict, focal_alpha=0.25, mask_out_stride=4, world_size_fn=None,
                 all_reduce_fn=None):
        self.num_classes = num_classes
        self.weight_dict = weight_dict
        self.focal_alpha = focal_alpha
        self.mask_out_stride = mask_out_stride
        self.world_size_fn = world_size_fn or (lambda: 1)
        self.all_reduce_fn = all_reduce_fn

    def matcher(self, outputs, targets):
        return simota_match(outputs["pred_logits"], outputs["pred_boxes"], targets)

    def loss_labels(self, outputs, targets, indices):
        logits = outputs["pred_logits"]
        classes = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=logits.device)
        n = 0
        for b, (sel, gt_of) in enumerate(indices):
            if len(gt_of) == 0:
                continue
            classes[b, sel] = targets[b]["labels"][gt_of]
            n += len(gt_of)
        num_boxes = n if n else 1                          # ref quirk: normalised by matched queries, not GT count
        onehot = torch.zeros(logits.shape[0], logits.shape[1], logits.shape[2] + 1, dtype=logits.dtype,
                             device=logits.device)
        onehot.scatter_(2, classes.unsqueeze(-1), 1)
        loss = sigmoid_focal_loss(logits, onehot[:, :, :-1], num_boxes, alpha=self.focal_alpha) * logits.shape[1]
        return {"loss_ce": loss}

    def loss_boxes(self, outputs, targets, indices):
        src, tgt = [], []
        for b, (sel, gt_of) in enumerate(indices):
            if len(gt_of) == 0:
                continue
            src.append(outputs["pred_boxes"][b][sel])
            tgt.append(targets[b]["boxes"][gt_of])
        if not src:
            z = outputs["pred_boxes"].sum() * 0
            return {"loss_bbox": z, "loss_giou": z}
        src, tgt = torch.cat(src), torch.cat(tgt)
        n = src.shape[0]
        return {"loss_bbox": F.l1_loss(src, tgt, reduction="none").sum() / n,
                "loss_giou": paired_giou_loss(cxcywh_to_xyxy(src), cxcywh_to_xyxy(tgt)).sum() / n}

    def loss_masks(self, outputs, targets, ref_targets, indices):
        src = outputs["pred_masks"]
        if isinstance(src, list):
            src = torch.cat(src, dim=1)[0]                                   # [sum n, 1, H/4, W/4]
        nk = len(targets)
        tm = pad_masks([t["masks"] for t in targets] + [t["masks"] for t in ref_targets])[:nk].to(src)
        s = self.mask_out_stride
        start = int(s // 2)
        full_h, full_w = tm.shape[-2:]
        tm = tm[:, :, start::s, start::s]
        assert tm.size(2) * s == full_h and tm.size(3) * s == full_w
        sel = []
        for b, (_, gt_of) in enumerate(indices):
            if len(gt_of) == 0:
                continue
            sel.append(tm[b][:targets[b]["masks"].shape[0]][gt_of].unsqueeze(1))
        if not sel:
            z = (src * 0).sum()
            return {"loss_mask": z, "loss_dice": z}
        tgt = torch.cat(sel)
        n = src.shape[0]
        assert src.shape == tgt.shape
        src, tgt = src.flatten(1), tgt.flatten(1)
        return {"loss_mask": sigmoid_focal_loss(src, tgt, n), "loss_dice": dice_loss(src, tgt, n)}

    def loss_reid(self, outputs):
        items = outputs["pred_qd"]
        if len(items) == 0:
            z = outputs["pred_logits"].sum() * 0
            return {"loss_reid": z, "loss_reid_aux": z}
        contrast, aux = 0, 0
        for it in items:
            pred = it["contrast"].permute(1, 0)                              # [1, n]
            label = it["label"].unsqueeze(0)
            pos, neg = label == 1, label == 0
            ppos = torch.where(pos, pred, pred.new_full((), float("inf")))
            pneg = torch.where(neg, pred, pred.new_full((), float("-inf")))
            n = pred.shape[1]
            x = F.pad(pneg.repeat(1, n) - torch.repeat_interleave(ppos, n, dim=1), (0, 1), "constant", 0)
            contrast = contrast + torch.logsumexp(x, dim=1)
            ap = it["aux_consin"].permute(1, 0)
            aux = aux + (torch.abs(ap - it["aux_label"].unsqueeze(0)) ** 2).mean()
        return {"loss_reid": contrast.sum() / len(items), "loss_reid_aux": aux / len(items)}

    def _level(self, outputs, targets, ref_targets, indices, reid):
        losses = {}
        losses.update(self.loss_labels(outputs, targets, indices))
        losses.update(self.loss_boxes(outputs, targets, indices))
        losses.update(self.loss_masks(outputs, targets, ref_targets, indices))
        if reid:
            losses.update(self.loss_reid(outputs))
        return losses

    def __call__(self, outputs, targets, ref_targets, indices_list):
        # models/deformable_detr.py:470-476: the cross-rank mean GT count is computed (a 4-byte all-reduce) but every
        # term renormalises by its own matched count, so it does not reach the values.
        if self.all_reduce_fn is not None:
            nb = torch.as_tensor([float(sum(len(t["labels"]) for t in targets))], device=outputs["pred_logits"].device)
            self.all_reduce_fn(nb)
        losses = self._level(outputs, targets, ref_targets, indices_list[-1], reid=True)
        for i, aux in enumerate(outputs.get("aux_outputs", [])):
            for k, v in self._level(aux, targets, ref_targets, indices_list[i], reid=False).items():
                losses[f"{k}_{i}"] = v
        return losses


def default_weight_dict(class_w=2.0, l1_w=5.0, giou_w=2.0, reid_w=2.0, mask_w=2.0, dice_w=5.0, dec_layers=6,
                        deep_supervision=True):
    """idol.py:160-176."""
    w = {"loss_ce": class_w, "loss_bbox": l1_w, "loss_giou": giou_w, "loss_reid": reid_w,
         "loss_reid_aux": reid_w * 1.5, "loss_mask": mask_w, "loss_dice": dice_w}
    if deep_supervision:
        base = dict(w)
        for i in range(dec_layers - 1):
            w.update({f"{k}_{i}": v for k, v in base.items()})
    return w
