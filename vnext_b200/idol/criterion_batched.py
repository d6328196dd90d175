"""Device-batched form of IDOL's assignment and losses ("matchers without host round-trips").

Same mathematics as criterion.py (which restates the reference line by line and stays the parity anchor), organised for
the GPU: the simOTA problems of ALL decoder levels and images (6 x B) are one padded [P, queries, max_gt] tensor
problem, every loss term is a masked dense reduction over [levels, images, queries], and the only host round-trips per
step are (1) one flag "did any ground truth end up without a query" (the reference's rare `while` repair loop,
matcher.py:151-162 — those problems, if any, are redone by the sequential code), (2) one `nonzero` listing the selected
(level, image, query) triples for the dynamic mask head and (3) the positive/negative counts `random.sample` needs
(pos_neg_select.py:47-54).  The reference issues ~9 000 tiny kernels and several hundred syncs for the same work.
"""
import random

import torch
import torch.nn.functional as F

from . import criterion as seq

INF = float("inf")


def pad_targets(targets, device):
    """list of per-image dicts -> boxes [B,G,4], labels [B,G], exists [B,G] (G = max instances, >= 1)."""
    B = len(targets)
    G = max(1, max(len(t["labels"]) for t in targets))
    boxes = torch.zeros(B, G, 4, device=device)
    labels = torch.zeros(B, G, dtype=torch.long, device=device)
    exists = torch.zeros(B, G, dtype=torch.bool, device=device)
    for b, t in enumerate(targets):
        n = len(t["labels"])
        if n:
            boxes[b, :n] = t["boxes"].reshape(n, 4)
            labels[b, :n] = t["labels"]
            exists[b, :n] = True
    return boxes, labels, exists


def _xyxy(b):
    return torch.cat([b[..., :2] - 0.5 * b[..., 2:], b[..., :2] + 0.5 * b[..., 2:]], -1)


def ota_cost(prob, boxes, gt_boxes, gt_labels, exists):
    """prob [P,nq,K] (sigmoid), boxes [P,nq,4], gt_* [P,G,...] -> cost [P,nq,G] (+inf in padded columns), iou [P,nq,G]."""
    g = _xyxy(gt_boxes)[:, None]                                                  # [P,1,G,4]
    q = _xyxy(boxes)[:, :, None]                                                  # [P,nq,1,4]
    cx, cy = boxes[:, :, None, 0], boxes[:, :, None, 1]
    inside = (cx > g[..., 0]) & (cx < g[..., 2]) & (cy > g[..., 1]) & (cy < g[..., 3])
    r = 2.5 / 32
    gc = gt_boxes[:, None]
    near = (cx > gc[..., 0] - r) & (cx < gc[..., 0] + r) & (cy > gc[..., 1] - r) & (cy < gc[..., 1] + r)
    ex = exists[:, None]
    inside, near = inside & ex, near & ex
    fg = inside.any(2) | near.any(2)
    wh = (torch.min(q[..., 2:], g[..., 2:]) - torch.max(q[..., :2], g[..., :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    area_q = (q[..., 2] - q[..., 0]) * (q[..., 3] - q[..., 1])
    area_g = (g[..., 2] - g[..., 0]) * (g[..., 3] - g[..., 1])
    union = area_q + area_g - inter
    iou = inter / union
    hwh = (torch.max(q[..., 2:], g[..., 2:]) - torch.min(q[..., :2], g[..., :2])).clamp(min=0)
    hull = hwh[..., 0] * hwh[..., 1]
    giou = iou - (hull - union) / (hull + 1e-7)
    alpha, gamma = 0.25, 2.0
    neg = (1 - alpha) * (prob ** gamma) * (-(1 - prob + 1e-8).log())
    pos = alpha * ((1 - prob) ** gamma) * (-(prob + 1e-8).log())
    idx = gt_labels[:, None].expand(-1, prob.shape[1], -1)
    cost = (pos.gather(2, idx) - neg.gather(2, idx)) + 3.0 * (-giou) + 100.0 * (~(inside & near))
    cost = cost + 10000.0 * (~fg)[..., None]
    cost = torch.where(ex, cost, cost.new_full((), INF))
    return cost, torch.where(ex, iou, iou.new_zeros(()))


def dynamic_k(cost, iou, exists, n_candidate):
    """-> match [P,nq,G] bool, unmatched [P] bool (problems the reference would send through its repair loop)."""
    k = min(n_candidate, iou.shape[1])
    ks = torch.clamp(torch.topk(iou, k, dim=1)[0].sum(1).int(), min=1)            # [P,G]
    rank = cost.argsort(dim=1).argsort(dim=1)
    match = (rank < ks[:, None]) & exists[:, None]
    multi = match.sum(2) > 1
    best = F.one_hot(cost.argmin(2), cost.shape[2]).bool()
    match = torch.where(multi[..., None], best, match)
    unmatched = ((match.sum(1) == 0) & exists).any(1)
    return match, unmatched


def simota_batched(logits, boxes, gt_boxes, gt_labels, exists):
    """All problems at once.  logits [P,nq,K], boxes [P,nq,4]; gt tensors [P,G,...].
    -> sel [P,nq] bool, gt_of [P,nq] long (meaningful where sel), best_query [P,G] long (meaningful where exists)."""
    with torch.no_grad():
        cost, iou = ota_cost(logits.sigmoid(), boxes, gt_boxes, gt_labels, exists)
        match, bad = dynamic_k(cost, iou, exists, 10)
        sel = match.any(2)
        gt_of = match.float().argmax(2)
        best_q = torch.where(match, cost, cost.new_full((), INF)).argmin(1)
        if bool(bad.any()):                                                      # the one sync of the matcher
            for p in torch.nonzero(bad).flatten().tolist():                      # rare: the reference's repair loop
                n = int(exists[p].sum())
                ind, mq = seq.simota_match(logits[p:p + 1], boxes[p:p + 1],
                                           [{"labels": gt_labels[p, :n], "boxes": gt_boxes[p, :n]}])
                sel[p] = ind[0][0]
                gt_of[p][ind[0][0]] = ind[0][1]
                best_q[p, :n] = mq[0]
        return sel, gt_of, best_q


def focal_sum(inputs, targets, alpha=0.25, gamma=2.0):
    """Un-normalised sigmoid focal loss, elementwise (segmentation_condInst.py:675-698)."""
    prob = inputs.sigmoid()
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    return (alpha * targets + (1 - alpha) * (1 - targets)) * loss if alpha >= 0 else loss


NEG = -1e30        # "minus infinity" that keeps logsumexp's gradient finite on fully masked rows


class BatchedCriterion:
    """All 32 loss terms from stacked per-level predictions.  Values equal criterion.SetCriterion's (same golden)."""

    def __init__(self, num_classes, focal_alpha=0.25, mask_out_stride=4):
        self.num_classes, self.focal_alpha, self.mask_out_stride = num_classes, focal_alpha, mask_out_stride

    def detection_losses(self, logits, boxes, det_targets):
        """logits [L,B,nq,K], boxes [L,B,nq,4] -> (dict of [L] tensors, assignment dict)."""
        L, B, nq, K = logits.shape
        gtb, gtl, ex = pad_targets(det_targets, logits.device)
        rep = lambda t: t.repeat(L, *([1] * (t.dim() - 1)))
        gtb_r, gtl_r, ex_r = rep(gtb), rep(gtl), rep(ex)
        lf, bf = logits.flatten(0, 1), boxes.flatten(0, 1)
        sel, gt_of, best_q = simota_batched(lf, bf, gtb_r, gtl_r, ex_r)
        n_l = sel.view(L, -1).sum(1).clamp(min=1).to(logits.dtype)
        cls_t = torch.where(sel, gtl_r.gather(1, gt_of), gtl_r.new_full((), K))
        onehot = F.one_hot(cls_t, K + 1)[..., :K].to(logits.dtype)
        out = {"loss_ce": focal_sum(lf, onehot, self.focal_alpha).view(L, -1).sum(1) / n_l}
        tgt = gtb_r.gather(1, gt_of[..., None].expand(-1, -1, 4))
        zero = bf.new_zeros(())
        out["loss_bbox"] = torch.where(sel, (bf - tgt).abs().sum(-1), zero).view(L, -1).sum(1) / n_l
        out["loss_giou"] = torch.where(sel, seq.paired_giou_loss(_xyxy(bf), _xyxy(tgt)), zero).view(L, -1).sum(1) / n_l
        return out, {"sel": sel.view(L, B, nq), "gt_of": gt_of.view(L, B, nq), "best_q": best_q.view(L, B, -1)}

    def mask_targets(self, det_targets, ref_targets, device):
        """Key-frame masks padded like the reference pads key+ref together, sampled at the mask stride -> [B,G,h,w]."""
        nk = len(det_targets)
        tm = seq.pad_masks([t["masks"] for t in det_targets] + [t["masks"] for t in ref_targets])[:nk]
        s = self.mask_out_stride
        full_h, full_w = tm.shape[-2:]
        tm = tm[:, :, int(s // 2)::s, int(s // 2)::s]
        assert tm.size(2) * s == full_h and tm.size(3) * s == full_w
        return tm.to(device=device, dtype=torch.float32)

    def mask_losses(self, mask_logits, target, level_of, L):
        """mask_logits / target [n, h, w] of the selected (level, image, query) triples -> dict of [L] tensors."""
        src, tgt = mask_logits.flatten(1), target.flatten(1)
        cnt = torch.zeros(L, device=src.device).index_add_(0, level_of, torch.ones_like(level_of, dtype=torch.float32))
        cnt = cnt.clamp(min=1)
        focal = focal_sum(src, tgt).mean(1)
        p = src.sigmoid()
        dice = 1 - (2 * (p * tgt).sum(1) + 1) / (p.sum(1) + tgt.sum(1) + 1)
        z = torch.zeros(L, device=src.device, dtype=src.dtype)
        return {"loss_mask": z.index_add(0, level_of, focal) / cnt, "loss_dice": z.index_add(0, level_of, dice) / cnt}

    def reid_losses(self, ref_boxes, ref_prob, key_embed, ref_embed, best_q, ref_targets):
        """ref_boxes [B,nq,4] (last refined reference boxes of the ref frames), ref_prob [B,nq,K], embeddings [B,nq,C],
        best_q [B,G] (last level's best key query of every ground truth).  -> loss_reid, loss_reid_aux, n_items"""
        dev = ref_boxes.device
        B, nq = ref_boxes.shape[:2]
        gtb, gtl, ex = pad_targets(ref_targets, dev)
        G = gtb.shape[1]
        valid = torch.zeros(B, G, dtype=torch.bool, device=dev)
        for b, t in enumerate(ref_targets):
            valid[b, :len(t["valid"])] = t["valid"]
        with torch.no_grad():
            cost, iou = ota_cost(ref_prob, ref_boxes, gtb, gtl, valid)
            pos, bad1 = dynamic_k(cost, iou, valid, 10)
            neg, bad2 = dynamic_k(cost, iou, valid, 100)
            negmask = ~neg
            counts = torch.stack([pos.sum(1), negmask.sum(1), valid.long(), (bad1 | bad2)[:, None].expand(-1, G).long()])
            n_pos, n_neg, v, bad = counts.tolist()                               # the one sync of the reID branch
        if any(any(r) for r in bad):
            return None                                                          # caller falls back to the sequential code
        key = torch.gather(key_embed, 1, best_q[..., None].expand(-1, -1, key_embed.shape[-1]))      # [B,G,C]
        d = torch.einsum("bqc,bgc->bgq", ref_embed, key)                          # [B,G,nq]
        pm, nm = pos.transpose(1, 2), negmask.transpose(1, 2)                     # [B,G,nq]
        neg_fill = d.new_full((), NEG)
        lse_neg = torch.logsumexp(torch.where(nm, d, neg_fill), -1)
        lse_pos = torch.logsumexp(torch.where(pm, -d, neg_fill), -1)
        contrast = torch.where(valid, F.softplus(lse_neg + lse_pos), d.new_zeros(()))
        # auxiliary cosine loss on all positives + a random.sample of the negatives, drawn in the reference's order
        chosen = torch.zeros(B, G, nq, dtype=torch.bool)
        denom = torch.ones(B, G)
        n_items = 0
        for b in range(B):
            for g in range(G):
                if not v[b][g]:
                    continue
                n_items += 1
                k = 10 if n_pos[b][g] == 0 else (n_neg[b][g] if n_pos[b][g] * 10 >= n_neg[b][g] else n_pos[b][g] * 10)
                ids = random.sample(list(range(0, n_neg[b][g])), k)
                chosen[b, g, ids] = True                                          # by RANK among the negatives
                denom[b, g] = n_pos[b][g] + k
        if n_items == 0:
            return None
        chosen, denom = chosen.to(dev), denom.to(dev)
        rank = (nm.long().cumsum(-1) - 1).clamp(min=0)
        sampled = nm & chosen.gather(2, rank)
        cos = torch.einsum("bqc,bgc->bgq", F.normalize(ref_embed.float(), dim=-1), F.normalize(key.float(), dim=-1))
        aux = (torch.where(pm, (cos - 1) ** 2, cos.new_zeros(())) + torch.where(sampled, cos ** 2, cos.new_zeros(()))).sum(-1)
        aux = torch.where(valid, aux / denom, aux.new_zeros(()))
        return contrast.sum() / n_items, aux.sum() / n_items, n_items
