"""IDOL's online association and video post-processing, restated for the GPU box:
`IDOL_Tracker` (projects/IDOL/idol/models/tracker.py:50-298) and `IDOL.inference` (projects/IDOL/idol/idol.py:326-473).

The reference walks the frames with everything on the model's device and decides with Python `if`s on device scalars:
the mask-NMS double loop alone (tracker.py:26-46) is O(n^2) kernel launches + host syncs per frame, and every kept
mask is upsampled and copied to the host one by one (idol.py:438-443).  Here the dense, data-parallel parts run as a
few batched device ops for the WHOLE video up front —
  * per-frame candidate selection: sigmoid / max / threshold / class-aware box NMS (idol.py:349-366);
  * binary masks and ALL pairwise mask IoUs of a frame's candidates as one GEMM on 0/1 matrices (exact integers);
  * the final x4 bilinear upsample + sigmoid + crop + nearest resize + threshold of every (instance, frame) mask in
    one batch and ONE device→host copy —
and the inherently sequential bookkeeping (greedy bi-softmax matching against the memory bank, tracklet momentum
updates, id allocation) runs on the host on small CPU tensors, one copy of the frame's candidates per frame.  Decisions
are the reference's, statement for statement; quirks kept are marked `# ref quirk`.
"""
import torch
import torch.nn.functional as F
import torchvision.ops as tvops


def _xyxy(b):
    cx, cy, w, h = b.unbind(1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], 1)


def pairwise_mask_iou(binary):
    """binary [n, h*w] bool -> [n, n] float32 IoU with the reference's smoothing (tracker.py:17-24)."""
    f = binary.float()
    inter = f @ f.t()                                    # exact: counts << 2^24
    area = f.sum(1)
    union = area[:, None] + area[None, :] - inter
    return (inter + 1e-6) / (union + 1e-6)


def greedy_mask_nms(iou, thr):
    """tracker.py:26-46 on a precomputed IoU matrix (host)."""
    n = iou.shape[0]
    keep = [True] * n
    rows = iou.tolist()
    for i in range(n - 1):
        if not keep[i]:
            continue
        for j in range(i + 1, n):
            if keep[j] and rows[i][j] > thr:
                keep[j] = False
    return keep


class IDOLTracker:
    """Same constructor arguments and semantics as `IDOL_Tracker` (tracker.py:52-97); state lives on the host."""

    def __init__(self, nms_thr_pre=0.7, nms_thr_post=0.3, init_score_thr=0.2, addnew_score_thr=0.5, obj_score_thr=0.1,
                 match_score_thr=0.5, memo_tracklet_frames=10, memo_backdrop_frames=1, memo_momentum=0.5,
                 nms_conf_thr=0.5, nms_backdrop_iou_thr=0.5, nms_class_iou_thr=0.7, with_cats=True,
                 match_metric="bisoftmax", long_match=False, frame_weight=False, temporal_weight=False, memory_len=10):
        assert 0 <= memo_momentum <= 1.0 and memo_tracklet_frames >= 0 and memo_backdrop_frames >= 0
        assert match_metric in ("bisoftmax", "softmax", "cosine")
        self.memory_len, self.temporal_weight, self.long_match, self.frame_weight = \
            memory_len, temporal_weight, long_match, frame_weight
        self.nms_thr_pre, self.nms_thr_post = nms_thr_pre, nms_thr_post
        self.init_score_thr, self.addnew_score_thr, self.match_score_thr = init_score_thr, addnew_score_thr, match_score_thr
        self.memo_tracklet_frames, self.memo_backdrop_frames, self.memo_momentum = \
            memo_tracklet_frames, memo_backdrop_frames, memo_momentum
        self.match_metric = match_metric
        self.num_tracklets = 0
        self.tracklets = dict()
        self.backdrops = []

    @property
    def empty(self):
        return not self.tracklets

    # ------------------------------------------------------------------ memory bank (tracker.py:102-205)
    def update_memo(self, ids, bboxes, embeds, labels, frame_id):
        for i in torch.nonzero(ids > -1).flatten().tolist():
            tid, bbox, embed, label = int(ids[i]), bboxes[i], embeds[i], labels[i]
            t = self.tracklets.get(tid)
            if t is not None:
                velocity = (bbox - t["bbox"]) / (frame_id - t["last_frame"])
                t["bbox"] = bbox
                t["long_score"].append(bbox[-1])
                t["embed"] = (1 - self.memo_momentum) * t["embed"] + self.memo_momentum * embed
                t["long_embed"].append(embed)
                t["last_frame"] = frame_id
                t["label"] = label
                t["velocity"] = (t["velocity"] * t["acc_frame"] + velocity) / (t["acc_frame"] + 1)
                t["acc_frame"] += 1
                t["exist_frame"] += 1
            else:
                self.tracklets[tid] = dict(bbox=bbox, embed=embed, long_embed=[embed], long_score=[bbox[-1]], label=label,
                                           last_frame=frame_id, velocity=torch.zeros_like(bbox), acc_frame=0,
                                           exist_frame=1)
        back = torch.nonzero(ids == -1).flatten()
        self.backdrops.insert(0, dict(bboxes=bboxes[back], embeds=embeds[back], labels=labels[back]))
        dead = []
        for k, v in self.tracklets.items():
            if frame_id - v["last_frame"] >= self.memo_tracklet_frames:
                dead.append(k)
            if len(v["long_embed"]) > self.memory_len:
                v["long_embed"].pop(0)
            if len(v["long_score"]) > self.memory_len:
                v["long_score"].pop(0)
        for k in dead:
            self.tracklets.pop(k)
        if len(self.backdrops) > self.memo_backdrop_frames:
            self.backdrops.pop()

    def memo(self):
        embeds, ids, exist = [], [], []
        for k, v in self.tracklets.items():
            if self.long_match:
                weights = torch.stack(v["long_score"])
                if self.temporal_weight:
                    n = len(weights)
                    weights = weights + torch.arange(0.0, 1 + 1 / n, 1 / n)[1:n + 1].to(weights)   # torch.range(0,1,1/n)[1:]
                embeds.append((torch.stack(v["long_embed"]) * weights.unsqueeze(1)).sum(0) / weights.sum())
            else:
                embeds.append(v["embed"])
            ids.append(k)
            exist.append(v["exist_frame"])
        return torch.stack(embeds), torch.tensor(ids, dtype=torch.long), torch.tensor(exist, dtype=torch.long)

    # ------------------------------------------------------------------ one frame (tracker.py:207-298)
    def match(self, bboxes, labels, mask_iou, track_feats, frame_id, indices):
        """bboxes [n,5] (cx,cy,w,h,score), labels [n], mask_iou [n,n] (pairwise IoU of the candidates' binary masks),
        track_feats [n,C] — all host tensors, candidates in descending score order; indices: their query slots.
        -> (bboxes, labels, ids, indices) of the candidates that survive mask NMS, like the reference."""
        valids = greedy_mask_nms(mask_iou, self.nms_thr_pre)
        keep = torch.tensor(valids, dtype=torch.bool)
        indices = torch.tensor(indices)[keep].tolist()
        bboxes, labels, embeds = bboxes[keep], labels[keep], track_feats[keep]
        iou = mask_iou[keep][:, keep]
        n = bboxes.size(0)
        ids = torch.full((n,), -2, dtype=torch.long)
        if n > 0 and not self.empty:
            memo_embeds, memo_ids, memo_exist = self.memo()
            memo_exist = memo_exist.to(memo_embeds)
            memo_ids_f = memo_ids.to(memo_embeds)
            if self.match_metric == "bisoftmax":
                feats = embeds @ memo_embeds.t()
                scores = (feats.softmax(dim=1) + feats.softmax(dim=0)) / 2
            elif self.match_metric == "softmax":
                scores = (embeds @ memo_embeds.t()).softmax(dim=1)
            else:
                scores = F.normalize(embeds, p=2, dim=1) @ F.normalize(memo_embeds, p=2, dim=1).t()
            for i in range(n):
                row = scores[i, :]
                if self.frame_weight:
                    non_backs = (memo_ids_f > -1) & (row > 0.5)
                    if (row[non_backs] > 0.5).sum() > 1:
                        weighted = row.clone()
                        fw = memo_exist[row[memo_ids_f > -1] > 0.5]
                        weighted[non_backs] = weighted[non_backs] * fw
                        weighted[~non_backs] = weighted[~non_backs] * fw.mean()
                        conf, memo_ind = torch.max(weighted, dim=0)
                    else:
                        conf, memo_ind = torch.max(row, dim=0)
                else:
                    conf, memo_ind = torch.max(row, dim=0)
                tid = memo_ids[memo_ind]
                if conf > self.match_score_thr and tid > -1:
                    ids[i] = tid
                    scores[:i, memo_ind] = 0
                    scores[i + 1:, memo_ind] = 0
            new = (ids == -2) & (bboxes[:, 4] > self.addnew_score_thr)
        elif self.empty:
            new = (ids == -2) & (bboxes[:, 4] > self.init_score_thr)
        else:
            return bboxes, labels, ids, indices          # ref quirk: no candidates and a non-empty bank: nothing happens
        num_news = int(new.sum())
        ids[new] = torch.arange(self.num_tracklets, self.num_tracklets + num_news, dtype=torch.long)
        self.num_tracklets += num_news
        for ind in torch.nonzero(ids == -2).flatten().tolist():
            if bool((iou[ind, :ind] < self.nms_thr_post).all()):                # not overlapping any better candidate
                ids[ind] = -1                                                    # -> backdrop
        self.update_memo(ids, bboxes, embeds, labels, frame_id)
        return bboxes, labels, ids, indices


@torch.no_grad()
def video_inference(outputs, tracker, ori_size, image_size, select_thres=0.1, temporal_score_type="mean",
                    multi_cls=True, apply_cls_thres=0.05, mask_stride=4, mask_format="bool"):
    """`IDOL.inference(outputs, tracker, ori_size, image_sizes)` (idol.py:326-473).  outputs: the dense per-frame
    predictions (pred_logits [T,nq,K], pred_boxes [T,nq,4], pred_inst_embed [T,nq,C], pred_masks [T,nq,1,h,w]) on any
    device.  -> {"image_size", "pred_scores", "pred_labels", "pred_masks": per instance a list over frames of bool
    [H,W] CPU tensors or None}.  mask_format="rle" (CUDA outputs only): the frames' masks as COCO RLE dicts instead, encoded
    on the device (vnext_b200/rle.py) — what ytvis_eval.py:196-204 makes of them next; None frames stay None."""
    logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
    embeds, masks = outputs["pred_inst_embed"], outputs["pred_masks"]
    T, nq = logits.shape[:2]
    h, w = masks.shape[-2:]
    prob = logits.sigmoid()
    max_score, max_cls = prob.max(2)                                            # [T,nq]
    prob_host = prob.cpu()
    video = {}                                                                  # id -> {"slots": [q or None]*T, "valid"}
    for t in range(T):
        idx = torch.nonzero(max_score[t] > select_thres).flatten()
        if len(idx) == 0:
            idx = prob_host[t].max(1)[0].argmax().reshape(1).to(logits.device)  # top-1 fallback (idol.py:355-358)
        else:
            keep = tvops.batched_nms(_xyxy(boxes[t][idx]), max_score[t][idx], max_cls[t][idx], 0.9)
            idx = idx[keep]                                                     # descending score
        cand = torch.cat([boxes[t][idx], max_score[t][idx, None], max_cls[t][idx, None].float(), embeds[t][idx],
                          pairwise_mask_iou((masks[t][idx].sigmoid() > 0.5).flatten(1))], 1).cpu()   # ONE copy per frame
        # (binary masks of the frame's CANDIDATES only: the whole video's [T, nq, h, w] sigmoid is never materialised)
        n = len(idx)
        det_bboxes, det_labels = cand[:, :5], cand[:, 5].long()
        C = embeds.shape[-1]
        feats, iou = cand[:, 6:6 + C], cand[:, 6 + C:6 + C + n]
        _, _, ids, slots = tracker.match(det_bboxes, det_labels, iou, feats, t, idx.tolist())
        for q, tid in zip(torch.tensor(slots)[ids > -1].tolist(), ids[ids > -1].tolist()):
            v = video.get(tid)
            if v is None:
                v = video[tid] = {"slots": [None] * t, "valid": 0}
            v["slots"].append(q)
            v["valid"] += 1
        for v in video.values():
            if len(v["slots"]) < t + 1:
                v["slots"].append(None)
        if t > 8:                                   # drop short noisy tracks (idol.py:404-411)
            for k in [k for k, v in video.items() if v["slots"][-1] is None and v["slots"][-2] is None and v["valid"] < 3]:
                video.pop(k)

    # every kept (instance, frame) mask in one batch: x4 bilinear -> sigmoid -> crop -> nearest resize -> > 0.5
    pairs = [(i, t, q) for i, v in enumerate(video.values()) for t, q in enumerate(v["slots"]) if q is not None]
    per_inst = [[None] * T for _ in video]
    if pairs:
        tt = torch.tensor([p[1] for p in pairs], device=masks.device)
        qq = torch.tensor([p[2] for p in pairs], device=masks.device)
        m = F.interpolate(masks[tt, qq], size=(h * mask_stride, w * mask_stride), mode="bilinear", align_corners=False)
        m = m.sigmoid()[:, :, :image_size[0], :image_size[1]]
        m = (F.interpolate(m, size=(ori_size[0], ori_size[1]), mode="nearest") > 0.5)[:, 0]
        if mask_format == "rle":
            from ..rle import encode
            m = encode(m.contiguous())                    # strings instead of H*W bytes per mask
        else:
            m = m.cpu()                                   # ONE copy
        for k, (i, t, _) in enumerate(pairs):
            per_inst[i][t] = m[k]
    cls = []
    for v in video.values():
        s = torch.stack([prob_host[t][q] for t, q in enumerate(v["slots"]) if q is not None])
        if temporal_score_type == "mean":
            cls.append(s.mean(0))
        elif temporal_score_type == "max":
            cls.append(s.max(0)[0])
        else:
            raise ValueError(f"non valid temporal_score_type {temporal_score_type!r}")
    if cls:
        pred_cls = torch.stack(cls)
        if multi_cls:
            above = torch.where(pred_cls > apply_cls_thres)
            scores, labels = pred_cls[above], above[1]
            out_masks = [per_inst[i] for i in above[0].tolist()]
        else:
            scores, labels = pred_cls.max(-1)
            out_masks = per_inst
        scores, labels = scores.tolist(), labels.tolist()
    else:
        scores, labels, out_masks = [], [], []
    return {"image_size": ori_size, "pred_scores": scores, "pred_labels": labels, "pred_masks": out_masks}
