"""Clip-level instance tracking for long videos — SeqFormer's `CLIP_MATCHING` inference
(projects/SeqFormer/seqformer/seqformer.py:236-262, :310-361; models/clip_output.py `Videos` / `Clips`).

The reference keeps every clip's masks in dense `[num_clips, 120, video_length, H/4, W/4]` float tensors (two of them:
36 GB each for a 36-frame 720p video, which is why it parks them on the CPU, MERGE_ON_CPU).  Only two things are ever
read back: (1) the masks of the last `clip_length` clips on the frames they share with the incoming clip (soft IoU for
the Hungarian match) and (2) per (instance, frame) the SUM of logits over clips and the number of clips that covered it.
So the state here is the running sums `[instances, video_length, h, w]` (+ counts, class sums) and a window of the last
`clip_length` clips' probabilities — ~40x less memory, resident on the GPU, same numbers.  The assignment itself is
scipy's `linear_sum_assignment` on a tiny [instances x 10] matrix, as in the reference.
"""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment


def select_clip_instances(pred_logits, pred_masks, topk=10):
    """`SeqFormer.inference_clip` (seqformer.py:310-334): the clip's top-k queries by best class score.
    pred_logits [1,nq,K], pred_masks [1,nq,T,h,w] -> cls_probs [k,K], mask_logits [k,T,h,w]."""
    mask_cls = pred_logits[0].sigmoid()
    _, idx = torch.topk(mask_cls.max(1)[0], k=min(topk, mask_cls.shape[0]))
    return mask_cls[idx], pred_masks[0][idx]


class ClipTracks:
    """Same update / get_result semantics as models/clip_output.py:Videos (match threshold 0.01, ≤ 120 instances)."""

    def __init__(self, clip_length, video_length, num_classes, mask_size, device, max_inst=120, match_threshold=0.01):
        self.clip_length, self.video_length, self.device = clip_length, video_length, device
        self.match_threshold, self.max_inst = match_threshold, max_inst
        self.num_inst = 0
        self.num_clip = 0
        self.seen = set()
        self.logit_sum = torch.zeros(0, video_length, *mask_size, device=device)
        self.count = torch.zeros(0, video_length, device=device)
        self.cls_sum = torch.zeros(0, num_classes, device=device)
        self.cls_count = torch.zeros(0, device=device)
        self.window = []                      # last clip_length clips: (frame_idx, rows tensor, probs [n, T, h, w])

    def _grow(self, n):
        if n <= self.logit_sum.shape[0]:
            return
        if n > self.max_inst:
            raise RuntimeError(f"clip matching tracks more than {self.max_inst} instances (models/clip_output.py:26)")
        add = n - self.logit_sum.shape[0]
        pad = lambda t: torch.cat([t, t.new_zeros((add,) + tuple(t.shape[1:]))], 0)
        self.logit_sum, self.count, self.cls_sum, self.cls_count = map(pad, (self.logit_sum, self.count, self.cls_sum,
                                                                            self.cls_count))

    def _siou(self, probs, frame_idx):
        """Soft IoU [num_inst, n_in] of the incoming clip against the window on the shared frames
        (clip_output.py:37-66, called with the slices of :77-84)."""
        n_in = probs.shape[0]
        shared_in = [i for i, f in enumerate(frame_idx) if f in self.seen]
        shared_f = [frame_idx[i] for i in shared_in]
        recent = self.window[-len(frame_idx):] if frame_idx else []
        C, Ns, Ti = len(recent), self.num_inst, len(shared_f)
        hw = probs.shape[-2] * probs.shape[-1]
        saved = probs.new_zeros(C, Ns, Ti, hw)
        valid = torch.zeros(C, Ns, Ti, dtype=torch.bool, device=probs.device)
        for c, (frames, rows, p) in enumerate(recent):
            pos = {f: i for i, f in enumerate(frames)}
            cols = [(j, pos[f]) for j, f in enumerate(shared_f) if f in pos]
            if cols and len(rows):
                dst = torch.tensor([j for j, _ in cols], device=probs.device)
                src = torch.tensor([i for _, i in cols], device=probs.device)
                saved[c, rows[:, None], dst[None, :]] = p[:, src].flatten(-2)
                valid[c, rows[:, None], dst[None, :]] = True
        inp = probs[:, shared_in].flatten(-2)[None, None]                        # 1, 1, n_in, Ti, hw
        sv = saved.unsqueeze(2)                                                  # C, Ns, 1, Ti, hw
        vm = valid[:, :, None, :, None]
        num = (sv * inp * vm).sum(dim=(-1, -2))
        den = ((sv + inp - sv * inp) * vm).sum(dim=(-1, -2))
        siou = num / (den + 1e-6)                                                # C, Ns, n_in
        n_valid_clip = (vm.flatten(2).sum(dim=2) > 0).sum(dim=0)                 # Ns
        return siou.sum(dim=0) / (n_valid_clip[..., None] + 1e-6) if C else probs.new_zeros(Ns, n_in)

    def update(self, frame_idx, cls_probs, mask_logits):
        """One clip: frame_idx (list of video frame numbers), cls_probs [n,K], mask_logits [n,T,h,w]."""
        probs = mask_logits.sigmoid()
        n_in = probs.shape[0]
        scores = self._siou(probs, frame_idx)
        above = scores > self.match_threshold
        scores = scores * above.float()
        row_idx, col_idx = linear_sum_assignment(scores.cpu(), maximize=True)
        keep = above.cpu()[row_idx, col_idx].tolist()
        rows = [int(r) for r, k in zip(row_idx, keep) if k]
        cols = [int(c) for c, k in zip(col_idx, keep) if k]
        left = [i for i in range(n_in) if i not in cols]
        rows = rows + list(range(self.num_inst, self.num_inst + len(left)))
        cols = cols + left
        self._grow(self.num_inst + len(left))
        r = torch.tensor(rows, dtype=torch.long, device=self.device)
        c = torch.tensor(cols, dtype=torch.long, device=self.device)
        f = torch.tensor(frame_idx, dtype=torch.long, device=self.device)
        self.logit_sum[r[:, None], f[None, :]] += mask_logits[c]
        self.count[r[:, None], f[None, :]] += 1
        self.cls_sum[r] += cls_probs[c]
        self.cls_count[r] += 1
        self.window.append((list(frame_idx), r, probs[c]))
        self.window = self.window[-self.clip_length:]
        self.seen.update(frame_idx)
        self.num_clip += 1
        self.num_inst += len(left)

    def get_result(self):
        """-> class probabilities [N,K] (mean over the clips that saw the instance) and mask logits [N,L,h,w] (mean over
        covering clips; NaN where no clip covered that (instance, frame), exactly like the reference's 0/0)."""
        n = self.num_inst
        return self.cls_sum[:n] / self.cls_count[:n, None], self.logit_sum[:n] / self.count[:n, :, None, None]


@torch.no_grad()
def clip_matching_postprocess(pred_cls, pred_masks, ori_size, image_size, mask_stride=4, multi_cls=True,
                              apply_cls_thres=0.05, mask_format="bool"):
    """seqformer.py:336-361: class thresholding, x4 bilinear + sigmoid, crop, nearest resize, > 0.5, to the host (as boolean
    tensors, or with mask_format="rle" as per-frame COCO RLE dicts encoded on the device: vnext_b200/rle.py)."""
    if len(pred_cls) == 0:
        return {"image_size": ori_size, "pred_scores": [], "pred_labels": [], "pred_masks": []}
    if multi_cls:
        above = torch.where(pred_cls > apply_cls_thres)
        scores, labels = pred_cls[above], above[1]
        pred_masks = pred_masks[above[0]]
    else:
        scores, labels = pred_cls.max(-1)
    h, w = pred_masks.shape[-2:]
    m = F.interpolate(pred_masks, size=(h * mask_stride, w * mask_stride), mode="bilinear", align_corners=False).sigmoid()
    m = m[:, :, :image_size[0], :image_size[1]]
    m = F.interpolate(m, size=(ori_size[0], ori_size[1]), mode="nearest") > 0.5
    if mask_format == "rle":
        from ..rle import encode
        flat, t = encode(m), m.shape[1]
        masks = [flat[i * t:(i + 1) * t] for i in range(m.shape[0])]
    else:
        masks = [x for x in m.cpu()]
    return {"image_size": ori_size, "pred_scores": scores.tolist(), "pred_labels": labels.tolist(), "pred_masks": masks}


def clip_windows(video_length, clip_length, clip_stride):
    """The frame windows of seqformer.py:239-245 (the last one is right-aligned to the end of the video)."""
    out = []
    for start in range(0, video_length, clip_stride):
        end = start + clip_length
        last = end >= video_length
        if last:
            start, end = max(0, video_length - clip_length), video_length
        out.append(list(range(start, end)))
        if last:
            break
    return out
