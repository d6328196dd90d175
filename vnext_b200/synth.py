"""Synthetic inputs and weights (no datasets or checkpoints exist offline): used by bench.py, the
tests, the golden-vector scripts and the oracle alike.  Nothing here computes any part of the hot path.

Deterministic synthetic weights keyed by state-dict name:

There are no checkpoints offline, and a seeded `reset_parameters()` depends on module construction
order (which differs between the reference's module tree and this repo's engine).  So every tensor of
a state dict is filled from a generator seeded by crc32(key): the reference model (in
tests/golden/make_*_golden.py), the CPU oracle and the CUDA engine all get bit-identical weights from
the key names alone.  Distributions keep activations O(1) through 50+ layers and — unlike the
reference's init, which zeroes `sampling_offsets.weight` / `attention_weights.*` — make every
parameter matter, so parity tests exercise data-dependent sampling locations.
"""
import zlib

import torch


def _gen(key, seed):
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))


def fill_tensor(key, shape, seed=0, dtype=torch.float32):
    g = _gen(key, seed)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]

    def randn(std=1.0, mean=0.0):
        return (torch.randn(shape, generator=g) * std + mean).to(dtype)

    def uniform(lo, hi):
        return (torch.rand(shape, generator=g) * (hi - lo) + lo).to(dtype)

    if leaf == "relative_position_index" or leaf == "num_batches_tracked":
        raise KeyError(key)                     # integer buffers are structural, never synthesised
    if leaf == "running_var":
        return uniform(0.5, 1.5)
    if leaf == "running_mean":
        return randn(0.1)
    is_norm = (".norm" in key or "norm." in key or key.split(".")[-2].startswith("norm")
               or (".input_proj." in key and key.split(".")[-2] == "1"))
    if len(shape) == 1:
        if leaf == "weight":                     # affine scale of BN / GN / LN
            return uniform(0.8, 1.2)
        if "sampling_offsets" in key:            # pixel-scale offsets so samples spread over cells
            return randn(1.5)
        if "class_embed" in key:                 # prior-probability bias of the reference init (-4.6), jittered
            return randn(0.5, -4.0)
        if is_norm:
            return randn(0.05)
        return randn(0.05)
    if "query_embed" in key or "level_embed" in key or "relative_position_bias_table" in key:
        return randn(1.0 if "relative_position_bias_table" not in key else 0.2)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    std = 1.0 / (fan_in ** 0.5)
    if "sampling_offsets" in key:
        std *= 0.5
    if "bbox_embed" in key and key.split(".")[-2] == "2":
        std *= 0.3
    if "controller" in key and key.split(".")[-2] == "2":
        std *= 0.5
    if "backbone" in key and len(shape) == 4:    # ReLU networks: He scaling keeps the signal alive
        std *= 2 ** 0.5
        if key.endswith("conv3.weight"):
            std *= 0.5                            # residual branches add up; keep res5 bounded
    return randn(std)


def fill_state_dict(spec, seed=0):
    """spec: dict key -> shape (or tensor).  Integer / structural buffers are passed through."""
    out = {}
    for k, v in spec.items():
        if isinstance(v, torch.Tensor) and not v.dtype.is_floating_point:
            out[k] = v.clone()
            continue
        if isinstance(v, tuple) and len(v) == 2 and v[0] == "index":       # Swin relative_position_index (structural)
            out[k] = swin_relative_position_index(v[1])
            continue
        shape = v.shape if isinstance(v, torch.Tensor) else v
        out[k] = fill_tensor(k, shape, seed)
    return out


def synthetic_clip(num_frames, height, width, seed=0):
    """uint8-valued float frames, the bench's `data: synthetic` (SURVEY §8d cfg 2)."""
    g = torch.Generator().manual_seed(1234 + seed)
    return [torch.randint(0, 256, (3, height, width), generator=g).float() for _ in range(num_frames)]


def synthetic_targets(num_videos, height, width, seed=0, insts=3, num_classes=40):
    """Ground truth of `num_videos` (key, ref) frame pairs in the format IDOL.prepare_targets produces
    (projects/IDOL/idol/idol.py:296-324): per frame {"labels","boxes" (cx,cy,w,h normalised),"masks" [n,H,W] bool,
    "inst_id","valid"}.  Instances are axis-aligned rectangles that drift between key and ref; in every odd video the
    last instance has left the reference frame (inst_id -1, valid False, empty box and mask), the case
    pos_neg_select.py:28-36 handles.  -> (det_targets, ref_targets)"""
    g = _gen("targets", seed)
    det, ref = [], []
    for v in range(num_videos):
        labels = torch.randint(0, num_classes, (insts,), generator=g)
        cx = (torch.rand(insts, generator=g) * 0.6 + 0.2)
        cy = (torch.rand(insts, generator=g) * 0.6 + 0.2)
        bw = (torch.rand(insts, generator=g) * 0.25 + 0.1)
        bh = (torch.rand(insts, generator=g) * 0.25 + 0.1)
        shift = (torch.rand(insts, 2, generator=g) - 0.5) * 0.1
        frames = []
        for f in range(2):
            x0 = ((cx + f * shift[:, 0] - bw / 2).clamp(0, 1) * width).round()
            x1 = ((cx + f * shift[:, 0] + bw / 2).clamp(0, 1) * width).round()
            y0 = ((cy + f * shift[:, 1] - bh / 2).clamp(0, 1) * height).round()
            y1 = ((cy + f * shift[:, 1] + bh / 2).clamp(0, 1) * height).round()
            ids = torch.arange(insts) + 10 * v
            valid = torch.ones(insts, dtype=torch.bool)
            if f == 1 and v % 2 == 1:
                x0[-1] = x1[-1] = y0[-1] = y1[-1] = 0
                ids[-1] = -1
                valid[-1] = False
            ys = torch.arange(height).view(1, height, 1)
            xs = torch.arange(width).view(1, 1, width)
            masks = (xs >= x0.view(-1, 1, 1)) & (xs < x1.view(-1, 1, 1)) & (ys >= y0.view(-1, 1, 1)) & \
                    (ys < y1.view(-1, 1, 1))
            boxes = torch.stack([(x0 + x1) / 2 / width, (y0 + y1) / 2 / height, (x1 - x0) / width,
                                 (y1 - y0) / height], -1)
            frames.append({"labels": labels.clone(), "boxes": boxes, "masks": masks, "inst_id": ids, "valid": valid})
        det.append(frames[0])
        ref.append(frames[1])
    return det, ref


def synthetic_track_outputs(num_frames=14, num_queries=40, num_classes=40, embed_dim=32, h=24, w=32, seed=0):
    """Dense per-frame outputs (`detr.inference_forward` format, projects/IDOL/idol/models/segmentation_condInst.py:
    233-321) of a scripted video for exercising the online tracker: 4 objects drifting across the frame, one of them
    hidden for three frames, one entering late, duplicate detections (mask-NMS fodder), low-score clutter, and one
    frame where nothing passes INFERENCE_SELECT_THRES (the top-1 fallback, idol.py:355-358).  Detections land in a
    different query slot every frame."""
    g = _gen("track", seed)
    T, nq, K, C = num_frames, num_queries, num_classes, embed_dim
    logits = torch.full((T, nq, K), -6.0) + torch.randn(T, nq, K, generator=g) * 0.2
    boxes = torch.rand(T, nq, 4, generator=g) * 0.1 + torch.tensor([0.45, 0.45, 0.02, 0.02])
    embeds = torch.randn(T, nq, C, generator=g) * 0.3
    masks = torch.full((T, nq, 1, h, w), -4.0) + torch.randn(T, nq, 1, h, w, generator=g) * 0.3
    n_obj = 4
    ident = torch.nn.functional.normalize(torch.randn(n_obj, C, generator=g), dim=1) * 4.0
    cls = torch.randint(0, K, (n_obj,), generator=g)
    start = torch.rand(n_obj, 2, generator=g) * 0.5 + 0.2
    vel = (torch.rand(n_obj, 2, generator=g) - 0.5) * 0.03
    size = torch.rand(n_obj, 2, generator=g) * 0.15 + 0.15
    ys, xs = torch.arange(h).view(h, 1), torch.arange(w).view(1, w)

    def paint(t, q, c, wh, score, k, noise=0.15):
        logits[t, q, cls[k]] = torch.logit(torch.tensor(score))
        logits[t, q, (cls[k] + 1) % K] = torch.logit(torch.tensor(score * 0.3))
        boxes[t, q] = torch.cat([c, wh])
        embeds[t, q] = ident[k] + torch.randn(C, generator=g) * noise
        x0, x1 = (c[0] - wh[0] / 2) * w, (c[0] + wh[0] / 2) * w
        y0, y1 = (c[1] - wh[1] / 2) * h, (c[1] + wh[1] / 2) * h
        inside = (xs >= x0) & (xs < x1) & (ys >= y0) & (ys < y1)
        masks[t, q, 0] = torch.where(inside, torch.tensor(4.0), torch.tensor(-4.0)) + torch.randn(h, w, generator=g) * 0.3

    for t in range(T):
        if t == 11:
            continue                                       # nothing confident in this frame
        slots = torch.randperm(nq, generator=g).tolist()
        for k in range(n_obj):
            if (k == 2 and 5 <= t <= 7) or (k == 3 and t < 4):
                continue
            c = (start[k] + vel[k] * t).clamp(0.1, 0.9)
            score = float(0.45 + 0.4 * torch.rand((), generator=g))
            paint(t, slots.pop(), c, size[k], score, k)
            if (t + k) % 3 == 0:                           # a duplicate of the same object, slightly off, lower score
                paint(t, slots.pop(), c + 0.01, size[k] * 0.95, score * 0.6, k, noise=0.4)
        for _ in range(2):                                 # weak clutter just above the selection threshold
            q = slots.pop()
            logits[t, q, int(torch.randint(0, K, (), generator=g))] = torch.logit(torch.tensor(0.13))
    return {"pred_logits": logits, "pred_boxes": boxes, "pred_inst_embed": embeds, "pred_masks": masks}


def synthetic_clip_outputs(video_length=9, clip_length=5, stride=1, num_queries=30, num_classes=40, h=12, w=16, seed=0):
    """Per-clip dense outputs of SeqFormer (`CondInst_segm.inference`: pred_logits [1,nq,K], pred_masks [1,nq,T,h,w])
    for every window that CLIP_MATCHING inference visits (projects/SeqFormer/seqformer/seqformer.py:236-262), from a
    scripted video with 3 moving objects (one enters late) + clutter (background logits -9, so every soft IoU is either
    large or orders of magnitude below the 0.01 match threshold: no borderline decisions for float noise to flip).  Objects land in different query slots in every
    clip, so the clips can only be stitched by mask overlap.  -> list of (frame_idx list, outputs dict)."""
    g = _gen("clips", seed)
    n_obj, K = 3, num_classes
    cls = torch.randint(0, K, (n_obj,), generator=g)
    start = torch.rand(n_obj, 2, generator=g) * 0.5 + 0.2
    vel = (torch.rand(n_obj, 2, generator=g) - 0.5) * 0.06
    size = torch.rand(n_obj, 2, generator=g) * 0.2 + 0.2
    ys, xs = torch.arange(h).view(h, 1), torch.arange(w).view(1, w)
    windows, last = [], False
    for s0 in range(0, video_length, stride):
        e0 = s0 + clip_length
        if e0 >= video_length:
            last = True
            s0, e0 = max(0, video_length - clip_length), video_length
        windows.append(list(range(s0, e0)))
        if last:
            break
    out = []
    for frames in windows:
        T = len(frames)
        logits = torch.full((1, num_queries, K), -6.0) + torch.randn(1, num_queries, K, generator=g) * 0.2
        masks = torch.full((1, num_queries, T, h, w), -9.0) + torch.randn(1, num_queries, T, h, w, generator=g) * 0.3
        slots = torch.randperm(num_queries, generator=g).tolist()
        for k in range(n_obj):
            if k == 2 and frames[-1] < 3:
                continue                                   # third object enters at frame 3
            q = slots.pop()
            score = float(0.5 + 0.4 * torch.rand((), generator=g))
            logits[0, q, cls[k]] = torch.logit(torch.tensor(score))
            logits[0, q, (cls[k] + 3) % K] = torch.logit(torch.tensor(score * 0.2))
            for ti, f in enumerate(frames):
                if k == 2 and f < 3:
                    continue
                c = (start[k] + vel[k] * f).clamp(0.1, 0.9)
                inside = (xs >= (c[0] - size[k, 0] / 2) * w) & (xs < (c[0] + size[k, 0] / 2) * w) & \
                         (ys >= (c[1] - size[k, 1] / 2) * h) & (ys < (c[1] + size[k, 1] / 2) * h)
                masks[0, q, ti] = torch.where(inside, torch.tensor(5.0), torch.tensor(-9.0)) + \
                    torch.randn(h, w, generator=g) * 0.3
        for j in range(3):                                 # clutter that still makes the top-10 (distinct scores: no
            score = 0.08 - 0.005 * j                       # top-k ties, whose order is implementation-defined)
            logits[0, slots.pop(), int(torch.randint(0, K, (), generator=g))] = torch.logit(torch.tensor(score))
        out.append((frames, {"pred_logits": logits, "pred_masks": masks}))
    return out


def synthetic_clip_targets(num_videos, num_frames, height, width, seed=0, insts=3, num_classes=40):
    """Clip-level ground truth in the format SeqFormer.prepare_targets produces (projects/SeqFormer/seqformer/
    seqformer.py:266-305): per video {"labels" [n], "boxes" [n, nf, 4] (cx,cy,w,h normalised), "masks" [n, nf, H, W]
    bool, "size" tensor([H, W])}: rectangles drifting over the frames."""
    g = _gen("clip_targets", seed)
    out = []
    for v in range(num_videos):
        labels = torch.randint(0, num_classes, (insts,), generator=g)
        cx = torch.rand(insts, generator=g) * 0.5 + 0.25
        cy = torch.rand(insts, generator=g) * 0.5 + 0.25
        bw = torch.rand(insts, generator=g) * 0.25 + 0.1
        bh = torch.rand(insts, generator=g) * 0.25 + 0.1
        shift = (torch.rand(insts, 2, generator=g) - 0.5) * 0.08
        boxes, masks = [], []
        ys = torch.arange(height).view(1, height, 1)
        xs = torch.arange(width).view(1, 1, width)
        for f in range(num_frames):
            x0 = ((cx + f * shift[:, 0] - bw / 2).clamp(0, 1) * width).round()
            x1 = ((cx + f * shift[:, 0] + bw / 2).clamp(0, 1) * width).round()
            y0 = ((cy + f * shift[:, 1] - bh / 2).clamp(0, 1) * height).round()
            y1 = ((cy + f * shift[:, 1] + bh / 2).clamp(0, 1) * height).round()
            masks.append((xs >= x0.view(-1, 1, 1)) & (xs < x1.view(-1, 1, 1)) & (ys >= y0.view(-1, 1, 1)) &
                         (ys < y1.view(-1, 1, 1)))
            boxes.append(torch.stack([(x0 + x1) / 2 / width, (y0 + y1) / 2 / height, (x1 - x0) / width,
                                      (y1 - y0) / height], -1))
        out.append({"labels": labels, "boxes": torch.stack(boxes, 1), "masks": torch.stack(masks, 1),
                    "size": torch.as_tensor([height, width], dtype=torch.long)})
    return out


def make_msda_inputs(N, Lq, M, D, shapes, P, seed, dtype=torch.float32, loc_range=(0.0, 1.0)):
    """MSDA input recipe of the reference test (projects/IDOL/idol/models/ops/test.py:33-36):
    value ~ U(0, .01), sampling locations ~ U(loc_range), attention weights normalised over (L, P)."""
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes_t.shape[0]
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    S = int(shapes_t.prod(1).sum())
    value = torch.rand(N, S, M, D, generator=g) * 0.01
    lo, hi = loc_range
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * (hi - lo) + lo
    attn = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    attn = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value.to(dtype), shapes_t, lsi, loc.to(dtype), attn.to(dtype)


def swin_relative_position_index(ws):
    """The integer buffer WindowAttention registers (projects/IDOL/idol/backbone/swin.py:108-117)."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij"))
    flat = torch.flatten(coords, 1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)
