"""Size-independent properties of the two association stages (no golden needed): the result may not depend on WHICH
query slot a detection sits in, only on what it predicts — the online tracker (IDOL) and the clip stitcher (SeqFormer)
are run on larger scripted videos than the goldens and on slot-permuted copies of them."""
import torch

from vnext_b200.idol.tracker import IDOLTracker, video_inference
from vnext_b200.seqformer.clip_matching import ClipTracks, clip_matching_postprocess, select_clip_instances
from vnext_b200.synth import synthetic_clip_outputs, synthetic_track_outputs


def _tracker():
    return IDOLTracker(init_score_thr=0.2, obj_score_thr=0.1, nms_thr_pre=0.5, nms_thr_post=0.05, addnew_score_thr=0.2,
                       memo_tracklet_frames=10, memo_momentum=0.8, long_match=True, frame_weight=True,
                       temporal_weight=True, memory_len=3)


def _as_set(res):
    out = set()
    for s, l, masks in zip(res["pred_scores"], res["pred_labels"], res["pred_masks"]):
        if isinstance(masks, list):
            key = tuple(None if m is None else int(m.sum()) * 100003 + int(m.nonzero().sum()) for m in masks)
        else:
            key = (int(masks.sum()), int(masks.nonzero().sum()))
        out.add((round(s, 5), l, key))
    return out


def test_tracker_is_invariant_to_query_slot_permutation():
    out = synthetic_track_outputs(num_frames=30, num_queries=80, h=36, w=48, seed=11)
    base = video_inference(out, _tracker(), (90, 120), (140, 190))
    assert len(base["pred_scores"]) >= 8 and all(len(m) == 30 for m in base["pred_masks"])
    g = torch.Generator().manual_seed(0)
    perm = torch.stack([torch.randperm(80, generator=g) for _ in range(30)])            # a different shuffle per frame
    shuffled = {k: torch.stack([v[t][perm[t]] for t in range(30)]) for k, v in out.items()}
    again = video_inference(shuffled, _tracker(), (90, 120), (140, 190))
    assert _as_set(base) == _as_set(again)


def test_clip_stitching_is_invariant_to_query_slot_permutation():
    kw = dict(video_length=24, clip_length=5, stride=2, num_queries=60, h=24, w=32, seed=5)
    g = torch.Generator().manual_seed(1)

    def run(permute):
        tracks = None
        for frame_idx, o in synthetic_clip_outputs(**kw):
            logits, masks = o["pred_logits"], o["pred_masks"]
            if permute:
                p = torch.randperm(logits.shape[1], generator=g)
                logits, masks = logits[:, p], masks[:, p]
            if tracks is None:
                tracks = ClipTracks(5, 24, 40, tuple(masks.shape[-2:]), "cpu")
            tracks.update(frame_idx, *select_clip_instances(logits, masks))
        return clip_matching_postprocess(*tracks.get_result(), (60, 80), (90, 120))
    a, b = run(False), run(True)
    assert len(a["pred_scores"]) >= 6 and _as_set(a) == _as_set(b)
    # the three scripted objects are tracked through the whole video: their masks are non-empty on (almost) every frame
    full = [m for m in a["pred_masks"] if (m.flatten(1).sum(1) > 0).sum() >= 20]
    assert len(full) >= 4
