"""`instances_to_coco_json_video` with the reference's signature and output
(projects/SeqFormer/seqformer/data/ytvis_eval.py:171-211, projects/IDOL/idol/data/ytvis_eval.py:174-214): the per-video list of
{"video_id", "score", "category_id", "segmentations": [RLE dict per frame]} that YTVISEvaluator.process collects and dumps to
results.json.  The reference RLE-encodes every boolean mask on the host with pycocotools; here the masks arrive already
encoded by the device (SeqFormer(..., mask_format="rle")), or — boolean tensors, the reference's format — are encoded on
the device now (vnext_b200/rle.py).  A frame whose mask is None gets the all-zero mask's RLE, like the reference.
"""
import torch

from .. import rle


def _empty_rle(h, w):
    # all-zero mask: one run of h*w zeros -> rleToString of a single count (csrc/rle_encode.cu writes the same characters)
    x, out = h * w, []
    more = True
    while more:
        c = x & 0x1F
        x >>= 5
        more = (x != -1) if (c & 0x10) else (x != 0)
        out.append(chr((c | 0x20 if more else c) + 48))
    return {"size": [h, w], "counts": "".join(out)}


def instances_to_coco_json_video(inputs, outputs, device="cuda"):
    assert len(inputs) == 1, "More than one inputs are loaded for inference!"
    video_id = inputs[0]["video_id"]
    h, w = inputs[0]["height"], inputs[0]["width"]
    results = []
    for s, l, m in zip(outputs["pred_scores"], outputs["pred_labels"], outputs["pred_masks"]):
        if torch.is_tensor(m):                        # [T, H, W] boolean masks: encode on the device
            segms = rle.encode(m.to(device))
        else:
            frames = [f for f in m if f is not None and torch.is_tensor(f)]
            enc = iter(rle.encode(torch.stack(frames).to(device))) if frames else iter(())
            segms = [_empty_rle(h, w) if f is None else (next(enc) if torch.is_tensor(f) else f) for f in m]
        results.append({"video_id": video_id, "score": s, "category_id": l, "segmentations": segms})
    return results
