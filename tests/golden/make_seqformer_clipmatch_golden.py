"""Golden of SeqFormer's CLIP_MATCHING inference from the REFERENCE's own code (build container only): the loop of
`SeqFormer.forward` (projects/SeqFormer/seqformer/seqformer.py:236-262) with `inference_clip` (:310-334),
`Videos` / `Clips` (models/clip_output.py) and `clip_matching_postprocess` (:336-361), fed with scripted per-clip
outputs (vnext_b200.synth.synthetic_clip_outputs) in place of `self.detr.inference`.  seqformer.py and clip_output.py
import detectron2 names at module level: stubbed (a 10-line `Instances` stand-in with attribute storage and `.to`)."""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_seqformer_golden as msg  # noqa: E402
from vnext_b200.synth import synthetic_clip_outputs  # noqa: E402

CASES = {"stride1": dict(video_length=9, clip_length=5, stride=1, seed=0),
         "stride3": dict(video_length=11, clip_length=5, stride=3, seed=1),
         "short": dict(video_length=4, clip_length=5, stride=1, seed=2)}
ORI, IMG = (40, 60), (44, 62)            # dataset output size, unpadded input size (mask grid 12x16 -> 48x64)


class Instances:
    def __init__(self, image_size):
        self.image_size = image_size

    def to(self, device):
        return self


def mount():
    msg.mount_reference()

    class _Reg:
        def register(self):
            return lambda c: c
    for name in ("detectron2", "detectron2.modeling", "detectron2.structures", "detectron2.utils", "detectron2.utils.memory"):
        sys.modules.setdefault(name, types.ModuleType(name))
    dm = sys.modules["detectron2.modeling"]
    dm.META_ARCH_REGISTRY, dm.build_backbone, dm.detector_postprocess = _Reg(), None, None
    ds = sys.modules["detectron2.structures"]
    ds.Instances = Instances
    ds.Boxes = ds.ImageList = ds.BitMasks = object
    sys.modules["detectron2.utils.memory"].retry_if_cuda_oom = lambda f: f
    data = types.ModuleType("refseq.data")
    data.__path__ = []
    dc = types.ModuleType("refseq.data.coco")
    dc.convert_coco_poly_to_mask = None
    sys.modules["refseq.data"], sys.modules["refseq.data.coco"] = data, dc
    return importlib.import_module("refseq.seqformer")


def run_reference(windows, video_length, clip_length, num_classes=40):
    mod = mount()
    me = types.SimpleNamespace(is_multi_cls=True, apply_cls_thres=0.05, mask_stride=4)
    video_output = None
    for frame_idx, output in windows:
        if video_output is None:
            interim = tuple(output["pred_masks"].shape[-2:])
            video_output = mod.Videos(clip_length, video_length, num_classes, interim, "cpu")
        res = mod.SeqFormer.inference_clip(me, output, IMG)
        video_output.update(mod.Clips(frame_idx, res.to("cpu")))
    pred_cls, pred_logits = video_output.get_result()
    return mod.SeqFormer.clip_matching_postprocess(me, pred_cls, pred_logits, ORI, IMG), pred_cls, pred_logits


def main():
    out = {}
    for name, kw in CASES.items():
        windows = synthetic_clip_outputs(**kw)
        res, cls, logits = run_reference(windows, kw["video_length"], kw["clip_length"])
        n = len(res["pred_scores"])
        out[f"{name}_scores"] = np.array(res["pred_scores"], dtype=np.float32)
        out[f"{name}_labels"] = np.array(res["pred_labels"], dtype=np.int64)
        out[f"{name}_masks"] = np.packbits(torch.stack(res["pred_masks"]).numpy().reshape(n, -1), axis=1) if n else np.zeros((0, 1), np.uint8)
        out[f"{name}_cls"] = cls.numpy()
        out[f"{name}_num_inst"] = np.array(cls.shape[0])
        print(name, "clips", len(windows), "tracked instances", cls.shape[0], "outputs", n, res["pred_labels"])
    path = os.path.join(HERE, "seqformer_clipmatch_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
