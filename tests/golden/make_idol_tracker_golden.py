"""Golden of IDOL's online association + video post-processing from the REFERENCE's own code (build container only):
`IDOL.inference` (projects/IDOL/idol/idol.py:326-473) driving `IDOL_Tracker` (models/tracker.py) with the defaults of
idol/config.py:51-65 and the tracker arguments of idol.py:276-290, on a scripted video
(vnext_b200.synth.synthetic_track_outputs).  idol.py imports detectron2 / fvcore / pycocotools names at module level;
they are stubbed (none is touched by `inference`)."""
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
import make_idol_golden as mig  # noqa: E402
from vnext_b200.synth import synthetic_track_outputs  # noqa: E402

CASES = {"default": dict(seed=0), "seed1": dict(seed=1), "nofw": dict(seed=2, inference_fw=False, inference_tw=False)}
ORI, IMG = (72, 100), (90, 120)          # output size of the dataset, unpadded model input size (mask grid 24x32 -> 96x128)


def mount_idol():
    mig.mount()

    class _Reg:
        def register(self):
            return lambda c: c
    dm = types.ModuleType("detectron2.modeling")
    dm.META_ARCH_REGISTRY, dm.build_backbone, dm.detector_postprocess = _Reg(), None, None
    ds = sys.modules["detectron2.structures"]
    ds.Boxes = ds.ImageList = ds.BitMasks = object
    sys.modules["detectron2.modeling"] = dm
    sys.modules["fvcore.nn"].giou_loss = sys.modules["fvcore.nn"].smooth_l1_loss = None
    dc = types.ModuleType("refidol.data.coco")
    dc.convert_coco_poly_to_mask = None
    data = types.ModuleType("refidol.data")
    data.__path__ = []
    sys.modules["refidol.data"], sys.modules["refidol.data.coco"] = data, dc
    return importlib.import_module("refidol.idol")


def run_reference(outputs, inference_fw=True, inference_tw=True):
    mod = mount_idol()
    from refidol.models.tracker import IDOL_Tracker
    me = types.SimpleNamespace(inference_select_thres=0.1, temporal_score_type="mean", is_multi_cls=True,
                               apply_cls_thres=0.05)
    tracker = IDOL_Tracker(init_score_thr=0.2, obj_score_thr=0.1, nms_thr_pre=0.5, nms_thr_post=0.05,
                           addnew_score_thr=0.2, memo_tracklet_frames=10, memo_momentum=0.8, long_match=inference_tw,
                           frame_weight=(inference_tw | inference_fw), temporal_weight=inference_tw, memory_len=3)
    return mod.IDOL.inference(me, {k: v.clone() for k, v in outputs.items()}, tracker, ORI, IMG)


def main():
    out = {}
    for name, kw in CASES.items():
        flags = {k: v for k, v in kw.items() if k != "seed"}
        res = run_reference(synthetic_track_outputs(seed=kw["seed"]), **flags)
        n, T = len(res["pred_scores"]), 14
        present = np.zeros((n, T), dtype=bool)
        masks = np.zeros((n, T) + ORI, dtype=bool)
        for i, per_frame in enumerate(res["pred_masks"]):
            for t, m in enumerate(per_frame):
                if m is not None:
                    present[i, t] = True
                    masks[i, t] = m.numpy()
        out[f"{name}_scores"] = np.array(res["pred_scores"], dtype=np.float32)
        out[f"{name}_labels"] = np.array(res["pred_labels"], dtype=np.int64)
        out[f"{name}_present"] = present
        out[f"{name}_masks"] = np.packbits(masks.reshape(n, -1), axis=1)
        print(name, "instances", n, "labels", res["pred_labels"], "frames present", present.sum(1).tolist())
    path = os.path.join(HERE, "idol_tracker_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
