"""`IDOL` META_ARCH facade (projects/IDOL/idol/idol.py:71-295): same cfg keys and `batched_inputs`; the dense
per-frame path runs on the B200 engine in chunks of BATCH_INFER_LEN frames (idol.py:252-274).  The online tracker
(`IDOL_Tracker`, idol.py:275-295 / models/tracker.py) is host-side association and is NOT part of this round's path
(SURVEY §8f): `forward` hands the dense outputs to `self.tracker_fn(output, (height, width), image_size)` when one is
attached (e.g. the reference's own `IDOL.inference` bound to its tracker) and raises otherwise; `dense_forward`
returns exactly what the reference's `detr.inference_forward` returns."""
import torch
from torch import nn

from ..seqformer.meta_arch import _cfg_get
from .engine import IDOLEngine
from .state_spec import idol_r50_state_spec


class IDOL(nn.Module):
    def __init__(self, cfg=None, state_dict=None, device="cuda", precision="tf32", tracker_fn=None):
        super().__init__()
        g = (lambda p, d: _cfg_get(cfg, p, d)) if cfg is not None else (lambda p, d: d)
        self.num_classes = g("MODEL.IDOL.NUM_CLASSES", 40)
        self.num_queries = g("MODEL.IDOL.NUM_OBJECT_QUERIES", 300)
        self.mask_stride = g("MODEL.IDOL.MASK_STRIDE", 4)
        self.batch_infer_len = g("MODEL.IDOL.BATCH_INFER_LEN", 10)
        self.pixel_mean = tuple(g("MODEL.PIXEL_MEAN", (123.675, 116.280, 103.530)))
        self.pixel_std = tuple(g("MODEL.PIXEL_STD", (58.395, 57.120, 57.375)))
        self.device_name = g("MODEL.DEVICE", device)
        self.precision = precision
        self.tracker_fn = tracker_fn
        self.engine = None
        if state_dict is not None:
            self.load_reference_state_dict(state_dict)

    def load_reference_state_dict(self, state_dict):
        spec = idol_r50_state_spec(self.num_queries, self.num_classes)
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise KeyError(f"checkpoint is missing {len(missing)} keys, e.g. {missing[:3]}")
        self.engine = IDOLEngine(state_dict, device=self.device_name, pixel_mean=self.pixel_mean,
                                 pixel_std=self.pixel_std, mask_stride=self.mask_stride, precision=self.precision)

    @torch.no_grad()
    def dense_forward(self, batched_inputs):
        if self.engine is None:
            raise RuntimeError("no weights loaded: call load_reference_state_dict(checkpoint['model'])")
        frames = batched_inputs[0]["image"]
        dev = self.engine.device
        outs = []
        for s in range(0, len(frames), self.batch_infer_len):
            chunk = torch.stack([f.to(dev, non_blocking=True) for f in frames[s:s + self.batch_infer_len]])
            if chunk.dtype not in (torch.uint8, torch.float32):
                chunk = chunk.float()
            o = self.engine.inference_forward(chunk)
            outs.append({k: o[k] for k in ("pred_logits", "pred_boxes", "pred_inst_embed", "pred_masks")})
        return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}

    def forward(self, batched_inputs):
        if self.training:
            raise NotImplementedError("vnext_b200 IDOL: the training step is not on this round's hot path")
        output = self.dense_forward(batched_inputs)
        if self.tracker_fn is None:
            raise NotImplementedError("IDOL's online tracker (models/tracker.py) is host-side post-processing outside "
                                      "this round's path; attach tracker_fn or use dense_forward()")
        v = batched_inputs[0]
        H, W = v["image"][0].shape[-2:]
        return self.tracker_fn(output, (v.get("height", H), v.get("width", W)), (H, W))


def register_with_detectron2():
    try:
        from detectron2.modeling import META_ARCH_REGISTRY
    except Exception:
        return False
    if "IDOL" not in META_ARCH_REGISTRY:
        META_ARCH_REGISTRY.register(IDOL)
    return True


register_with_detectron2()
