"""`IDOL` META_ARCH facade (projects/IDOL/idol/idol.py:71-295): same cfg keys and `batched_inputs`; the dense
per-frame path runs on the B200 engine in chunks of BATCH_INFER_LEN frames (idol.py:252-274).  In eval mode `forward`
then runs the online association + video post-processing of idol.py:275-295 / :326-473 (vnext_b200/idol/tracker.py:
batched device ops + host bookkeeping, pinned to the reference's own outputs), or hands the dense outputs to
`self.tracker_fn(output, (height, width), image_size)` when one is attached (e.g. the reference's `IDOL.inference`); `dense_forward`
returns exactly what the reference's `detr.inference_forward` returns.  In training mode `forward` returns the weighted
loss dict of idol.py:205-215 from the differentiable dense path (vnext_b200/idol/train.py: MSDA forward/backward
kernels under autograd) and the host-side criterion (criterion.py)."""
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
        self.dropout = g("MODEL.IDOL.DROPOUT", 0.1)
        self.is_multi_cls = g("MODEL.IDOL.MULTI_CLS_ON", True)                  # evaluation keys, idol/config.py:51-65
        self.apply_cls_thres = g("MODEL.IDOL.APPLY_CLS_THRES", 0.05)
        self.temporal_score_type = g("MODEL.IDOL.TEMPORAL_SCORE_TYPE", "mean")
        self.inference_select_thres = g("MODEL.IDOL.INFERENCE_SELECT_THRES", 0.1)
        self.inference_fw = g("MODEL.IDOL.INFERENCE_FW", True)
        self.inference_tw = g("MODEL.IDOL.INFERENCE_TW", True)
        self.memory_len = g("MODEL.IDOL.MEMORY_LEN", 3)
        self.nms_pre = g("MODEL.IDOL.NMS_PRE", 0.5)
        self.add_new_score = g("MODEL.IDOL.ADD_NEW_SCORE", 0.2)
        from .criterion import default_weight_dict
        self.weight_dict = default_weight_dict(g("MODEL.IDOL.CLASS_WEIGHT", 2.0), g("MODEL.IDOL.L1_WEIGHT", 5.0),
                                               g("MODEL.IDOL.GIOU_WEIGHT", 2.0), g("MODEL.IDOL.REID_WEIGHT", 2.0),
                                               g("MODEL.IDOL.MASK_WEIGHT", 2.0), g("MODEL.IDOL.DICE_WEIGHT", 5.0),
                                               g("MODEL.IDOL.DEC_LAYERS", 6), g("MODEL.IDOL.DEEP_SUPERVISION", True))
        # backbone: build_resnet_backbone (R50) or D2SwinTransformer (cfg.MODEL.SWIN.*, I/config.py, configs/*swinL.yaml)
        self.mask_format = "bool"         # "rle": pred_masks as per-frame COCO RLE dicts, encoded on the device (vnext_b200/rle.py)
        self.swin = None
        self.resnet_depth = g("MODEL.RESNETS.DEPTH", 50)          # 50 / 101 (configs/ytvis19_r101.yaml) / 152
        if "swin" in str(g("MODEL.BACKBONE.NAME", "build_resnet_backbone")).lower():
            self.swin = dict(embed_dim=g("MODEL.SWIN.EMBED_DIM", 96), depths=tuple(g("MODEL.SWIN.DEPTHS", (2, 2, 6, 2))),
                             num_heads=tuple(g("MODEL.SWIN.NUM_HEADS", (3, 6, 12, 24))),
                             window_size=g("MODEL.SWIN.WINDOW_SIZE", 7))
        self.engine = None
        self._state_dict = None
        self._train_model = None
        if state_dict is not None:
            self.load_reference_state_dict(state_dict)

    def load_reference_state_dict(self, state_dict, keep_train_model=False):
        self._state_dict = state_dict
        if not keep_train_model:
            self._train_model = None
        spec = idol_r50_state_spec(self.num_queries, self.num_classes, swin=self.swin, depth=self.resnet_depth)
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise KeyError(f"checkpoint is missing {len(missing)} keys, e.g. {missing[:3]}")
        self.engine = IDOLEngine(state_dict, device=self.device_name, pixel_mean=self.pixel_mean,
                                 pixel_std=self.pixel_std, mask_stride=self.mask_stride, precision=self.precision,
                                 swin=self.swin)

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

    # ------------------------------------------------------------------ training (idol.py:205-215, :296-324)
    def train_model(self):
        """The differentiable dense path (vnext_b200/idol/train.py), created on first use from the loaded weights."""
        if self._train_model is None:
            if self._state_dict is None:
                raise RuntimeError("no weights loaded: call load_reference_state_dict(checkpoint['model'])")
            if self.swin is not None:
                raise NotImplementedError("the training path is built for the ResNet-50 backbone (BASELINE config 3); "
                                          "Swin training needs the window-attention backward, not built this round")
            from .train import IDOLTrainModel
            self._train_model = IDOLTrainModel(self._state_dict, device=self.device_name, num_classes=self.num_classes,
                                               num_queries=self.num_queries, dropout=self.dropout,
                                               pixel_mean=self.pixel_mean, pixel_std=self.pixel_std,
                                               mask_stride=self.mask_stride, weight_dict=self.weight_dict)
        return self._train_model

    def sync_engine(self):
        """Rebuild the inference engine from the trained weights (call before evaluating after training steps)."""
        self.load_reference_state_dict(self.train_model().reference_state_dict(), keep_train_model=True)

    def prepare_targets(self, targets):
        """Per-frame `Instances`-like objects (gt_classes, gt_boxes.tensor xyxy pixels, gt_masks.tensor, gt_ids,
        image_size) in key, ref, key, ref order -> (det_targets, ref_targets), objects absent from the key frame
        dropped from both (idol.py:296-324)."""
        from .criterion import xyxy_to_cxcywh
        dev = self.device_name
        new = []
        for t in targets:
            h, w = t.image_size
            boxes = t.gt_boxes.tensor.to(dev) / torch.as_tensor([w, h, w, h], dtype=torch.float, device=dev)
            ids = t.gt_ids.to(dev)
            new.append({"labels": t.gt_classes.to(dev), "boxes": xyxy_to_cxcywh(boxes),
                        "masks": t.gt_masks.tensor.to(dev), "inst_id": ids, "valid": ids != -1})
        det, ref = new[0::2], new[1::2]
        for d, r in zip(det, ref):
            if not bool(d["valid"].all()):
                keep = d["valid"].clone()
                for k in list(d):
                    d[k] = d[k][keep]
                for k in list(r):
                    r[k] = r[k][keep]
        return det, ref

    def forward(self, batched_inputs):
        if self.training:
            frames = [f for video in batched_inputs for f in video["image"]]
            gt = [inst for video in batched_inputs for inst in video["instances"]]
            det, ref = self.prepare_targets(gt)
            model = self.train_model().train()
            loss_dict, _ = model(frames, det, ref)
            return model.weighted(loss_dict)
        output = self.dense_forward(batched_inputs)
        v = batched_inputs[0]
        H, W = v["image"][0].shape[-2:]
        ori_size, image_size = (v.get("height", H), v.get("width", W)), (H, W)
        if self.tracker_fn is not None:                      # e.g. the reference's own IDOL.inference bound to its tracker
            return self.tracker_fn(output, ori_size, image_size)
        from .tracker import IDOLTracker, video_inference
        tracker = IDOLTracker(init_score_thr=0.2, obj_score_thr=0.1, nms_thr_pre=self.nms_pre, nms_thr_post=0.05,
                              addnew_score_thr=self.add_new_score, memo_tracklet_frames=10, memo_momentum=0.8,
                              long_match=self.inference_tw, frame_weight=(self.inference_tw | self.inference_fw),
                              temporal_weight=self.inference_tw, memory_len=self.memory_len)        # idol.py:276-290
        return video_inference(output, tracker, ori_size, image_size, select_thres=self.inference_select_thres,
                               temporal_score_type=self.temporal_score_type, multi_cls=self.is_multi_cls,
                               apply_cls_thres=self.apply_cls_thres, mask_stride=self.mask_stride,
                               mask_format=self.mask_format)


def register_with_detectron2():
    try:
        from detectron2.modeling import META_ARCH_REGISTRY
    except Exception:
        return False
    if "IDOL" not in META_ARCH_REGISTRY:
        META_ARCH_REGISTRY.register(IDOL)
    return True


register_with_detectron2()
