"""`SeqFormer` META_ARCH facade with the reference's model API (projects/SeqFormer/seqformer/seqformer.py).

`SeqFormer.forward(batched_inputs)` in eval mode takes the same `list[dict]` the reference's
YTVISDatasetMapper produces — {"image": [Tensor CHW x T], "height", "width", "file_names"} — and returns
the same dict {"image_size", "pred_scores", "pred_labels", "pred_masks"} (seqformer.py:226-235, 403-410),
running the B200 engine underneath.  When detectron2 is importable the class registers itself under
META_ARCH_REGISTRY as "SeqFormer" (detectron2/modeling/meta_arch/build.py:16-25), so
`cfg.MODEL.META_ARCHITECTURE: "SeqFormer"` in projects/SeqFormer/train_net.py resolves to it unchanged.
Training (`self.training`) is not part of this round's path and raises.
"""
import torch
from torch import nn

from .engine import SeqFormerEngine
from .runner import ClipRunner
from .state_spec import seqformer_r50_state_spec


class SeqFormer(nn.Module):
    def __init__(self, cfg=None, state_dict=None, device="cuda", precision="tf32", **overrides):
        super().__init__()
        g = (lambda path, default: _cfg_get(cfg, path, default)) if cfg is not None else (lambda path, default: default)
        self.num_frames = g("INPUT.SAMPLING_FRAME_NUM", 5)
        self.num_classes = g("MODEL.SeqFormer.NUM_CLASSES", 40)
        self.num_queries = g("MODEL.SeqFormer.NUM_OBJECT_QUERIES", 300)
        self.mask_stride = g("MODEL.SeqFormer.MASK_STRIDE", 4)
        self.is_multi_cls = g("MODEL.SeqFormer.MULTI_CLS_ON", True)
        self.apply_cls_thres = g("MODEL.SeqFormer.APPLY_CLS_THRES", 0.05)
        self.clip_matching = g("MODEL.SeqFormer.CLIP_MATCHING", False)
        self.clip_length = g("MODEL.SeqFormer.CLIP_LENGTH", 5)          # S/config.py:58-60
        self.clip_stride = g("MODEL.SeqFormer.CLIP_STRIDE", 1)
        self.pixel_mean = tuple(g("MODEL.PIXEL_MEAN", (123.675, 116.280, 103.530)))
        self.pixel_std = tuple(g("MODEL.PIXEL_STD", (58.395, 57.120, 57.375)))
        self.device_name = g("MODEL.DEVICE", device)
        self.precision = precision
        # "bool": pred_masks = boolean [T, H, W] tensors like the reference; "rle": pred_masks = per frame COCO RLE dicts, encoded on
        # the device — what ytvis_eval.instances_to_coco_json_video makes of them next (see vnext_b200/seqformer/ytvis_eval.py)
        self.mask_format = "bool"
        # backbone: build_resnet_backbone (R50) or D2SwinTransformer (cfg.MODEL.SWIN.*, S/config.py:67-83)
        self.swin = None
        self.resnet_depth = g("MODEL.RESNETS.DEPTH", 50)          # 50 / 101 (configs/*_r101.yaml) / 152
        if g("MODEL.BACKBONE.NAME", "build_resnet_backbone") == "D2SwinTransformer":
            self.swin = dict(embed_dim=g("MODEL.SWIN.EMBED_DIM", 96), depths=tuple(g("MODEL.SWIN.DEPTHS", (2, 2, 6, 2))),
                             num_heads=tuple(g("MODEL.SWIN.NUM_HEADS", (3, 6, 12, 24))),
                             window_size=g("MODEL.SWIN.WINDOW_SIZE", 7))
        self._cfg_get = g
        self.dropout = None                     # None: cfg.MODEL.SeqFormer.DROPOUT (0.1)
        self._state_dict = None
        self._train_model = None
        self.__dict__.update(overrides)
        self.engine = None
        self._runners = {}
        if state_dict is not None:
            self.load_reference_state_dict(state_dict)

    def expected_state_spec(self):
        return seqformer_r50_state_spec(self.num_queries, self.num_classes, swin=self.swin, depth=self.resnet_depth)

    def load_reference_state_dict(self, state_dict):
        """Takes a reference-layout checkpoint (`model` dict of a DetectionCheckpointer .pth)."""
        self._state_dict, self._train_model = state_dict, None
        spec = self.expected_state_spec()
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise KeyError(f"checkpoint is missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, shp in spec.items():
            if shp and shp[0] == "index":              # structural integer buffer (Swin relative_position_index)
                shp = (shp[1] ** 2, shp[1] ** 2)
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {tuple(shp)}, got {tuple(state_dict[k].shape)}")
        self.engine = SeqFormerEngine(state_dict, device=self.device_name, pixel_mean=self.pixel_mean,
                                      pixel_std=self.pixel_std, mask_stride=self.mask_stride,
                                      precision=self.precision, swin=self.swin)
        self._runners = {}

    def runner(self, T, H, W, ori_size, slot=0):
        key = (T, H, W, tuple(ori_size), slot, self.mask_format)
        if key not in self._runners:
            self._runners[key] = ClipRunner(self.engine, T, H, W, ori_size, mask_format=self.mask_format)
        return self._runners[key]

    def stream(self, videos, depth=2):
        """Throughput-oriented evaluation: `for out in model.stream(iterable_of_batched_inputs)` yields exactly what
        `model(batched_inputs)` returns, in order, while `depth` clips are in flight on alternating runners (own CUDA
        graph, stream and pinned buffers each): clip i+1's staging / H2D / compute overlaps clip i's D2H and host-side
        list building.  Same results as the synchronous call; whole-video (non clip-matching) inference only."""
        if self.training or self.clip_matching:
            raise NotImplementedError("stream() serves whole-video eval inference")
        if self.engine is None:
            raise RuntimeError("no weights loaded: call load_reference_state_dict(checkpoint['model'])")
        from collections import deque
        pending = deque()

        def finish():
            run, ori = pending.popleft()
            scores, masks = run.collect()
            return SeqFormerEngine.finish_video(scores, masks, ori, self.apply_cls_thres, self.is_multi_cls)
        for n, batched_inputs in enumerate(videos):
            video = batched_inputs[0]
            frames = video["image"]
            H, W = frames[0].shape[-2:]
            ori = (video.get("height", H), video.get("width", W))
            if len(pending) == depth:
                yield finish()
            if (len(frames), H, W, tuple(ori), n % depth, self.mask_format) not in self._runners:
                # a runner's construction (eager warm-up + graph capture) never overlaps a clip in flight: drain first,
                # then build every slot of this geometry at once
                while pending:
                    yield finish()
                for slot in range(depth):
                    self.runner(len(frames), H, W, ori, slot=slot)
            run = self.runner(len(frames), H, W, ori, slot=n % depth)
            run.submit(frames_list=frames)
            pending.append((run, ori))
        while pending:
            yield finish()

    # ------------------------------------------------------------------ training (seqformer.py:217-225, :266-305)
    def train_model(self, num_frames):
        """The differentiable dense path (vnext_b200/seqformer/train.py), created on first use from the loaded weights."""
        if self._train_model is None or self._train_model.nf != num_frames:
            if self._state_dict is None:
                raise RuntimeError("no weights loaded: call load_reference_state_dict(checkpoint['model'])")
            if self.swin is not None:
                raise NotImplementedError("the training path is built for the ResNet-50 backbone; Swin training needs the "
                                          "window-attention backward, not built this round")
            from .criterion import default_weight_dict
            from .train import SeqFormerTrainModel
            g = self._cfg_get
            wd = default_weight_dict(g("MODEL.SeqFormer.CLASS_WEIGHT", 2.0), g("MODEL.SeqFormer.L1_WEIGHT", 5.0),
                                     g("MODEL.SeqFormer.GIOU_WEIGHT", 2.0), g("MODEL.SeqFormer.MASK_WEIGHT", 2.0),
                                     g("MODEL.SeqFormer.DICE_WEIGHT", 5.0), g("MODEL.SeqFormer.DEC_LAYERS", 6),
                                     g("MODEL.SeqFormer.DEEP_SUPERVISION", True))
            self._train_model = SeqFormerTrainModel(
                self._state_dict, num_frames, device=self.device_name, num_classes=self.num_classes,
                num_queries=self.num_queries, dropout=g("MODEL.SeqFormer.DROPOUT", 0.1) if self.dropout is None else self.dropout,
                pixel_mean=self.pixel_mean, pixel_std=self.pixel_std, mask_stride=self.mask_stride, weight_dict=wd)
        return self._train_model

    def prepare_targets(self, batched_inputs):
        """Per video: `Instances`-like objects of its frames -> {"labels","boxes" [n,nf,4],"masks" [n,nf,H,W],"size"}
        (seqformer.py:266-305: an instance's label is the max over frames of class x valid)."""
        from ..idol.criterion import xyxy_to_cxcywh
        dev = self.device_name
        out = []
        for video in batched_inputs:
            boxes, masks, classes = [], [], []
            for t in video["instances"]:
                h, w = t.image_size
                b = t.gt_boxes.tensor.to(dev) / torch.as_tensor([w, h, w, h], dtype=torch.float, device=dev)
                boxes.append(xyxy_to_cxcywh(b))
                masks.append(t.gt_masks.tensor.to(dev))
                classes.append(t.gt_classes.to(dev) * (t.gt_ids.to(dev) != -1))
            out.append({"labels": torch.stack(classes, 0).max(0)[0], "boxes": torch.stack(boxes, 1),
                        "masks": torch.stack(masks, 1), "size": torch.as_tensor([h, w], dtype=torch.long, device=dev)})
        return out

    def forward(self, batched_inputs):
        if self.training:
            frames = [f for video in batched_inputs for f in video["image"]]
            model = self.train_model(len(batched_inputs[0]["image"])).train()
            loss_dict, _ = model(frames, self.prepare_targets(batched_inputs))
            return model.weighted(loss_dict)
        if self.engine is None:
            raise RuntimeError("no weights loaded: call load_reference_state_dict(checkpoint['model'])")
        video = batched_inputs[0]
        frames = video["image"]
        T = len(frames)
        H, W = frames[0].shape[-2:]
        ori = (video.get("height", H), video.get("width", W))
        if self.clip_matching:
            return self._forward_clip_matching(frames, ori)
        run = self.runner(T, H, W, ori)
        scores, masks = run.run_host(frames_list=frames)
        return SeqFormerEngine.finish_video(scores, masks, ori, self.apply_cls_thres, self.is_multi_cls)


def _clip_matching_forward(self, frames, ori):
    """Long videos: sliding clips stitched by mask overlap (seqformer.py:236-262; clip_matching.py)."""
    from .clip_matching import ClipTracks, clip_matching_postprocess, clip_windows
    L = len(frames)
    H, W = frames[0].shape[-2:]
    dev = self.engine.device
    tracks = None
    for frame_idx in clip_windows(L, self.clip_length, self.clip_stride):
        clip = torch.stack([frames[i] for i in frame_idx]).to(dev, non_blocking=True)
        if clip.dtype not in (torch.uint8, torch.float32):
            clip = clip.float()
        cls, logits = self.engine.infer_clip_device(clip)
        if tracks is None:
            tracks = ClipTracks(self.clip_length, L, self.num_classes, tuple(logits.shape[-2:]), dev)
        tracks.update(frame_idx, cls, logits)
    pred_cls, pred_logits = tracks.get_result()
    return clip_matching_postprocess(pred_cls, pred_logits, ori, (H, W), self.mask_stride, self.is_multi_cls,
                                     self.apply_cls_thres, mask_format=self.mask_format)


SeqFormer._forward_clip_matching = _clip_matching_forward


def _cfg_get(cfg, path, default):
    node = cfg
    for part in path.split("."):
        if isinstance(node, dict):
            if part not in node:
                return default
            node = node[part]
        else:
            if not hasattr(node, part):
                return default
            node = getattr(node, part)
    return node


def register_with_detectron2():
    """Register under the reference's registry name; returns False when detectron2 is absent."""
    try:
        from detectron2.modeling import META_ARCH_REGISTRY
    except Exception:
        return False
    if "SeqFormer" not in META_ARCH_REGISTRY:
        META_ARCH_REGISTRY.register(SeqFormer)
    return True


register_with_detectron2()
