"""Backbone drop-ins at the reference's BACKBONE_REGISTRY boundary (SURVEY §8b "Registry boundary"):
`build_resnet_backbone(cfg, input_shape)` (detectron2/modeling/backbone/resnet.py:613-694) and
`D2SwinTransformer(cfg, input_shape)` (projects/IDOL/idol/backbone/swin.py:741-768) return an object with
`forward(x[N,3,H,W]) -> {"res2".."res5": [N,C,h,w]}`, `output_shape()` and `size_divisibility`, which is what
`MaskedBackbone` (projects/SeqFormer/seqformer/seqformer.py:35-55) consumes.  The work runs on the same sm_100a kernels
as inside the engines (tcgen05 GEMM 1x1 convs with fused FrozenBN/ReLU/residual, implicit-GEMM 3x3 convs, window
attention); weights come from a D2-layout state dict (`stem.conv1.weight`, `res2.0.conv1.norm.running_mean`, ... or the
Swin tree) via `load_reference_state_dict`.  No CPU fallback: construction on a CPU device raises."""
from collections import namedtuple

import torch
from torch import nn

from .seqformer import packing
from .seqformer.engine import SeqFormerEngine

ShapeSpec = namedtuple("ShapeSpec", ["channels", "height", "width", "stride"])      # detectron2.layers.ShapeSpec


class _BackboneEngine(SeqFormerEngine):
    """The slice of SeqFormerEngine the backbones use (device, precision switches, `lin`, 3x3 conv cache)."""

    def __init__(self, device, precision):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("vnext_b200 backbones: CUDA only (no CPU fallback)")
        self.precision = precision
        self._msda_impl = None
        self.native = True
        self.native_gemm = precision != "fp32"
        self._plans, self._khwc = {}, {}
        self.swin = None


class B200Backbone(nn.Module):
    """`forward(x)`: x [N,3,H,W] float, already normalised (the reference normalises in `preprocess_image`)."""

    def __init__(self, kind, swin_cfg=None, device="cuda", precision="tf32", out_features=("res2", "res3", "res4", "res5")):
        super().__init__()
        self.kind, self.swin_cfg, self.device_name, self.precision = kind, swin_cfg, device, precision
        self._out_features = tuple(out_features)
        self.eng = None
        if kind == "resnet50":
            self._channels = {"res2": 256, "res3": 512, "res4": 1024, "res5": 2048}
        else:
            e = swin_cfg["embed_dim"]
            self._channels = {f"res{i + 2}": e * 2 ** i for i in range(4)}
        self._strides = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}

    @property
    def size_divisibility(self):
        return 32 if self.kind == "swin" else 0             # swin.py:766-768; Backbone default (backbone.py:38-46)

    def output_shape(self):
        return {n: ShapeSpec(self._channels[n], None, None, self._strides[n]) for n in self._out_features}

    def load_reference_state_dict(self, state_dict, prefix=""):
        eng = _BackboneEngine(self.device_name, self.precision)
        dev = eng.device
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        cl = lambda w: f32(w).contiguous(memory_format=torch.channels_last)
        if self.kind == "swin":
            from .swin import SwinBackbone
            eng.swin = SwinBackbone(state_dict, prefix, self.swin_cfg, dev, eng)
        else:
            bb = packing.pack_resnet50(state_dict, prefix)
            eng.stem = (cl(bb["stem"][0]), f32(bb["stem"][1]))
            eng.stages = []
            for name, blocks in bb["stages"]:
                out = []
                for blk in blocks:
                    o = {"stride": blk["stride"]}
                    for k in ("conv1", "conv2", "conv3", "shortcut"):
                        o[k] = None if blk[k] is None else (cl(blk[k][0]), f32(blk[k][1]))
                    out.append(o)
                eng.stages.append((name, out))
        self.eng = eng
        return self

    @torch.no_grad()
    def forward(self, x):
        if self.eng is None:
            raise RuntimeError("no weights loaded: call load_reference_state_dict(backbone_state_dict)")
        x = x.to(self.eng.device, torch.float32)
        if self.kind == "swin" and (x.shape[-2] % 32 or x.shape[-1] % 32):
            raise ValueError("D2SwinTransformer drop-in: H and W must be multiples of size_divisibility (32)")
        feats = self.eng.swin.forward(x) if self.kind == "swin" else self.eng.backbone_native(x)
        return {n: feats[n].permute(0, 3, 1, 2) for n in self._out_features}       # NCHW views of the NHWC activations


def _get(cfg, path, default):
    from .seqformer.meta_arch import _cfg_get
    return _cfg_get(cfg, path, default) if cfg is not None else default


def build_resnet_backbone(cfg=None, input_shape=None, **kw):
    depth = _get(cfg, "MODEL.RESNETS.DEPTH", 50)
    if depth not in (50, 101, 152) or _get(cfg, "MODEL.RESNETS.STRIDE_IN_1X1", False):
        raise NotImplementedError("vnext_b200 builds the reference's bottleneck ResNets (depth 50 / 101 / 152), STRIDE_IN_1X1: "
                                  "False (configs/base_ytvis.yaml)")
    return B200Backbone("resnet50", out_features=_get(cfg, "MODEL.RESNETS.OUT_FEATURES", ("res2", "res3", "res4", "res5")), **kw)


def D2SwinTransformer(cfg=None, input_shape=None, **kw):
    swin = dict(embed_dim=_get(cfg, "MODEL.SWIN.EMBED_DIM", 96), depths=tuple(_get(cfg, "MODEL.SWIN.DEPTHS", (2, 2, 6, 2))),
                num_heads=tuple(_get(cfg, "MODEL.SWIN.NUM_HEADS", (3, 6, 12, 24))),
                window_size=_get(cfg, "MODEL.SWIN.WINDOW_SIZE", 7))
    return B200Backbone("swin", swin, out_features=_get(cfg, "MODEL.SWIN.OUT_FEATURES", ("res2", "res3", "res4", "res5")), **kw)


def register_with_detectron2():
    try:
        from detectron2.modeling import BACKBONE_REGISTRY
    except Exception:
        return False
    for fn in (build_resnet_backbone, D2SwinTransformer):
        if fn.__name__ not in BACKBONE_REGISTRY:
            BACKBONE_REGISTRY.register(fn)
    return True


register_with_detectron2()
