"""Generate golden vectors for the SeqFormer inference path from the REFERENCE's own modules.

Run ONLY in the build container (needs /root/reference):

    python tests/golden/make_seqformer_golden.py

The reference package `projects/SeqFormer/seqformer` is mounted as a namespace-like package
(`refseq`) WITHOUT running its __init__ (which needs detectron2/fvcore); three tiny stubs stand in for
absent third-party modules: `MultiScaleDeformableAttention` (compiled op; replaced by the reference's
own CPU function ms_deform_attn_core_pytorch), `fvcore.nn` (loss-only symbols, never called here) and
the D2 backbone, replaced by torchvision's resnet50 + FrozenBatchNorm2d (identical topology for
STRIDE_IN_1X1=False; key mapping = the reference's tools/convert-torchvision-to-d2.py:38-49).
Everything else — DeformableTransformer (encoder/decoder, SeqFormer MSDeformAttn), DeformableDETR,
CondInst_segm.inference, MaskHeadSmallConv, dynamic mask head, PositionEmbeddingSine, Joiner,
nested_tensor_from_tensor_list — is the reference code, unmodified.

Weights are synthesised from state-dict key names (vnext_b200.synth.fill_state_dict) so tests can
rebuild them; only outputs are stored.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
import torchvision
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from vnext_b200.synth import fill_state_dict, synthetic_clip  # noqa: E402

REF_PKG = "/root/reference/projects/SeqFormer/seqformer"
PIXEL_MEAN = (123.675, 116.280, 103.530)
PIXEL_STD = (58.395, 57.120, 57.375)

CASES = {
    # name: (T, H, W, num_queries, seed)
    "nopad": (3, 96, 128, 300, 0),   # 300 queries: inference hard-codes num_insts=[300] (segm:338)
    "pad": (2, 80, 112, 300, 1),     # padded to 96 x 128 -> masks / valid ratios matter
}


def mount_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    fv = types.ModuleType("fvcore")
    fvnn = types.ModuleType("fvcore.nn")
    fvnn.giou_loss = fvnn.smooth_l1_loss = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("loss stub"))
    fv.nn = fvnn
    sys.modules.setdefault("fvcore", fv)
    sys.modules.setdefault("fvcore.nn", fvnn)
    pkg = types.ModuleType("refseq")
    pkg.__path__ = [REF_PKG]
    sys.modules["refseq"] = pkg
    func = importlib.import_module("refseq.models.ops.functions.ms_deform_attn_func")
    mod = importlib.import_module("refseq.models.ops.modules.ms_deform_attn")

    class CpuMSDA:   # the reference's own CPU path behind the autograd Function's call signature
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            return func.ms_deform_attn_core_pytorch(value, shapes, loc, attn)
    mod.MSDeformAttnFunction = CpuMSDA
    return pkg


def tv_to_d2_key(k):
    """torchvision resnet key -> detectron2 key (tools/convert-torchvision-to-d2.py:38-49)."""
    k0 = k
    if "layer" not in k:
        k = "stem." + k
    for t in (1, 2, 3, 4):
        k = k.replace(f"layer{t}", f"res{t + 1}")
    for t in (1, 2, 3):
        k = k.replace(f"bn{t}", f"conv{t}.norm")
    k = k.replace("downsample.0", "shortcut").replace("downsample.1", "shortcut.norm")
    assert k != k0 or "layer" in k0
    return k


class TVMaskedBackbone(nn.Module):
    """Stand-in for seqformer.py:35-55 MaskedBackbone with a torchvision ResNet-50 body."""

    def __init__(self):
        super().__init__()
        from torchvision.ops.misc import FrozenBatchNorm2d
        self.body = torchvision.models.resnet50(weights=None, norm_layer=FrozenBatchNorm2d)
        del self.body.fc
        self.feature_strides = [4, 8, 16, 32]
        self.num_channels = [256, 512, 1024, 2048]

    def forward(self, tensor_list):
        from refseq.util.misc import NestedTensor
        b = self.body
        x = b.maxpool(b.relu(b.bn1(b.conv1(tensor_list.tensors))))
        out = {}
        for name, layer in (("res2", b.layer1), ("res3", b.layer2), ("res4", b.layer3), ("res5", b.layer4)):
            x = layer(x)
            m = tensor_list.mask
            mask = F.interpolate(m[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
            out[name] = NestedTensor(x, mask)
        return out


def build_reference_model(num_queries, num_frames, num_classes=40):
    from refseq.models.backbone import Joiner
    from refseq.models.deformable_detr import DeformableDETR
    from refseq.models.deformable_transformer import DeformableTransformer
    from refseq.models.position_encoding import PositionEmbeddingSine
    from refseq.models.segmentation_condInst import CondInst_segm
    mb = TVMaskedBackbone()
    backbone = Joiner(mb, PositionEmbeddingSine(128, normalize=True))
    backbone.num_channels = mb.num_channels[1:]
    backbone.strides = mb.feature_strides[1:]
    transformer = DeformableTransformer(d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6,
                                        dim_feedforward=1024, dropout=0.1, activation="relu",
                                        return_intermediate_dec=True, num_frames=num_frames, num_feature_levels=4,
                                        dec_n_points=4, enc_n_points=4)
    model = DeformableDETR(backbone, transformer, num_classes=num_classes, num_frames=num_frames,
                           num_queries=num_queries, num_feature_levels=4, aux_loss=True, with_box_refine=True)
    seg = CondInst_segm(model, freeze_detr=False, rel_coord=True)

    class Wrapper(nn.Module):    # gives the `detr.` prefix of SeqFormer.detr (seqformer.py:164)
        def __init__(self):
            super().__init__()
            self.detr = seg
    return Wrapper().eval()


def reference_state_spec(model):
    """Reference-layout (D2 backbone naming) key -> tensor, and the mapping back to the stand-in's keys."""
    spec, back = {}, {}
    for k, v in model.state_dict().items():
        pre = "detr.detr.backbone.0.body."
        if k.startswith(pre):
            k2 = "detr.detr.backbone.0.backbone." + tv_to_d2_key(k[len(pre):])
        else:
            k2 = k
        spec[k2] = v
        back[k2] = k
    return spec, back


class Clip(list):
    image_sizes = None


def run_case(name):
    T, H, W, nq, seed = CASES[name]
    model = build_reference_model(nq, T)
    spec, back = reference_state_spec(model)
    sd = fill_state_dict(spec, seed)
    model.load_state_dict({back[k]: v for k, v in sd.items()}, strict=True)
    frames = synthetic_clip(T, H, W, seed)
    mean, std = torch.tensor(PIXEL_MEAN).view(3, 1, 1), torch.tensor(PIXEL_STD).view(3, 1, 1)
    clip = Clip([(f - mean) / std for f in frames])
    clip.image_sizes = [(H, W)] * T
    model.detr.detr.num_frames = T
    with torch.no_grad():
        out = model.detr.inference(clip)
        # intermediates for finer-grained pins
        from refseq.util.misc import nested_tensor_from_tensor_list
        samples = nested_tensor_from_tensor_list(clip, size_divisibility=32)
        feats, pos = model.detr.detr.backbone(samples)
    res = {
        f"{name}_pred_logits": out["pred_logits"].numpy(),
        f"{name}_pred_boxes": out["pred_boxes"].numpy(),
        f"{name}_pred_masks": out["pred_masks"].numpy()[:, ::10],       # every 10th query
        f"{name}_res3_sample": feats[1].tensors.flatten()[::613].numpy(),
        f"{name}_res5": feats[3].tensors.numpy()[:, ::16],
        f"{name}_pos3": pos[3].numpy()[:1, ::8],
    }
    print(name, {k: (v.shape, float(np.abs(v).mean())) for k, v in res.items()})
    return res, len(spec)


def main():
    mount_reference()
    out = {}
    for name in CASES:
        res, nkeys = run_case(name)
        out.update(res)
    out["num_state_keys"] = np.array([nkeys])
    path = os.path.join(HERE, "seqformer_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
