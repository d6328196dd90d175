"""Golden vectors of IDOL's per-frame path from the REFERENCE's own modules (run in the build container only).
Same mounting trick as make_seqformer_golden.py, on projects/IDOL/idol; extra stub: detectron2.structures
(two names imported at module level by models/segmentation_condInst.py:719, unused on this path)."""
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
from make_seqformer_golden import PIXEL_MEAN, PIXEL_STD, TVMaskedBackbone, Clip, tv_to_d2_key  # noqa: E402
from vnext_b200.synth import fill_state_dict, synthetic_clip  # noqa: E402

CASES = {"nopad": (3, 96, 128, 0), "pad": (2, 80, 112, 1)}


def mount():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    for name in ("fvcore", "fvcore.nn", "detectron2", "detectron2.structures"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["fvcore.nn"].giou_loss = sys.modules["fvcore.nn"].smooth_l1_loss = None
    sys.modules["detectron2.structures"].Instances = sys.modules["detectron2.structures"].ROIMasks = object
    pkg = types.ModuleType("refidol")
    pkg.__path__ = ["/root/reference/projects/IDOL/idol"]
    sys.modules["refidol"] = pkg
    func = importlib.import_module("refidol.models.ops.functions.ms_deform_attn_func")
    mod = importlib.import_module("refidol.models.ops.modules.ms_deform_attn")

    class CpuMSDA:
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            return func.ms_deform_attn_core_pytorch(value, shapes, loc, attn)
    mod.MSDeformAttnFunction = CpuMSDA


def build(num_queries=300, num_classes=40):
    from refidol.models.backbone import Joiner
    from refidol.models.deformable_detr import DeformableDETR
    from refidol.models.deformable_transformer import DeformableTransformer
    from refidol.models.position_encoding import PositionEmbeddingSine
    from refidol.models.segmentation_condInst import CondInst_segm
    import make_seqformer_golden as msg
    # TVMaskedBackbone builds NestedTensor from the SeqFormer mount; give it IDOL's class (same code)
    sys.modules.setdefault("refseq", types.ModuleType("refseq"))
    sys.modules["refseq"].__path__ = [msg.REF_PKG]
    mb = TVMaskedBackbone()
    backbone = Joiner(mb, PositionEmbeddingSine(128, normalize=True))
    backbone.num_channels = mb.num_channels[1:]
    backbone.strides = mb.feature_strides[1:]
    transformer = DeformableTransformer(d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6,
                                        dim_feedforward=1024, dropout=0.1, activation="relu", return_intermediate_dec=True,
                                        num_frames=1, num_feature_levels=4, dec_n_points=4, enc_n_points=4)
    model = DeformableDETR(backbone, transformer, num_classes=num_classes, num_frames=1, num_queries=num_queries,
                           num_feature_levels=4, aux_loss=True, with_box_refine=True)
    seg = CondInst_segm(model, freeze_detr=False, rel_coord=True)

    class Wrapper(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.detr = seg
    return Wrapper().eval()


def main():
    mount()
    out = {}
    for name, (T, H, W, seed) in CASES.items():
        model = build()
        spec, back = {}, {}
        for k, v in model.state_dict().items():
            pre = "detr.detr.backbone.0.body."
            k2 = "detr.detr.backbone.0.backbone." + tv_to_d2_key(k[len(pre):]) if k.startswith(pre) else k
            spec[k2], back[k2] = v, k
        sd = fill_state_dict(spec, seed)
        model.load_state_dict({back[k]: v for k, v in sd.items()}, strict=True)
        frames = synthetic_clip(T, H, W, seed)
        mean, std = torch.tensor(PIXEL_MEAN).view(3, 1, 1), torch.tensor(PIXEL_STD).view(3, 1, 1)
        clip = Clip([(f - mean) / std for f in frames])
        clip.image_sizes = [(H, W)] * T
        with torch.no_grad():
            o = model.detr.inference_forward(clip)
        out[f"{name}_pred_logits"] = o["pred_logits"].numpy()
        out[f"{name}_pred_boxes"] = o["pred_boxes"].numpy()
        out[f"{name}_pred_inst_embed"] = o["pred_inst_embed"].numpy()[:, ::5]
        out[f"{name}_pred_masks"] = o["pred_masks"].numpy()[:, ::10]
        out["num_state_keys"] = np.array([len(spec)])
        print(name, {k: v.shape for k, v in o.items() if hasattr(v, "shape")})
    path = os.path.join(HERE, "idol_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
