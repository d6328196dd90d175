"""Reference-layout state-dict specification of IDOL with a ResNet-50 backbone:
module tree of `IDOL` (projects/IDOL/idol/idol.py:77-203) = CondInst_segm(DeformableDETR(...)) + reid_embed_head."""
from ..seqformer.state_spec import _linear, _ln, _mha, _msda, resnet50_spec


def idol_r50_state_spec(num_queries=300, num_classes=40, C=256, ffn=1024, M=8, L=4, P=4, enc=6, dec=6, swin=None, depth=50):
    if swin is None:
        spec = resnet50_spec("detr.detr.backbone.0.backbone.", depth)
        in_ch = (512, 1024, 2048)
    else:
        from ..swin_spec import swin_state_spec
        spec = swin_state_spec("detr.detr.backbone.0.backbone.", **swin)
        in_ch = tuple(swin["embed_dim"] * 2 ** i for i in (1, 2, 3))
    T = "detr.detr.transformer."
    for i in range(enc):
        p = f"{T}encoder.layers.{i}"
        _msda(spec, p + ".self_attn", C, M, L, P, box=False)
        _ln(spec, p + ".norm1", C)
        _linear(spec, p + ".linear1", ffn, C)
        _linear(spec, p + ".linear2", C, ffn)
        _ln(spec, p + ".norm2", C)
    for i in range(dec):
        p = f"{T}decoder.layers.{i}"
        _msda(spec, p + ".cross_attn", C, M, L, P, box=False)
        _ln(spec, p + ".norm1", C)
        _mha(spec, p + ".self_attn", C)
        _ln(spec, p + ".norm2", C)
        _linear(spec, p + ".linear1", ffn, C)
        _linear(spec, p + ".linear2", C, ffn)
        _ln(spec, p + ".norm3", C)
    for i in range(dec):                       # decoder.bbox_embed aliases detr.bbox_embed (with_box_refine)
        for j, (co, ci) in enumerate(((C, C), (C, C), (4, C))):
            _linear(spec, f"{T}decoder.bbox_embed.{i}.layers.{j}", co, ci)
    spec[T + "level_embed"] = (L, C)
    _linear(spec, T + "reference_points", 2, C)
    D = "detr.detr."
    for i in range(dec):
        _linear(spec, f"{D}class_embed.{i}", num_classes, C)
        for j, (co, ci) in enumerate(((C, C), (C, C), (4, C))):
            _linear(spec, f"{D}bbox_embed.{i}.layers.{j}", co, ci)
    spec[D + "query_embed.weight"] = (num_queries, 2 * C)
    for l, (cin, k) in enumerate(((in_ch[0], 1), (in_ch[1], 1), (in_ch[2], 1), (in_ch[2], 3))):
        spec[f"{D}input_proj.{l}.0.weight"] = (C, cin, k, k)
        spec[f"{D}input_proj.{l}.0.bias"] = (C,)
        spec[f"{D}input_proj.{l}.1.weight"] = (C,)
        spec[f"{D}input_proj.{l}.1.bias"] = (C,)
    for j, (co, ci) in enumerate(((C, C), (C, C), (169, C))):
        _linear(spec, f"detr.controller.layers.{j}", co, ci)
    for j in range(3):
        _linear(spec, f"detr.reid_embed_head.layers.{j}", C, C)
    for name, co, ci in (("lay1", C // 4, C), ("lay2", C // 32, C // 4), ("lay3", C, C), ("lay4", C, C), ("dcn", C, C)):
        spec[f"detr.mask_head.{name}.weight"] = (co, ci, 3, 3)
        spec[f"detr.mask_head.{name}.bias"] = (co,)
    return spec
