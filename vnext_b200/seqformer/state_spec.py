"""Reference-layout state-dict specification (key -> shape) of SeqFormer with a ResNet-50 backbone.

This is the checkpoint boundary of SURVEY §8b: the module tree of the reference's `SeqFormer`
META_ARCH (projects/SeqFormer/seqformer/seqformer.py:133-164) yields exactly these keys:
  detr.detr.backbone.0.backbone.{stem,res2..res5}.*   D2 ResNet (detectron2/modeling/backbone/resnet.py)
  detr.detr.input_proj.{0-3}.{0,1}.*                  models/deformable_detr.py:63-84
  detr.detr.transformer.*                             models/deformable_transformer.py
  detr.detr.{class_embed,bbox_embed}.{0-5}.*, detr.detr.query_embed.weight
  detr.controller.layers.{0-2}.*, detr.mask_head.{lay1-4,dcn}.*   models/segmentation_condInst.py:61-66
Used to synthesise weights offline and to validate a loaded checkpoint before packing it for the GPU.
"""

R50_STAGES = (("res2", 3, 64, 256), ("res3", 4, 128, 512), ("res4", 6, 256, 1024), ("res5", 3, 512, 2048))
# bottleneck ResNets of detectron2/modeling/backbone/resnet.py:656-662 (`num_blocks_per_stage`); channels are the same
BLOCKS_PER_STAGE = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


def resnet_stages(depth=50):
    if depth not in BLOCKS_PER_STAGE:
        raise NotImplementedError(f"ResNet depth {depth}: the bottleneck depths 50 / 101 / 152 are built "
                                  "(projects/IDOL/configs/ytvis19_r50.yaml, ytvis19_r101.yaml)")
    return tuple((name, n, mid, cout) for (name, _, mid, cout), n in zip(R50_STAGES, BLOCKS_PER_STAGE[depth]))


def resnet_depth_of(state_dict, prefix="detr.detr.backbone.0.backbone."):
    """The bottleneck ResNet depth a checkpoint holds, from its res3 / res4 block counts (50 when it has no ResNet keys)."""
    n3 = sum(1 for k in state_dict if k.startswith(prefix + "res3.") and k.endswith(".conv1.weight"))
    n4 = sum(1 for k in state_dict if k.startswith(prefix + "res4.") and k.endswith(".conv1.weight"))
    for depth, blocks in BLOCKS_PER_STAGE.items():
        if (n3, n4) == blocks[1:3]:
            return depth
    return 50


def _conv_bn(spec, p, cout, cin, k):
    spec[p + ".weight"] = (cout, cin, k, k)
    for n in ("weight", "bias", "running_mean", "running_var"):
        spec[f"{p}.norm.{n}"] = (cout,)


def resnet50_spec(prefix, depth=50):
    spec = {}
    _conv_bn(spec, prefix + "stem.conv1", 64, 3, 7)
    cin = 64
    for stage, nblocks, mid, cout in resnet_stages(depth):
        for b in range(nblocks):
            p = f"{prefix}{stage}.{b}."
            if cin != cout:
                _conv_bn(spec, p + "shortcut", cout, cin, 1)
            _conv_bn(spec, p + "conv1", mid, cin, 1)
            _conv_bn(spec, p + "conv2", mid, mid, 3)
            _conv_bn(spec, p + "conv3", cout, mid, 1)
            cin = cout
    return spec


def _linear(spec, p, cout, cin):
    spec[p + ".weight"] = (cout, cin)
    spec[p + ".bias"] = (cout,)


def _ln(spec, p, c):
    spec[p + ".weight"] = (c,)
    spec[p + ".bias"] = (c,)


def _msda(spec, p, C, M, L, P, box):
    _linear(spec, p + ".sampling_offsets", M * L * P * 2, C)
    _linear(spec, p + ".attention_weights", M * L * P, C)
    _linear(spec, p + ".value_proj", C, C)
    _linear(spec, p + ".output_proj", C, C)
    if box:
        _linear(spec, p + ".output_proj_box", C, C)


def _mha(spec, p, C):
    spec[p + ".in_proj_weight"] = (3 * C, C)
    spec[p + ".in_proj_bias"] = (3 * C,)
    _linear(spec, p + ".out_proj", C, C)


def seqformer_r50_state_spec(num_queries=300, num_classes=40, C=256, ffn=1024, M=8, L=4, P=4, enc=6, dec=6, swin=None,
                             depth=50):
    """swin: None for a ResNet (depth 50 / 101 / 152: cfg.MODEL.RESNETS.DEPTH), else a Swin cfg dict (swin_spec.SWIN_L ...)."""
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
        # SeqFormer's MSDeformAttn always owns output_proj_box, encoder copies included (ms_deform_attn.py:61)
        _msda(spec, p + ".self_attn", C, M, L, P, box=True)
        _ln(spec, p + ".norm1", C)
        _linear(spec, p + ".linear1", ffn, C)
        _linear(spec, p + ".linear2", C, ffn)
        _ln(spec, p + ".norm2", C)
    for i in range(dec):
        p = f"{T}decoder.layers.{i}"
        _msda(spec, p + ".cross_attn", C, M, L, P, box=True)
        _ln(spec, p + ".norm1", C)
        _ln(spec, p + ".norm1_box", C)
        _mha(spec, p + ".self_attn", C)
        _ln(spec, p + ".norm2", C)
        _mha(spec, p + ".self_attn_box", C)
        _ln(spec, p + ".norm2_box", C)
        _linear(spec, p + ".linear1", ffn, C)
        _linear(spec, p + ".linear2", C, ffn)
        _ln(spec, p + ".norm3", C)
        _linear(spec, p + ".linear1_box", ffn, C)
        _linear(spec, p + ".linear2_box", C, ffn)
        _ln(spec, p + ".norm3_box", C)
        _linear(spec, p + ".time_attention_weights", 1, C)
    # decoder.bbox_embed aliases detr.bbox_embed (with_box_refine hack, deformable_detr.py:104) -> both key sets exist
    for i in range(dec):
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
    for name, co, ci in (("lay1", C // 4, C), ("lay2", C // 32, C // 4), ("lay3", C, C), ("lay4", C, C), ("dcn", C, C)):
        spec[f"detr.mask_head.{name}.weight"] = (co, ci, 3, 3)
        spec[f"detr.mask_head.{name}.bias"] = (co,)
    return spec
