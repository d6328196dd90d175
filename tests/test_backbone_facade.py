"""BACKBONE_REGISTRY drop-ins (vnext_b200/backbone.py): interface on CPU, values on the GPU against the oracles."""
import pytest
import torch

from vnext_b200.backbone import B200Backbone, D2SwinTransformer, build_resnet_backbone
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict

BB = "detr.detr.backbone.0.backbone."
SMALL = dict(embed_dim=32, depths=(2, 2, 2, 2), num_heads=(1, 2, 4, 8), window_size=4)


def test_interface_matches_detectron2_backbone():
    r = build_resnet_backbone()
    shapes = r.output_shape()
    assert list(shapes) == ["res2", "res3", "res4", "res5"] and r.size_divisibility == 0
    assert [(s.channels, s.stride) for s in shapes.values()] == [(256, 4), (512, 8), (1024, 16), (2048, 32)]
    cfg = {"MODEL": {"SWIN": {"EMBED_DIM": 192, "DEPTHS": [2, 2, 18, 2], "NUM_HEADS": [6, 12, 24, 48], "WINDOW_SIZE": 12,
                              "OUT_FEATURES": ["res3", "res4", "res5"]}}}
    s = D2SwinTransformer(cfg)
    assert s.size_divisibility == 32 and {k: v.channels for k, v in s.output_shape().items()} == \
        {"res3": 384, "res4": 768, "res5": 1536}
    with pytest.raises(RuntimeError, match="no weights loaded"):
        r(torch.zeros(1, 3, 32, 32))
    assert build_resnet_backbone({"MODEL": {"RESNETS": {"DEPTH": 101}}}).kind == "resnet50"      # bottleneck family: 50 / 101 / 152
    with pytest.raises(NotImplementedError):
        build_resnet_backbone({"MODEL": {"RESNETS": {"DEPTH": 34}}})


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_backbones_match_oracle_on_gpu():
    from oracle import seqformer_oracle as SO, swin_oracle as WO
    from vnext_b200.swin_spec import swin_state_spec
    sd = fill_state_dict(seqformer_r50_state_spec(), 0)
    d2 = {k[len(BB):]: v for k, v in sd.items() if k.startswith(BB)}             # what a D2 backbone checkpoint holds
    x = torch.randn(2, 3, 96, 128, generator=torch.Generator().manual_seed(0))
    got = build_resnet_backbone(precision="fp32").load_reference_state_dict(d2)(x.cuda())
    with torch.no_grad():
        want = SO.resnet50(sd, x)
    for k in ("res2", "res3", "res4", "res5"):
        assert got[k].shape == want[k].shape
        assert ((got[k].cpu() - want[k]).abs().max() / want[k].abs().max()).item() < 1e-3, k
    ssd = fill_state_dict(swin_state_spec("", **SMALL), 0)
    sw = B200Backbone("swin", SMALL, precision="fp32").load_reference_state_dict(ssd)(x.cuda())
    with torch.no_grad():
        wsw = WO.swin(ssd, x, "", **SMALL)
    for k in ("res3", "res5"):
        assert ((sw[k].cpu() - wsw[k]).abs().max() / wsw[k].abs().max()).item() < 1e-3, k
