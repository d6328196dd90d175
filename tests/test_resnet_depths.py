"""ResNet-101 / -152 (projects/IDOL/configs/ytvis19_r101.yaml `DEPTH: 101`; detectron2/modeling/backbone/resnet.py:656-662
`num_blocks_per_stage`): state-dict spec, block discovery from checkpoint keys, oracle and training model on CPU."""
import pytest
import torch

from oracle import seqformer_oracle as SO
from vnext_b200.idol.state_spec import idol_r50_state_spec
from vnext_b200.seqformer import packing
from vnext_b200.seqformer.state_spec import resnet_depth_of, resnet_stages, seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict


def test_stage_layout_per_depth():
    assert [n for _, n, _, _ in resnet_stages(50)] == [3, 4, 6, 3]
    assert [n for _, n, _, _ in resnet_stages(101)] == [3, 4, 23, 3]
    assert [n for _, n, _, _ in resnet_stages(152)] == [3, 8, 36, 3]
    with pytest.raises(NotImplementedError):
        resnet_stages(34)


@pytest.mark.parametrize("spec_fn", [seqformer_r50_state_spec, idol_r50_state_spec])
def test_r101_spec_adds_exactly_the_17_extra_res4_blocks(spec_fn):
    s50, s101 = spec_fn(), spec_fn(depth=101)
    extra = set(s101) - set(s50)
    assert set(s50) <= set(s101) and all(s50[k] == s101[k] for k in s50)
    blocks = {int(k.split("res4.")[1].split(".")[0]) for k in extra}
    assert blocks == set(range(6, 23)) and len(extra) == 17 * 3 * 5           # 3 convs x (weight + 4 FrozenBN buffers)
    assert s101["detr.detr.backbone.0.backbone.res4.22.conv2.weight"] == (256, 256, 3, 3)
    assert resnet_depth_of(s101) == 101 and resnet_depth_of(s50) == 50 and resnet_depth_of(spec_fn(depth=152)) == 152


def test_packing_and_oracle_read_the_block_counts_off_the_checkpoint():
    sd = fill_state_dict(seqformer_r50_state_spec(depth=101), 0)
    assert packing.resnet_blocks(sd) == (("res2", 3), ("res3", 4), ("res4", 23), ("res5", 3))
    packed = packing.pack_resnet50(sd)
    assert [len(b) for _, b in packed["stages"]] == [3, 4, 23, 3]
    x = torch.rand(1, 3, 64, 96)
    f101 = SO.resnet50(sd, x)
    assert f101["res4"].shape == (1, 1024, 4, 6) and f101["res5"].shape == (1, 2048, 2, 3)
    # the oracle really runs the extra blocks: dropping them changes res4
    sd50 = {k: v for k, v in sd.items() if k in seqformer_r50_state_spec()}
    assert not torch.allclose(SO.resnet50(sd50, x)["res4"], f101["res4"])


def test_facade_reads_the_depth_from_the_config():
    from vnext_b200.seqformer.meta_arch import SeqFormer
    cfg = {"MODEL": {"RESNETS": {"DEPTH": 101}}}
    m = SeqFormer(cfg=cfg, device="cpu")
    assert m.resnet_depth == 101 and "detr.detr.backbone.0.backbone.res4.22.conv3.weight" in m.expected_state_spec()
    from vnext_b200.backbone import build_resnet_backbone
    assert build_resnet_backbone(cfg).kind == "resnet50"
    with pytest.raises(NotImplementedError):
        build_resnet_backbone({"MODEL": {"RESNETS": {"DEPTH": 18}}})
