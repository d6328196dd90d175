"""CPU tests of the SeqFormer engine's HOST logic (weight packing, geometry plan, frame batching,
fused projections, top-k mask restriction) against the CPU oracle.  The deformable-attention sampler
is injected from the oracle here because the product op is CUDA-only; the `-m gpu` tests run the same
comparisons with the real kernels."""
import pytest
import torch

from oracle import msda_oracle, seqformer_oracle as SO
from vnext_b200.seqformer.engine import SeqFormerEngine
from vnext_b200.seqformer.packing import fold_conv_bn
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip

CASES = {"nopad": (3, 96, 128, 0), "pad": (2, 80, 112, 1)}


@pytest.fixture(scope="module")
def weights():
    return {seed: fill_state_dict(seqformer_r50_state_spec(), seed) for seed in (0, 1)}


def test_bn_folding_is_exact(weights):
    sd = weights[0]
    p = "detr.detr.backbone.0.backbone.res3.0.conv2"
    x = torch.randn(2, 128, 9, 11, generator=torch.Generator().manual_seed(0))
    ref = SO._conv_bn(sd, p, x, stride=2, padding=1, relu=False)
    w, b = fold_conv_bn(sd, p)
    assert torch.allclose(torch.nn.functional.conv2d(x, w, b, stride=2, padding=1), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_engine_matches_oracle_on_cpu(weights, name):
    T, H, W, seed = CASES[name]
    sd = weights[seed]
    frames = synthetic_clip(T, H, W, seed)
    eng = SeqFormerEngine(sd, device="cpu", msda_impl=msda_oracle.forward_torch)
    with torch.no_grad():
        ref = SO.inference(sd, frames, return_intermediates=True)
        out = eng.inference(torch.stack(frames))
    plan = eng.plan(T, H, W)
    assert plan.any_pad == (name == "pad") and plan.S == ref["memory"].shape[2]
    assert torch.allclose(plan.pos_flat, ref["pos_flat"][0, :1], rtol=1e-5, atol=1e-5)
    assert torch.allclose(plan.valid_ratios, ref["valid_ratios"])
    assert torch.allclose(out["memory"], ref["memory"][0], rtol=1e-3, atol=1e-3)
    assert torch.allclose(out["pred_logits"], ref["pred_logits"], rtol=1e-3, atol=1e-3)
    assert torch.allclose(out["pred_boxes"], ref["pred_boxes"], rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["pred_masks"], ref["pred_masks"], rtol=1e-3, atol=5e-3)

    # video-level output: top-k restricted fast path == oracle post-processing of the full output
    want = SO.whole_video_postprocess(ref, (H, W))
    scores, masks = eng.infer_video_device(torch.stack(frames), (H, W))
    got = eng.finish_video(scores, masks, (H, W))
    assert got["pred_labels"] == want["pred_labels"]
    assert torch.allclose(torch.tensor(got["pred_scores"]), torch.tensor(want["pred_scores"]), rtol=1e-3, atol=1e-4)
    agree = sum((a == b).float().mean().item() for a, b in zip(got["pred_masks"], want["pred_masks"]))
    assert agree / len(want["pred_masks"]) > 0.999


@pytest.mark.parametrize("name", list(CASES))
def test_idol_engine_matches_oracle_on_cpu(name):
    from oracle import idol_oracle as IO
    from vnext_b200.idol.engine import IDOLEngine
    from vnext_b200.idol.state_spec import idol_r50_state_spec
    T, H, W, seed = CASES[name]
    sd = fill_state_dict(idol_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    eng = IDOLEngine(sd, device="cpu", msda_impl=msda_oracle.forward_torch)
    with torch.no_grad():
        ref = IO.inference_forward(sd, frames)
        out = eng.inference_forward(torch.stack(frames))
    for k, tol in (("pred_logits", 1e-3), ("pred_boxes", 1e-4), ("pred_inst_embed", 1e-3), ("pred_masks", 5e-3)):
        assert torch.allclose(out[k], ref[k], rtol=1e-3, atol=tol), k


def test_swin_backbone_host_logic_matches_oracle():
    """The Swin backbone engine's restructured math (no pad/roll/partition copies; bias-synthesised padded tokens)
    against the oracle, on CPU (plain-torch formulation of the fused window-attention kernel's contract)."""
    from oracle import swin_oracle as WO
    from vnext_b200.swin import SwinBackbone
    from vnext_b200.swin_spec import swin_state_spec

    class Host:              # the two things the backbone needs from its owning engine
        native = False
        lin = staticmethod(lambda x, wb, relu=False, residual=None, ln=None, gelu=False: (
            lambda y: torch.nn.functional.gelu(y) if gelu else y)(
            torch.nn.functional.linear(x, wb[0], wb[1]) + (residual if residual is not None else 0)))
    for cfg, B, H, W, seed in ((dict(embed_dim=96, depths=(2, 2, 2, 2), num_heads=(3, 6, 12, 24), window_size=12), 2, 160, 224, 0),
                               (dict(embed_dim=64, depths=(2, 2, 2, 2), num_heads=(2, 4, 8, 16), window_size=7), 1, 96, 128, 1)):
        sd = fill_state_dict(swin_state_spec("bb.", **cfg), seed)
        x = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed + 50))
        with torch.no_grad():
            want = WO.swin(sd, x, "bb.", **cfg)
            got = SwinBackbone(sd, "bb.", cfg, torch.device("cpu"), Host).forward(x)
        for k in want:
            assert torch.allclose(got[k].permute(0, 3, 1, 2), want[k], rtol=1e-3, atol=2e-4), k
