"""GPU parity of the SeqFormer engine (CUDA kernels + CUDA graph) against the CPU oracle and the golden
vectors from the reference's own modules.  Tolerance (north_star): 1e-3 for the fp32 path."""
import os

import numpy as np
import pytest
import torch

from oracle import seqformer_oracle as SO
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
from vnext_b200.synth import fill_state_dict, synthetic_clip

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"nopad": (3, 96, 128, 0), "pad": (2, 80, 112, 1)}


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "seqformer_golden.npz"))


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-6)).item()


@pytest.mark.parametrize("name", list(CASES))
def test_engine_fp32_vs_oracle_and_golden(golden, name):
    from vnext_b200 import _cabi
    from vnext_b200.seqformer.engine import SeqFormerEngine
    T, H, W, seed = CASES[name]
    sd = fill_state_dict(seqformer_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    eng = SeqFormerEngine(sd, device="cuda", precision="fp32")
    n0 = _cabi.launch_count()
    out = eng.inference(torch.stack(frames).cuda())
    assert _cabi.launch_count() - n0 >= 12 and _cabi.last_variant().startswith("fwd_"), "native kernels must run"
    with torch.no_grad():
        ref = SO.inference(sd, frames, return_intermediates=True)
    # fp32 path: 1e-3 relative to the tensor's scale
    assert _rel(out["memory"].cpu(), ref["memory"][0]) < 1e-3
    assert _rel(out["pred_logits"].cpu(), ref["pred_logits"]) < 1e-3
    assert _rel(out["pred_boxes"].cpu(), ref["pred_boxes"]) < 1e-3
    assert _rel(out["pred_masks"].cpu(), ref["pred_masks"]) < 1e-3
    g = lambda k: torch.from_numpy(golden[f"{name}_{k}"])
    assert torch.allclose(out["pred_logits"].cpu(), g("pred_logits"), rtol=1e-3, atol=1e-3)
    assert torch.allclose(out["pred_boxes"].cpu(), g("pred_boxes"), rtol=1e-3, atol=1e-3)
    assert _rel(out["pred_masks"].cpu()[:, ::10], g("pred_masks")) < 1e-3


def test_graph_runner_and_meta_arch_output():
    from vnext_b200.seqformer.meta_arch import SeqFormer
    T, H, W, seed = 3, 96, 128, 0
    sd = fill_state_dict(seqformer_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    model = SeqFormer(state_dict=sd, device="cuda", precision="fp32").eval()
    batched = [{"image": frames, "height": H, "width": W, "file_names": ["a", "b", "c"]}]
    got = model(batched)
    got2 = model(batched)                          # second call = pure CUDA-graph replay
    with torch.no_grad():
        want = SO.whole_video_postprocess(SO.inference(sd, frames), (H, W))
    assert got["image_size"] == (H, W)
    assert got["pred_labels"] == want["pred_labels"] == got2["pred_labels"]
    assert torch.allclose(torch.tensor(got["pred_scores"]), torch.tensor(want["pred_scores"]), rtol=1e-3, atol=1e-4)
    agree = np.mean([(a == b).float().mean().item() for a, b in zip(got["pred_masks"], want["pred_masks"])])
    assert agree > 0.999
    assert all(torch.equal(a, b) for a, b in zip(got["pred_masks"], got2["pred_masks"]))
    # results are fresh tensors: a later clip must not overwrite an earlier clip's masks (the reference returns new ones)
    snapshot = [m.clone() for m in got["pred_masks"]]
    other = model([{"image": synthetic_clip(T, H, W, seed + 7), "height": H, "width": W, "file_names": ["a", "b", "c"]}])
    assert all(torch.equal(a, b) for a, b in zip(got["pred_masks"], snapshot))
    assert other["pred_masks"][0].data_ptr() != got["pred_masks"][0].data_ptr()
    # stream(): same results, in order, with two clips in flight
    b2 = [{"image": synthetic_clip(T, H, W, seed + 7), "height": H, "width": W, "file_names": ["a", "b", "c"]}]
    outs = list(model.stream([batched, b2, batched]))
    assert len(outs) == 3 and outs[0]["pred_labels"] == got["pred_labels"] == outs[2]["pred_labels"]
    assert all(torch.equal(a, b) for a, b in zip(outs[0]["pred_masks"], got["pred_masks"]))
    assert all(torch.equal(a, b) for a, b in zip(outs[1]["pred_masks"], other["pred_masks"]))
    with pytest.raises(KeyError):
        SeqFormer(state_dict={k: v for k, v in sd.items() if "query_embed" not in k}, device="cuda")


def test_tf32_mode_error_is_reported():
    """TF32 tensor-core mode (the bench default, = PyTorch's default conv precision on the reference):
    measure and bound its deviation from the fp32 oracle."""
    from vnext_b200.seqformer.engine import SeqFormerEngine
    T, H, W, seed = 3, 96, 128, 0
    sd = fill_state_dict(seqformer_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    out = SeqFormerEngine(sd, device="cuda", precision="tf32").inference(torch.stack(frames).cuda())
    with torch.no_grad():
        ref = SO.inference(sd, frames)
    errs = {k: _rel(out[k].cpu(), ref[k]) for k in ("pred_logits", "pred_boxes", "pred_masks")}
    print("tf32 relative errors:", errs)
    assert max(errs.values()) < 2e-2


def test_bf16_mode_error_is_reported():
    """bf16 mode (bf16 storage + bf16 MMAs in the backbone, the input projection and the encoder's value / sampled / FFN-hidden
    tensors; fp32 residual stream; BASELINE config 4's precision): within the north star's 1e-2 of the fp32 oracle, relative
    to each output's scale."""
    from vnext_b200.seqformer.engine import SeqFormerEngine
    T, H, W, seed = 3, 96, 128, 0
    sd = fill_state_dict(seqformer_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    out = SeqFormerEngine(sd, device="cuda", precision="bf16").inference(torch.stack(frames).cuda())
    assert out["memory"].dtype == torch.float32            # the encoder's residual stream stays fp32
    with torch.no_grad():
        ref = SO.inference(sd, frames, return_intermediates=True)
    errs = {k: _rel(out[k].float().cpu(), ref[k]) for k in ("pred_logits", "pred_boxes", "pred_masks")}
    mem_err = _rel(out["memory"].float().cpu(), ref["memory"][0])
    print("bf16 relative errors:", errs, "encoder memory (intermediate):", mem_err)
    # class logits and boxes: the north star's bf16 bound.  Mask logits (a per-instance dynamic conv on the encoder memory)
    # amplify the encoder's bf16 rounding: 0.8-1.7e-2 of their scale depending on the clip (max over ~10^7 values); see
    # test_full_size_parity_of_the_benchmarked_modes and tools/bf16_error_budget.py
    assert errs["pred_logits"] < 1e-2 and errs["pred_boxes"] < 1e-2 and errs["pred_masks"] < 2e-2
    assert mem_err < 1.5e-2


@pytest.mark.parametrize("geom", ["480x864", "720x1280"])
def test_full_size_parity_of_the_benchmarked_modes(geom):
    """The modes bench.py times, at the BASELINE geometries (T=5), against the CPU oracle on the same clip.  tf32 (the bench
    default): every dense output within 1e-2 of its tensor's scale (measured 6.2-6.6e-3 on the mask logits, 1.8e-3 / 0.9e-3 on
    class logits / boxes).  bf16 (reported next to it): class logits and boxes within 1e-2 (2.5e-3 / 1.9e-3); the mask logits
    land at 1.0-1.7e-2 (max; mean 1.2-1.8e-3) and are bounded at 2e-2 here — the north star's 1e-2 is NOT met on that output,
    which is why bf16 is not the default.  Both: identical top-10 instances, binary-mask agreement > 0.99."""
    from vnext_b200.seqformer.engine import SeqFormerEngine
    H, W = (480, 864) if geom == "480x864" else (720, 1280)
    T = 5
    sd = fill_state_dict(seqformer_r50_state_spec(), 0)
    frames = synthetic_clip(T, H, W, 0)
    with torch.no_grad():
        ref = SO.inference(sd, frames)
    top_ref = ref["pred_logits"][0].sigmoid().max(1)[0].topk(10)[1]
    for precision in ("tf32", "bf16"):
        out = SeqFormerEngine(sd, device="cuda", precision=precision).inference(torch.stack(frames).cuda())
        errs = {k: _rel(out[k].float().cpu(), ref[k]) for k in ("pred_logits", "pred_boxes", "pred_masks")}
        top = out["pred_logits"][0].sigmoid().max(1)[0].topk(10)[1].cpu()
        same = len(set(top.tolist()) & set(top_ref.tolist()))
        agree = ((out["pred_masks"][0, top_ref.cuda()].float().cpu() > 0) == (ref["pred_masks"][0, top_ref] > 0)).float().mean().item()
        print(f"{geom} {precision}: rel errors {errs}; top-10 overlap {same}/10; mask agreement {agree:.5f}")
        if precision == "tf32":
            assert max(errs.values()) < 1e-2, (precision, errs)
        else:
            assert errs["pred_logits"] < 1e-2 and errs["pred_boxes"] < 1e-2 and errs["pred_masks"] < 2e-2, (precision, errs)
        assert same >= 9 and agree > 0.99


def test_clip_matching_inference_long_video():
    """CLIP_MATCHING: True — sliding clips through the engine, stitched on the device; against the CPU oracle run on the
    same windows + the (reference-pinned) clip tracker."""
    from vnext_b200.seqformer.clip_matching import (ClipTracks, clip_matching_postprocess, clip_windows,
                                                    select_clip_instances)
    from vnext_b200.seqformer.meta_arch import SeqFormer
    L, H, W = 5, 96, 128
    sd = fill_state_dict(seqformer_r50_state_spec(), 0)
    frames = synthetic_clip(L, H, W, 4)
    model = SeqFormer(state_dict=sd, device="cuda", precision="fp32").eval()
    model.clip_matching, model.clip_length, model.clip_stride = True, 3, 1
    got = model([{"image": frames, "height": H, "width": W, "file_names": [str(i) for i in range(L)]}])
    tracks = None
    with torch.no_grad():
        for idx in clip_windows(L, 3, 1):
            o = SO.inference(sd, [frames[i] for i in idx])
            cls, ml = select_clip_instances(o["pred_logits"], o["pred_masks"])
            if tracks is None:
                tracks = ClipTracks(3, L, 40, tuple(ml.shape[-2:]), "cpu")
            tracks.update(idx, cls, ml)
        want = clip_matching_postprocess(*tracks.get_result(), (H, W), (H, W))
    assert got["pred_labels"] == want["pred_labels"] and got["image_size"] == (H, W)
    assert torch.allclose(torch.tensor(got["pred_scores"]), torch.tensor(want["pred_scores"]), rtol=1e-3, atol=1e-4)
    agree = np.mean([(a == b).float().mean().item() for a, b in zip(got["pred_masks"], want["pred_masks"])])
    assert agree > 0.999 and got["pred_masks"][0].shape == (L, H, W)


@pytest.mark.parametrize("name", ["stride1", "stride3", "short"])
def test_clip_tracker_on_device_matches_reference(name):
    from test_seqformer_clipmatch_golden import run_case
    run_case(name, "cuda")


def test_adjacent_pinned_uint8_frames_are_copied_straight_from_the_callers_memory():
    """A clip held as one pinned uint8 buffer: one H2D copy from it (no staging copy); same result as separate float frames."""
    from vnext_b200.seqformer.meta_arch import SeqFormer
    T, H, W, seed = 3, 96, 128, 0
    sd = fill_state_dict(seqformer_r50_state_spec(), seed)
    frames = synthetic_clip(T, H, W, seed)
    block = torch.stack(frames).to(torch.uint8).pin_memory()
    model = SeqFormer(state_dict=sd, device="cuda", precision="fp32").eval()
    run = model.runner(T, H, W, (H, W))
    views = [f for f in block]
    assert run._pinned_block(views) is not None and run._pinned_block(views).data_ptr() == block.data_ptr()
    assert run._pinned_block([f.clone() for f in block]) is None                  # separate allocations
    assert run._pinned_block(views[::-1]) is None                                  # not in order
    assert run._pinned_block([f.float() for f in block]) is None
    want = model([{"image": frames, "height": H, "width": W}])
    got = model([{"image": views, "height": H, "width": W}])
    assert got["pred_scores"] == want["pred_scores"] and got["pred_labels"] == want["pred_labels"]
    assert all(torch.equal(a, b) for a, b in zip(got["pred_masks"], want["pred_masks"]))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("tf32", 1e-2)])
def test_resnet101_backbone_engine_vs_oracle(precision, tol):
    """cfg.MODEL.RESNETS.DEPTH: 101 (projects/IDOL/configs/ytvis19_r101.yaml): 23 res4 blocks found from the checkpoint keys."""
    from vnext_b200.seqformer.engine import SeqFormerEngine
    T, H, W, seed = 2, 96, 128, 3
    sd = fill_state_dict(seqformer_r50_state_spec(depth=101), seed)
    frames = synthetic_clip(T, H, W, seed)
    out = SeqFormerEngine(sd, device="cuda", precision=precision).inference(torch.stack(frames).cuda())
    with torch.no_grad():
        ref = SO.inference(sd, frames)
    for k in ("pred_logits", "pred_boxes", "pred_masks"):
        assert _rel(out[k].float().cpu(), ref[k]) < tol, k
