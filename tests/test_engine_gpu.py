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
