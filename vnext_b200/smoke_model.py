"""__graft_entry__.smoke() part 2: one tiny SeqFormer clip through the CUDA engine, checked against the CPU oracle."""
import torch


def run():
    from oracle import seqformer_oracle as SO
    from vnext_b200 import _cabi
    from vnext_b200.seqformer.meta_arch import SeqFormer
    from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
    from vnext_b200.synth import fill_state_dict, synthetic_clip
    T, H, W = 2, 96, 128
    sd = fill_state_dict(seqformer_r50_state_spec(), 0)
    frames = synthetic_clip(T, H, W, 0)
    n0 = _cabi.launch_count()
    model = SeqFormer(state_dict=sd, device="cuda", precision="tf32").eval()
    got = model([{"image": frames, "height": H, "width": W, "file_names": ["a", "b"]}])
    with torch.no_grad():
        want = SO.whole_video_postprocess(SO.inference(sd, frames), (H, W))
    assert got["pred_labels"] == want["pred_labels"], "labels differ from the oracle"
    assert torch.allclose(torch.tensor(got["pred_scores"]), torch.tensor(want["pred_scores"]), rtol=2e-2, atol=2e-3)
    agree = sum((a == b).float().mean().item() for a, b in zip(got["pred_masks"], want["pred_masks"])) / max(1, len(want["pred_masks"]))
    assert agree > 0.995, f"mask agreement {agree}"
    print(f"[smoke] SeqFormer-R50 clip {T}x{H}x{W}: {len(got['pred_scores'])} instances, mask agreement {agree:.4f}; "
          f"native kernels launched: {_cabi.launch_count() - n0}")
