"""__graft_entry__.smoke() part 2: one tiny SeqFormer clip through the CUDA engine, checked against the CPU oracle."""
import torch


def run():
    from oracle import seqformer_oracle as SO
    from vnext_b200 import _cabi
    from vnext_b200.seqformer.engine import SeqFormerEngine
    from vnext_b200.seqformer.meta_arch import SeqFormer
    from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec
    from vnext_b200.synth import fill_state_dict, synthetic_clip
    T, H, W = 2, 96, 128
    sd = fill_state_dict(seqformer_r50_state_spec(), 0)
    frames = synthetic_clip(T, H, W, 0)
    with torch.no_grad():
        ref = SO.inference(sd, frames)
        want = SO.whole_video_postprocess(ref, (H, W))
    # (1) strict fp32 mode through the reference-facing API (CUDA graph): identical instances, matching masks
    n0 = _cabi.launch_count()
    model = SeqFormer(state_dict=sd, device="cuda", precision="fp32").eval()
    got = model([{"image": frames, "height": H, "width": W, "file_names": ["a", "b"]}])
    assert got["pred_labels"] == want["pred_labels"], "labels differ from the oracle"
    assert torch.allclose(torch.tensor(got["pred_scores"]), torch.tensor(want["pred_scores"]), rtol=1e-3, atol=1e-4)
    agree = sum((a == b).float().mean().item() for a, b in zip(got["pred_masks"], want["pred_masks"])) / max(1, len(want["pred_masks"]))
    assert agree > 0.999, f"mask agreement {agree}"
    # (2) tf32 mode (tcgen05 GEMMs, the bench default): dense outputs within the documented tolerance
    out = SeqFormerEngine(sd, device="cuda", precision="tf32").inference(torch.stack(frames).cuda())
    errs = {k: ((out[k].cpu() - ref[k]).abs().max() / ref[k].abs().max()).item() for k in ("pred_logits", "pred_boxes", "pred_masks")}
    assert max(errs.values()) < 3e-2, errs
    print(f"[smoke] SeqFormer-R50 clip {T}x{H}x{W}: {len(got['pred_scores'])} instances, mask agreement {agree:.4f} (fp32 mode); "
          f"tf32 rel. errors {errs}; native kernels launched: {_cabi.launch_count() - n0}")
