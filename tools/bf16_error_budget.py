"""Where does the bf16 mode's deviation from the fp32 oracle come from?  Runs the engine at a BASELINE geometry with bf16
storage enabled for the backbone only, the encoder branch tensors only, both (the bf16 mode) and neither (= tf32 mode), and
prints max|err| / max|ref| and the 99.99th-percentile |err| / max|ref| of the three model outputs.

    python tools/bf16_error_budget.py [480p|720p]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import seqformer_oracle as SO  # noqa: E402
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec  # noqa: E402
from vnext_b200.synth import fill_state_dict, synthetic_clip  # noqa: E402

H, W = (480, 864) if (len(sys.argv) > 1 and sys.argv[1] == "480p") else (720, 1280)
T = 5
sd = fill_state_dict(seqformer_r50_state_spec(), 0)
frames = synthetic_clip(T, H, W, 0)
torch.set_num_threads(min(32, os.cpu_count() or 1))
with torch.no_grad():
    ref = SO.inference(sd, frames)
from vnext_b200.seqformer.engine import SeqFormerEngine  # noqa: E402

for parts in (sys.argv[2].split(";") if len(sys.argv) > 2 else ("", "backbone", "encoder", "backbone,encoder")):
    os.environ["VNEXT_BF16_PARTS"] = parts
    out = SeqFormerEngine(sd, device="cuda", precision="bf16").inference(torch.stack(frames).cuda())
    line = {}
    for k in ("pred_logits", "pred_boxes", "pred_masks"):
        e = (out[k].float().cpu() - ref[k]).abs().flatten()
        s = ref[k].abs().max().item()
        kth = max(1, int(e.numel() * 0.9999))
        line[k] = (round(e.max().item() / s, 5), round(e.kthvalue(kth)[0].item() / s, 5), round(e.mean().item() / s, 6))
    print(f"{H}x{W} bf16 parts [{parts or 'none'}]: (max, p99.99, mean)|err|/max|ref| {line}", flush=True)
