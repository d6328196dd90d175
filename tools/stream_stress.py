"""Stress the ClipRunner / SeqFormer.stream() path: N rounds of forward() vs stream() on the same clips, counting the
mask pixels and score bits that differ.  Used to root-cause the round-1 driver failure (tests/test_engine_gpu.py)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vnext_b200.seqformer.meta_arch import SeqFormer          # noqa: E402
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec  # noqa: E402
from vnext_b200.synth import fill_state_dict, synthetic_clip  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--fresh", type=int, default=1, help="rebuild the model every round (exercises runner construction)")
    a = ap.parse_args()
    T, H, W = 3, 96, 128
    sd = fill_state_dict(seqformer_r50_state_spec(), 0)
    mk = lambda s: [{"image": synthetic_clip(T, H, W, s), "height": H, "width": W, "file_names": ["a", "b", "c"]}]
    b0, b1 = mk(0), mk(7)
    stats = []
    model = None
    for r in range(a.rounds):
        if model is None or a.fresh:
            model = SeqFormer(state_dict=sd, device="cuda", precision=a.precision).eval()
        g0, g1 = model(b0), model(b1)
        outs = list(model.stream([b0, b1, b0, b1]))
        rec = {"round": r}
        for i, (o, g) in enumerate(zip(outs, (g0, g1, g0, g1))):
            px = sum(int((x != y).sum()) for x, y in zip(o["pred_masks"], g["pred_masks"])) if len(o["pred_masks"]) == len(g["pred_masks"]) else -1
            sc = max([abs(x - y) for x, y in zip(o["pred_scores"], g["pred_scores"])] or [0.0]) if len(o["pred_scores"]) == len(g["pred_scores"]) else -1
            rec[f"clip{i}"] = {"mask_px_diff": px, "score_maxdiff": sc, "labels_equal": o["pred_labels"] == g["pred_labels"]}
        stats.append(rec)
        print(json.dumps(rec), flush=True)
    bad = sum(1 for r in stats for k, v in r.items() if k.startswith("clip") and (v["mask_px_diff"] != 0 or v["score_maxdiff"] != 0))
    print(json.dumps({"precision": a.precision, "rounds": a.rounds, "clips_with_any_difference": bad}))


if __name__ == "__main__":
    main()
