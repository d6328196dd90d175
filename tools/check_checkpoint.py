"""Validate a reference checkpoint against this repo's state-dict specifications before loading it.

    python tools/check_checkpoint.py seqformer_r50.pth --arch seqformer
    python tools/check_checkpoint.py idol_swinl.pth --arch idol --swin L

Reads the `model` dict of a DetectionCheckpointer `.pth` (or a bare state dict), reports missing / unexpected keys and
shape mismatches against `vnext_b200.{seqformer,idol}.state_spec` (SURVEY §8b "Checkpoint/state-dict boundary").  CPU only.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vnext_b200.idol.state_spec import idol_r50_state_spec  # noqa: E402
from vnext_b200.seqformer.state_spec import seqformer_r50_state_spec  # noqa: E402
from vnext_b200.swin_spec import SWIN_L, SWIN_T  # noqa: E402


def check(state_dict, spec):
    """-> (missing, unexpected, mismatched [(key, want, got)])."""
    missing = [k for k in spec if k not in state_dict]
    unexpected = [k for k in state_dict if k not in spec and not k.endswith("num_batches_tracked")]
    bad = []
    for k, shp in spec.items():
        if k in state_dict:
            if shp and shp[0] == "index":
                shp = (shp[1] ** 2, shp[1] ** 2)
            if tuple(state_dict[k].shape) != tuple(shp):
                bad.append((k, tuple(shp), tuple(state_dict[k].shape)))
    return missing, unexpected, bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("checkpoint")
    ap.add_argument("--arch", choices=["seqformer", "idol"], required=True)
    ap.add_argument("--swin", choices=["T", "L"], default=None, help="Swin backbone instead of ResNet-50")
    ap.add_argument("--num-queries", type=int, default=300)
    ap.add_argument("--num-classes", type=int, default=40)
    ap.add_argument("--unsafe-pickle", action="store_true",
                    help="allow arbitrary pickled objects (legacy model-zoo files); only for checkpoints you trust")
    args = ap.parse_args()
    # a validation tool must not execute code from the file it validates: tensors-only unpickling unless asked otherwise
    obj = torch.load(args.checkpoint, map_location="cpu", weights_only=not args.unsafe_pickle)
    sd = obj.get("model", obj) if isinstance(obj, dict) else obj
    swin = {"T": dict(SWIN_T), "L": dict(SWIN_L)}.get(args.swin)
    fn = seqformer_r50_state_spec if args.arch == "seqformer" else idol_r50_state_spec
    missing, unexpected, bad = check(sd, fn(args.num_queries, args.num_classes, swin=swin))
    print(f"{len(sd)} tensors in the checkpoint; missing {len(missing)}, unexpected {len(unexpected)}, shape mismatches {len(bad)}")
    for name, items in (("missing", missing), ("unexpected", unexpected)):
        for k in items[:20]:
            print(f"  {name}: {k}")
    for k, want, got in bad[:20]:
        print(f"  shape: {k}: expected {want}, checkpoint has {got}")
    return 1 if (missing or bad) else 0


if __name__ == "__main__":
    sys.exit(main())
