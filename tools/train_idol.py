"""Minimal IDOL training loop on synthetic data — what `train_net.py` does around the model, with the B200-native
trainer (flat parameter buffers, one gradient all-reduce, fused AdamW).  One process per GPU:

    python tools/train_idol.py --steps 20
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_idol.py --steps 20

Replace `synthetic_clip` / `synthetic_targets` by a YTVIS data loader (the targets are exactly what
`vnext_b200.idol.meta_arch.IDOL.prepare_targets` builds from detectron2 `Instances`) and load a checkpoint's `model`
dict instead of the synthetic weights to fine-tune the released IDOL weights."""
import argparse
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vnext_b200.idol import train as T  # noqa: E402
from vnext_b200.idol.state_spec import idol_r50_state_spec  # noqa: E402
from vnext_b200.synth import fill_state_dict, synthetic_clip, synthetic_targets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--videos", type=int, default=4, help="(key, ref) pairs per GPU per step")
    ap.add_argument("--size", type=int, nargs=2, default=(480, 864))
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"])
    ap.add_argument("--save", default="")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    model = T.IDOLTrainModel(fill_state_dict(idol_r50_state_spec(), 0), device=dev, precision=args.precision).train()
    trainer = T.IDOLTrainer(model)                       # BASE_LR 1e-4, backbone x0.1, wd 1e-4, clip 0.01
    H, W = args.size
    random.seed(rank)
    for step in range(args.steps):
        seed = step * world + rank                        # every rank sees its own videos: data parallel
        frames = torch.stack(synthetic_clip(2 * args.videos, H, W, seed)).to(dev)
        det, ref = synthetic_targets(args.videos, H, W, seed)
        det = [{k: v.to(dev) for k, v in t.items()} for t in det]
        ref = [{k: v.to(dev) for k, v in t.items()} for t in ref]
        grad_norm, total, losses = trainer.step(frames, det, ref)
        if rank == 0:
            print(f"step {step:4d}  loss {float(total):9.4f}  grad-norm {float(grad_norm):8.4f}  "
                  f"ce {float(losses['loss_ce']):.3f}  mask {float(losses['loss_mask']):.3f}  reid {float(losses['loss_reid']):.3f}",
                  flush=True)
    if args.save and rank == 0:
        torch.save({"model": model.reference_state_dict()}, args.save)          # reference key layout
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
