"""Per-kernel-family time of one eager SeqFormer step (CUDA events via the KernelTimer hooks)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vnext_b200 import _cabi  # noqa: E402
from vnext_b200.seqformer.engine import SeqFormerEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "720p"
T, H, W = bench.WORKLOADS[wl]
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
eng = SeqFormerEngine(bench.make_weights(0), device="cuda", precision=precision)
frames = bench.make_clip_u8(T, H, W, 0).cuda()
for _ in range(2):
    eng.infer_video_device(frames, (H, W))
with _cabi.KernelTimer() as kt:
    for _ in range(3):
        eng.infer_video_device(frames, (H, W))
agg = kt.summary()
tot = sum(d["us"] for d in agg.values()) / 3
print(f"# {wl} {precision}: {tot:.0f} us of C-ABI kernel time per clip (eager, per-launch CUDA events)")
for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:int(os.environ.get("TOPN", "8"))]:
    print(f"{name:40s} calls/step {d['calls'] // 3:3d}  avg {d['us'] / d['calls']:8.1f} us  total {d['us'] / 3:8.1f} us  "
          f"{d['bytes'] / d['us'] / 1e3:8.1f} GB/s  {d['flops'] / d['us'] / 1e6:7.1f} TF/s")
