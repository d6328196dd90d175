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
eng = SeqFormerEngine(bench.make_weights(0), device="cuda", precision="tf32")
frames = bench.make_clip_u8(T, H, W, 0).cuda()
for _ in range(2):
    eng.infer_video_device(frames, (H, W))
with _cabi.KernelTimer() as kt:
    for _ in range(3):
        eng.infer_video_device(frames, (H, W))
agg = kt.summary()
for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:int(os.environ.get("TOPN", "8"))]:
    print(f"{name:34s} calls/step {d['calls'] // 3:3d}  avg {d['us'] / d['calls']:8.1f} us  {d['bytes'] / d['us'] / 1e3:8.1f} GB/s")
