"""One eager (non-graph) SeqFormer step for ncu: kernels of exactly one clip between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --workload 720p
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="720p")
    ap.add_argument("--precision", default="tf32")
    args = ap.parse_args()
    from vnext_b200.seqformer.engine import SeqFormerEngine
    T, H, W = bench.WORKLOADS[args.workload]
    eng = SeqFormerEngine(bench.make_weights(0), device="cuda", precision=args.precision)
    frames = bench.make_clip_u8(T, H, W, 0).cuda().float()
    for _ in range(2):
        eng.infer_video_device(frames, (H, W))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    eng.infer_video_device(frames, (H, W))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
