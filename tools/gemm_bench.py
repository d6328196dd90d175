"""Time the tcgen05 linear kernel on the shapes the SeqFormer engine uses (CUDA events, L2 flushed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vnext_b200.linear import linear  # noqa: E402

SHAPES = [  # (M, N, K, relu, residual, ln, tag)
    (97800, 1024, 256, True, False, False, "enc FFN1"),
    (97800, 256, 1024, False, True, True, "enc FFN2+res+LN"),
    (97800, 256, 256, False, True, True, "enc out_proj+res+LN"),
    (97800, 256, 256, False, False, False, "enc value_proj"),
    (97800, 384, 256, False, False, False, "enc offsets|logits"),
    (294400, 256, 64, True, True, False, "res2 conv3+res+relu"),
    (294400, 64, 256, True, False, False, "res2 conv1+relu"),
    (73600, 512, 128, True, True, False, "res3 conv3+res+relu"),
    (18400, 1024, 256, True, True, False, "res4 conv3+res+relu"),
    (4600, 2048, 512, True, True, False, "res5 conv3+res+relu"),
    (1500, 256, 256, False, False, False, "decoder small"),
]


def main():
    dbg = int(os.environ.get("GEMM_DEBUG", "0"), 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for M, N, K, relu, res, ln, tag in SHAPES:
        x = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if res else None
        g = torch.ones(N, device="cuda") if ln else None
        be = torch.zeros(N, device="cuda") if ln else None
        out = torch.empty(M, N, device="cuda")
        fn = lambda: linear(x, w, b, relu=relu, residual=r, ln_weight=g, ln_bias=be, out=out, _debug_flags=dbg)
        for _ in range(5):
            fn()
        ts = []
        for _ in range(20):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); e.record(); e.synchronize()
            ts.append(a.elapsed_time(e) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        byt = 4.0 * (M * K + N * K + M * N * (2 if res else 1))
        print(f"{tag:24s} M={M:6d} N={N:4d} K={K:4d}  {us:8.1f} us  {byt / us / 1e6:6.2f} TB/s  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
