"""For ncu: one launch each of (1) the encoder's offsets|logits projection with the sampler front-end in its epilogue
(vnext_linear_msda_front), (2) the cache-blocked MSDA sampler reading its output (pre-located mode), (3) the six RLE kernels
on 50 masks of 720x1280 — at the 720p T=5 geometry, between cudaProfilerStart/Stop.

    ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/front_rle -f \
        python tools/profile_front_rle.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vnext_b200 import msda  # noqa: E402
from vnext_b200.linear import linear_msda_front  # noqa: E402
from vnext_b200.rle import RleEncoder  # noqa: E402

PYR = [(92, 160), (46, 80), (23, 40), (12, 20)]


def main():
    T, S = 5, sum(h * w for h, w in PYR)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(T, S, 256, generator=g).cuda()
    w = (torch.randn(384, 256, generator=g) / 16).cuda()
    w[:256] *= 2.0
    b2 = (torch.randn(S, 384, generator=g) * 0.5).cuda()
    refs = []
    for h, wd in PYR:
        ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(wd) + 0.5) / wd, indexing="ij")
        refs.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    ref = torch.cat(refs, 0)[:, None, :].expand(S, 4, 2).contiguous().cuda()
    value = (torch.rand(T, S, 8, 32, generator=g) * 0.01).cuda()
    shapes = torch.tensor(PYR, dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    geom = msda.host_geometry(shapes, lsi)
    f = torch.randn((50, 1, 20, 34), generator=g)
    masks = (torch.nn.functional.interpolate(f, size=(720, 1280), mode="bilinear", align_corners=False)[:, 0] > 0.3).cuda()
    enc = RleEncoder(50, 720, 1280, "cuda")
    for _ in range(2):
        la = linear_msda_front(x, w, b2, ref, PYR)
        msda.ms_deform_attn_located(value, la, 4, 4, geom)
        enc.encode(masks)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    la = linear_msda_front(x, w, b2, ref, PYR)
    msda.ms_deform_attn_located(value, la, 4, 4, geom)
    enc.encode(masks)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
