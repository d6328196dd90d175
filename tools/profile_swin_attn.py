"""One Swin-L stage-3 shaped window-attention call (SIMT fp32, tcgen05 TF32, mma.sync bf16 kernels) for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vnext_b200 import fused  # noqa: E402

B, H, W, heads, ws = 5, 46, 80, 24, 12
C = heads * 32
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B, H, W, 3 * C, generator=g).cuda()
bias = torch.randn(3 * C, generator=g).cuda()
table = (torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5).cuda()
for tc in (False, True):
    for _ in range(2):
        fused.swin_window_attention(qkv, bias, table, heads, ws, 6, tensor_cores=tc)
q16 = qkv.bfloat16()
for _ in range(2):
    fused.swin_window_attention(q16, bias, table, heads, ws, 6, tensor_cores=True)     # bf16 storage: the mma.sync kernel
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fused.swin_window_attention(qkv, bias, table, heads, ws, 6, tensor_cores=False)
fused.swin_window_attention(qkv, bias, table, heads, ws, 6, tensor_cores=True)
fused.swin_window_attention(q16, bias, table, heads, ws, 6, tensor_cores=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
