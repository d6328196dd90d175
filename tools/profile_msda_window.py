"""ncu target: the encoder-shape MSDA forward at the BASELINE 720p geometry (T=5 frames as the batch), window kernel.

    ncu --set full --import-source on -k regex:msda_fwd_window -c 2 -o gpurun_out/win python tools/profile_msda_window.py bf16
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_msda_window_gpu import PYR720, encoder_inputs  # noqa: E402
from vnext_b200 import _cabi, msda  # noqa: E402

msda.WINDOW_MODE = os.environ.get("VNEXT_MSDA_WINDOW", "auto")
dt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 1.6
value, shapes, lsi, loc, attn = encoder_inputs(5, 8, PYR720, 4, 1, sigma, 0.0)
dv, dsh, dl, dloc, dat = [t.cuda() for t in (value.to(dt), shapes, lsi, loc, attn)]
for _ in range(3):
    out = msda.ms_deform_attn_forward(dv, dsh, dl, dloc, dat, 64)
torch.cuda.synchronize()
print(_cabi.last_variant(), out.float().abs().mean().item())
