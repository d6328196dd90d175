"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Python face of the CPU oracle for multi-scale deformable attention (MSDA).

Two independent restatements of the reference algorithm live here:

* ``forward_c`` / ``backward_c`` — ctypes calls into ``msda_oracle.c`` (scalar C,
  follows the reference CUDA kernel's arithmetic:
  projects/IDOL/idol/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:33-84,87-159,237-299).
* ``forward_torch`` — the reference's *CPU* path restated with ``F.grid_sample``
  (follows projects/IDOL/idol/models/ops/functions/ms_deform_attn_func.py:42-62:
  grids = 2*loc-1, bilinear / zeros padding / align_corners=False per level,
  weighted sum over L*P).  ``backward_torch`` is autograd through it
  (the reference has no closed-form CPU backward; its test uses gradcheck,
  projects/IDOL/idol/models/ops/test.py:63-78).

Pinned by tests/test_oracle_golden.py against golden vectors generated from the
reference's own function (tests/golden/make_msda_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libmsda_oracle.so")
_lib = None


def build(force=False):
    """Compile msda_oracle.c with gcc into oracle/_build/ (git-ignored, travels with gpurun)."""
    src = os.path.join(_HERE, "msda_oracle.c")
    if (not force) and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(_BUILD, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _np(t, dtype):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(value, shapes, loc):
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    return N, S, M, D, L, Lq, P


def forward_c(value, shapes, lsi, loc, attn, dtype=np.float32):
    lib = _load()
    v, lc, aw = _np(value, dtype), _np(loc, dtype), _np(attn, dtype)
    sh, ls = _np(shapes, np.int64), _np(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(v, sh, lc)
    out = np.empty((N, Lq, M * D), dtype=dtype)
    fn = lib.msda_oracle_forward_f32 if dtype == np.float32 else lib.msda_oracle_forward_f64
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lc), _ptr(aw), N, S, M, D, L, Lq, P, _ptr(out))
    return torch.from_numpy(out)


def backward_c(value, shapes, lsi, loc, attn, grad_out, dtype=np.float32):
    lib = _load()
    v, lc, aw, go = _np(value, dtype), _np(loc, dtype), _np(attn, dtype), _np(grad_out, dtype)
    sh, ls = _np(shapes, np.int64), _np(lsi, np.int64)
    N, S, M, D, L, Lq, P = _dims(v, sh, lc)
    gv, gl, ga = np.empty_like(v), np.empty_like(lc), np.empty_like(aw)
    fn = lib.msda_oracle_backward_f32 if dtype == np.float32 else lib.msda_oracle_backward_f64
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lc), _ptr(aw), _ptr(go), N, S, M, D, L, Lq, P,
       _ptr(gv), _ptr(gl), _ptr(ga))
    return torch.from_numpy(gv), torch.from_numpy(gl), torch.from_numpy(ga)


def forward_torch(value, shapes, loc, attn):
    """grid_sample restatement of the reference CPU path (ms_deform_attn_func.py:42-62)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    hw = [(int(h), int(w)) for h, w in shapes.tolist()]
    grids = loc * 2 - 1
    start = 0
    per_level = []
    for lvl, (H, W) in enumerate(hw):
        # [N, H*W, M, D] -> [N*M, D, H, W]
        feat = value[:, start:start + H * W].permute(0, 2, 3, 1).reshape(N * M, D, H, W)
        start += H * W
        # [N, Lq, M, P, 2] -> [N*M, Lq, P, 2]
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq, P, 2)
        per_level.append(F.grid_sample(feat, g, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))        # [N*M, D, Lq, P]
    sampled = torch.stack(per_level, dim=-2).flatten(-2)               # [N*M, D, Lq, L*P]
    w = attn.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L * P)
    out = (sampled * w).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


def backward_torch(value, shapes, loc, attn, grad_out):
    v = value.detach().clone().requires_grad_(True)
    lc = loc.detach().clone().requires_grad_(True)
    aw = attn.detach().clone().requires_grad_(True)
    out = forward_torch(v, shapes, lc, aw)
    out.backward(grad_out)
    return v.grad, lc.grad, aw.grad


def make_inputs(*args, **kwargs):
    """Input recipe of the reference test (ops/test.py:33-36); shared with bench.py via vnext_b200.synth."""
    from vnext_b200.synth import make_msda_inputs
    return make_msda_inputs(*args, **kwargs)
