"""Generate golden vectors for MSDA from the REFERENCE's own CPU function.

Run ONLY in the build container (needs /root/reference):

    python tests/golden/make_msda_golden.py

It loads projects/IDOL/idol/models/ops/functions/ms_deform_attn_func.py by file
path (an empty stub stands in for the compiled `MultiScaleDeformableAttention`
module that file imports at line 18) and evaluates `ms_deform_attn_core_pytorch`
(:42-62) plus its fp64 autograd on seeded inputs.  Inputs are re-creatable from
the seed (oracle.msda_oracle.make_inputs); only small outputs are stored.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.msda_oracle import make_inputs  # noqa: E402

REF = "/root/reference/projects/IDOL/idol/models/ops/functions/ms_deform_attn_func.py"

PYR_480x864 = [(60, 108), (30, 54), (15, 27), (8, 14)]


def load_reference_fn():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    spec = importlib.util.spec_from_file_location("ref_msda_func", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ms_deform_attn_core_pytorch


def ref_fwd_bwd(fn, value, shapes, loc, attn, seed):
    v = value.double().requires_grad_(True)
    lc = loc.double().requires_grad_(True)
    aw = attn.double().requires_grad_(True)
    out = fn(v, shapes, lc, aw)
    g = torch.Generator().manual_seed(seed + 1000)
    gout = torch.rand(out.shape, generator=g, dtype=torch.float64)
    out.backward(gout)
    return out.detach(), gout, v.grad, lc.grad, aw.grad


def main():
    fn = load_reference_fn()
    out = {}

    # (A) the reference test fixture, ops/test.py:21-36 (global seed 3, consumed in this order)
    torch.manual_seed(3)
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out["A_value"], out["A_loc"], out["A_attn"] = value.numpy(), loc.numpy(), attn.numpy()
    out["A_shapes"] = shapes.numpy()
    out["A_out_f64"] = fn(value.double(), shapes, loc.double(), attn.double()).numpy()
    out["A_out_f32"] = fn(value, shapes, loc, attn).numpy()

    # (B) BASELINE config 1: N=2, Lq=100, M=8, D=32, L=4, K=4 on the 480x864 pyramid, in-range
    #     and out-of-range (loc in [-0.2, 1.2]) variants.  Inputs from make_inputs(seed).
    for tag, seed, rng in (("B", 3, (0.0, 1.0)), ("Bo", 4, (-0.2, 1.2))):
        value, shapes, lsi, loc, attn = make_inputs(2, 100, 8, 32, PYR_480x864, 4, seed, loc_range=rng)
        o32 = fn(value, shapes, loc, attn)
        o64, gout, gv, gl, ga = ref_fwd_bwd(fn, value, shapes, loc, attn, seed)
        out[tag + "_out_f32"] = o32.numpy()
        out[tag + "_out_f64"] = o64.numpy()
        out[tag + "_gloc"] = gl.numpy()
        out[tag + "_gattn"] = ga.numpy()
        # grad_value is 17 MB: keep a strided sample and per-level sums
        out[tag + "_gvalue_sample"] = gv.flatten()[::997].numpy()
        out[tag + "_gvalue_sum"] = np.array([gv.sum().item(), gv.abs().sum().item()])

    # (C) backward branch coverage: the channel counts the reference gradchecks (test.py:85) that
    #     are cheap on CPU, on the test fixture's shapes.
    for Dc in (30, 32, 64, 71):
        value, shapes, lsi, loc, attn = make_inputs(1, 2, 2, Dc, [(6, 4), (3, 2)], 2, 100 + Dc,
                                                    dtype=torch.float64, loc_range=(-0.1, 1.1))
        o64, gout, gv, gl, ga = ref_fwd_bwd(fn, value, shapes, loc, attn, 100 + Dc)
        out[f"C{Dc}_out"] = o64.numpy()
        out[f"C{Dc}_gvalue"] = gv.numpy()
        out[f"C{Dc}_gloc"] = gl.numpy()
        out[f"C{Dc}_gattn"] = ga.numpy()

    # (D) ragged / degenerate shapes: L=1, K=8; 1x1 level; single query
    value, shapes, lsi, loc, attn = make_inputs(3, 5, 4, 32, [(1, 1), (2, 7), (5, 3)], 8, 77,
                                                dtype=torch.float64, loc_range=(-0.3, 1.3))
    o64, gout, gv, gl, ga = ref_fwd_bwd(fn, value, shapes, loc, attn, 77)
    out["D_out"], out["D_gvalue"], out["D_gloc"], out["D_gattn"] = (o64.numpy(), gv.numpy(),
                                                                     gl.numpy(), ga.numpy())

    path = os.path.join(HERE, "msda_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; torch", torch.__version__)


if __name__ == "__main__":
    main()
