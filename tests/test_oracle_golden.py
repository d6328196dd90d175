"""Pin the CPU oracle (oracle/msda_oracle.{c,py}) against golden vectors produced by the
reference's own ms_deform_attn_core_pytorch (tests/golden/make_msda_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

PYR = [(60, 108), (30, 54), (15, 27), (8, 14)]


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_reference_fixture_A(golden_msda):
    g = golden_msda
    shapes = _t(g["A_shapes"])
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    v, lc, aw = _t(g["A_value"]), _t(g["A_loc"]), _t(g["A_attn"])
    o64 = O.forward_c(v, shapes, lsi, lc, aw, np.float64)
    assert torch.allclose(o64, _t(g["A_out_f64"]))          # reference's own fp64 criterion
    o32 = O.forward_c(v, shapes, lsi, lc, aw, np.float32)
    assert torch.allclose(o32, _t(g["A_out_f32"]), rtol=1e-2, atol=1e-3)   # test.py:56
    assert torch.allclose(o32, _t(g["A_out_f32"]), rtol=1e-5, atol=1e-8)
    assert torch.allclose(O.forward_torch(v.double(), shapes, lc.double(), aw.double()), _t(g["A_out_f64"]))


@pytest.mark.parametrize("tag,seed,rng", [("B", 3, (0.0, 1.0)), ("Bo", 4, (-0.2, 1.2))])
def test_config1_forward_backward(golden_msda, tag, seed, rng):
    g = golden_msda
    value, shapes, lsi, loc, attn = O.make_inputs(2, 100, 8, 32, PYR, 4, seed, loc_range=rng)
    assert lsi.tolist() == [0, 6480, 8100, 8505] and value.shape[1] == 8617
    o32 = O.forward_c(value, shapes, lsi, loc, attn, np.float32)
    assert torch.allclose(o32, _t(g[tag + "_out_f32"]), rtol=1e-4, atol=1e-7)
    o64 = O.forward_c(value, shapes, lsi, loc, attn, np.float64)
    assert torch.allclose(o64, _t(g[tag + "_out_f64"]), rtol=1e-10, atol=1e-13)
    ot = O.forward_torch(value, shapes, loc, attn)
    assert torch.allclose(ot, _t(g[tag + "_out_f32"]), rtol=1e-5, atol=1e-8)

    gen = torch.Generator().manual_seed(seed + 1000)
    gout = torch.rand(o64.shape, generator=gen, dtype=torch.float64)
    gv, gl, ga = O.backward_c(value, shapes, lsi, loc, attn, gout, np.float64)
    assert torch.allclose(gl, _t(g[tag + "_gloc"]), rtol=1e-9, atol=1e-12)
    assert torch.allclose(ga, _t(g[tag + "_gattn"]), rtol=1e-9, atol=1e-12)
    assert torch.allclose(gv.flatten()[::997], _t(g[tag + "_gvalue_sample"]), rtol=1e-9, atol=1e-12)
    sums = torch.tensor([gv.sum().item(), gv.abs().sum().item()], dtype=torch.float64)
    assert torch.allclose(sums, _t(g[tag + "_gvalue_sum"]), rtol=1e-9)


@pytest.mark.parametrize("D", [30, 32, 64, 71])
def test_backward_channel_branches(golden_msda, D):
    g = golden_msda
    value, shapes, lsi, loc, attn = O.make_inputs(1, 2, 2, D, [(6, 4), (3, 2)], 2, 100 + D,
                                                  dtype=torch.float64, loc_range=(-0.1, 1.1))
    o = O.forward_c(value, shapes, lsi, loc, attn, np.float64)
    assert torch.allclose(o, _t(g[f"C{D}_out"]), rtol=1e-10, atol=1e-14)
    gen = torch.Generator().manual_seed(100 + D + 1000)
    gout = torch.rand(o.shape, generator=gen, dtype=torch.float64)
    gv, gl, ga = O.backward_c(value, shapes, lsi, loc, attn, gout, np.float64)
    assert torch.allclose(gv, _t(g[f"C{D}_gvalue"]), rtol=1e-9, atol=1e-13)
    assert torch.allclose(gl, _t(g[f"C{D}_gloc"]), rtol=1e-9, atol=1e-13)
    assert torch.allclose(ga, _t(g[f"C{D}_gattn"]), rtol=1e-9, atol=1e-13)


def test_ragged_levels(golden_msda):
    g = golden_msda
    value, shapes, lsi, loc, attn = O.make_inputs(3, 5, 4, 32, [(1, 1), (2, 7), (5, 3)], 8, 77,
                                                  dtype=torch.float64, loc_range=(-0.3, 1.3))
    o = O.forward_c(value, shapes, lsi, loc, attn, np.float64)
    assert torch.allclose(o, _t(g["D_out"]), rtol=1e-10, atol=1e-14)
    gen = torch.Generator().manual_seed(77 + 1000)
    gout = torch.rand(o.shape, generator=gen, dtype=torch.float64)
    gv, gl, ga = O.backward_c(value, shapes, lsi, loc, attn, gout, np.float64)
    for mine, key in ((gv, "D_gvalue"), (gl, "D_gloc"), (ga, "D_gattn")):
        assert torch.allclose(mine, _t(g[key]), rtol=1e-9, atol=1e-13)
    gv2, gl2, ga2 = O.backward_torch(value, shapes, loc, attn, gout)
    assert torch.allclose(gv2, gv, rtol=1e-9, atol=1e-13) and torch.allclose(gl2, gl, rtol=1e-8, atol=1e-12)
