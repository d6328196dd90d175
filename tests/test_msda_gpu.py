"""GPU parity tests for the MSDA kernels: CUDA path (through the C ABI) vs the CPU oracle and the
golden vectors generated from the reference's own function."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu

PYR = [(60, 108), (30, 54), (15, 27), (8, 14)]
PYR720 = [(92, 160), (46, 80), (23, 40), (12, 20)]


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from vnext_b200 import _cabi, msda
    _cabi.lib()
    return msda


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _cuda(*ts):
    return [t.cuda() for t in ts]


def _variant():
    from vnext_b200 import _cabi
    return _cabi.last_variant()


def test_reference_fixture_fp64_fp32(ops, golden_msda):
    g = golden_msda
    shapes = _t(g["A_shapes"])
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    v, lc, aw = _t(g["A_value"]), _t(g["A_loc"]), _t(g["A_attn"])
    out64 = ops.ms_deform_attn_forward(*_cuda(v.double(), shapes, lsi, lc.double(), aw.double()), 2).cpu()
    assert torch.allclose(out64, _t(g["A_out_f64"]))                       # test.py:31-44 criterion
    out32 = ops.ms_deform_attn_forward(*_cuda(v, shapes, lsi, lc, aw), 2).cpu()
    assert torch.allclose(out32, _t(g["A_out_f32"]), rtol=1e-2, atol=1e-3)  # test.py:56 criterion
    assert torch.allclose(out32, _t(g["A_out_f32"]), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("tag,seed,rng", [("B", 3, (0.0, 1.0)), ("Bo", 4, (-0.2, 1.2))])
def test_config1_forward(ops, golden_msda, tag, seed, rng):
    value, shapes, lsi, loc, attn = O.make_inputs(2, 100, 8, 32, PYR, 4, seed, loc_range=rng)
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 64).cpu()
    assert _variant() == "fwd_f32_d32_tma"
    gold = _t(golden_msda[tag + "_out_f32"])
    assert torch.allclose(out, gold, rtol=1e-2, atol=1e-3)     # the reference's fp32 criterion
    assert (out - gold).abs().max() < 1e-6                      # values are O(0.01): this is ~1e-4 relative
    oc = O.forward_c(value, shapes, lsi, loc, attn, np.float32)
    assert torch.allclose(out, oc, rtol=1e-4, atol=1e-7)
    o64 = _t(golden_msda[tag + "_out_f64"])
    assert (out.double() - o64).abs().max() < 1e-7


@pytest.mark.parametrize("tag,seed,rng", [("B", 3, (0.0, 1.0)), ("Bo", 4, (-0.2, 1.2))])
def test_config1_backward(ops, golden_msda, tag, seed, rng):
    g = golden_msda
    value, shapes, lsi, loc, attn = O.make_inputs(2, 100, 8, 32, PYR, 4, seed, loc_range=rng)
    gen = torch.Generator().manual_seed(seed + 1000)
    gout = torch.rand((2, 100, 256), generator=gen, dtype=torch.float64)
    gv, gl, ga = ops.ms_deform_attn_backward(*_cuda(value, shapes, lsi, loc, attn, gout.float()), 64)
    assert _variant() == "bwd_f32_d32_tma"
    gv, gl, ga = gv.cpu(), gl.cpu(), ga.cpu()
    assert torch.allclose(gl.double(), _t(g[tag + "_gloc"]), rtol=1e-3, atol=1e-5)
    assert torch.allclose(ga.double(), _t(g[tag + "_gattn"]), rtol=1e-3, atol=1e-6)
    assert torch.allclose(gv.double().flatten()[::997], _t(g[tag + "_gvalue_sample"]), rtol=1e-3, atol=1e-6)
    sums = torch.tensor([gv.double().sum().item(), gv.double().abs().sum().item()], dtype=torch.float64)
    assert torch.allclose(sums, _t(g[tag + "_gvalue_sum"]), rtol=1e-4)
    cv, cl, ca = O.backward_c(value, shapes, lsi, loc, attn, gout, np.float64)
    assert torch.allclose(gv.double(), cv, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("D", [30, 32, 64, 71])
def test_fp64_channel_branches(ops, golden_msda, D):
    g = golden_msda
    value, shapes, lsi, loc, attn = O.make_inputs(1, 2, 2, D, [(6, 4), (3, 2)], 2, 100 + D,
                                                  dtype=torch.float64, loc_range=(-0.1, 1.1))
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 2).cpu()
    assert torch.allclose(out, _t(g[f"C{D}_out"]), rtol=1e-10, atol=1e-14)
    gen = torch.Generator().manual_seed(100 + D + 1000)
    gout = torch.rand(out.shape, generator=gen, dtype=torch.float64)
    gv, gl, ga = ops.ms_deform_attn_backward(*_cuda(value, shapes, lsi, loc, attn, gout), 2)
    assert torch.allclose(gv.cpu(), _t(g[f"C{D}_gvalue"]), rtol=1e-9, atol=1e-13)
    assert torch.allclose(gl.cpu(), _t(g[f"C{D}_gloc"]), rtol=1e-9, atol=1e-13)
    assert torch.allclose(ga.cpu(), _t(g[f"C{D}_gattn"]), rtol=1e-9, atol=1e-13)


def test_ragged_levels_fp64(ops, golden_msda):
    g = golden_msda
    value, shapes, lsi, loc, attn = O.make_inputs(3, 5, 4, 32, [(1, 1), (2, 7), (5, 3)], 8, 77,
                                                  dtype=torch.float64, loc_range=(-0.3, 1.3))
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 64).cpu()
    assert torch.allclose(out, _t(g["D_out"]), rtol=1e-10, atol=1e-14)
    gen = torch.Generator().manual_seed(77 + 1000)
    gout = torch.rand(out.shape, generator=gen, dtype=torch.float64)
    gv, gl, ga = ops.ms_deform_attn_backward(*_cuda(value, shapes, lsi, loc, attn, gout), 64)
    for mine, key in ((gv, "D_gvalue"), (gl, "D_gloc"), (ga, "D_gattn")):
        assert torch.allclose(mine.cpu(), _t(g[key]), rtol=1e-9, atol=1e-13)
    # the same case in fp32 goes through the tile kernels (D=32, L*K=24 is NOT supported -> generic)
    out32 = ops.ms_deform_attn_forward(*_cuda(value.float(), shapes, lsi, loc.float(), attn.float()), 64).cpu()
    assert torch.allclose(out32.double(), _t(g["D_out"]), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("L,K", [(1, 4), (1, 8), (2, 4), (4, 4), (4, 8), (2, 2), (3, 4)])
@pytest.mark.parametrize("N,Lq,M", [(1, 37, 3), (2, 300, 8), (3, 1, 1)])
def test_tile_kernels_vs_oracle(ops, L, K, N, Lq, M):
    shapes = [(13, 9), (7, 5), (4, 3), (2, 2)][:L]
    value, shapes_t, lsi, loc, attn = O.make_inputs(N, Lq, M, 32, shapes, K, 1000 + L * 10 + K + Lq,
                                                    loc_range=(-0.25, 1.25))
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes_t, lsi, loc, attn), 64).cpu()
    fast = (L * K) in (4, 8, 16, 32)
    assert _variant() == ("fwd_f32_d32_tma" if fast else "fwd_generic")
    oc = O.forward_c(value, shapes_t, lsi, loc, attn, np.float32)
    assert torch.allclose(out, oc, rtol=1e-4, atol=1e-7)
    gen = torch.Generator().manual_seed(5)
    gout = torch.rand(out.shape, generator=gen)
    gv, gl, ga = ops.ms_deform_attn_backward(*_cuda(value, shapes_t, lsi, loc, attn, gout), 64)
    assert _variant() == ("bwd_f32_d32_tma" if fast else "bwd_generic")
    cv, cl, ca = O.backward_c(value, shapes_t, lsi, loc, attn, gout, np.float64)
    assert torch.allclose(gv.cpu().double(), cv, rtol=1e-3, atol=1e-6)
    assert torch.allclose(gl.cpu().double(), cl, rtol=1e-3, atol=1e-5)
    assert torch.allclose(ga.cpu().double(), ca, rtol=1e-3, atol=1e-6)


def test_misaligned_views_take_ldg_staging(ops):
    value, shapes, lsi, loc, attn = O.make_inputs(1, 50, 8, 32, PYR, 4, 21)
    dv, dsh, dlsi, dloc, dattn = _cuda(value, shapes, lsi, loc, attn)
    # 8-byte-aligned but not 16-byte-aligned storage for loc, 4-byte for attn
    buf_l = torch.empty(dloc.numel() + 2, device="cuda")
    loc_m = buf_l[2:].view_as(dloc).copy_(dloc)
    buf_a = torch.empty(dattn.numel() + 1, device="cuda")
    attn_m = buf_a[1:].view_as(dattn).copy_(dattn)
    assert loc_m.data_ptr() % 16 == 8 and attn_m.data_ptr() % 16 == 4
    ref = ops.ms_deform_attn_forward(dv, dsh, dlsi, dloc, dattn, 64)
    out = ops.ms_deform_attn_forward(dv, dsh, dlsi, loc_m, attn_m, 64)
    assert _variant() == "fwd_f32_d32_ldg"
    assert torch.equal(out, ref)


def test_bf16_forward(ops):
    value, shapes, lsi, loc, attn = O.make_inputs(2, 300, 8, 32, PYR, 4, 31, loc_range=(-0.1, 1.1))
    value = (value * 100).bfloat16()           # O(1) values
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 64)
    assert out.dtype == torch.bfloat16 and _variant() == "fwd_bf16_d32_tma"
    # oracle for bf16 = fp32 oracle evaluated on the bf16-rounded inputs
    oc = O.forward_c(value.float(), shapes, lsi, loc, attn, np.float32)
    assert torch.allclose(out.cpu().float(), oc, rtol=1e-2, atol=1e-2)
    assert (out.cpu().float() - oc).abs().max() <= 2 ** -8 * oc.abs().max()   # one bf16 rounding of the result
    # backward (generic path, fp32 accumulation)
    gout = torch.rand(out.shape, generator=torch.Generator().manual_seed(1)).bfloat16()
    gv, gl, ga = ops.ms_deform_attn_backward(*_cuda(value, shapes, lsi, loc, attn, gout), 64)
    cv, cl, ca = O.backward_c(value.float(), shapes, lsi, loc, attn, gout.float(), np.float64)
    assert gv.dtype == torch.bfloat16 and gl.dtype == torch.float32
    assert torch.allclose(gv.cpu().double(), cv, rtol=1e-2, atol=1e-2)
    assert torch.allclose(gl.cpu().double(), cl, rtol=1e-3, atol=1e-3)
    assert torch.allclose(ga.cpu().double(), ca, rtol=1e-3, atol=1e-4)


def test_full_size_encoder_call_properties(ops):
    """BASELINE sizes (T=5 frames folded into the batch, Lq = S = 8617 / 19560): size-independent checks."""
    for pyr, n in ((PYR, 5), (PYR720, 2)):
        S = sum(h * w for h, w in pyr)
        value, shapes, lsi, loc, attn = O.make_inputs(n, S, 8, 32, pyr, 4, 41, loc_range=(-0.05, 1.05))
        dv, dsh, dlsi, dloc, dattn = _cuda(value, shapes, lsi, loc, attn)
        out = ops.ms_deform_attn_forward(dv, dsh, dlsi, dloc, dattn, 64)
        # (1) a strided subset of queries against the oracle (queries are independent)
        idx = torch.arange(0, S, 211)
        oc = O.forward_c(value, shapes, lsi, loc[:, idx].contiguous(), attn[:, idx].contiguous(), np.float32)
        assert torch.allclose(out[:, idx.cuda()].cpu(), oc, rtol=1e-4, atol=1e-7)
        # (2) linearity in value and (3) convex-combination bound
        out2 = ops.ms_deform_attn_forward(dv * 2, dsh, dlsi, dloc, dattn, 64)
        assert torch.allclose(out2, out * 2, rtol=1e-6, atol=1e-9)
        assert out.max() <= value.max() + 1e-7 and out.min() >= 0
        # (4) backward: sum(grad_value) == sum over samples of (bilinear weight mass * attn * gout)
        gout = torch.ones_like(out)
        gv, gl, ga = ops.ms_deform_attn_backward(dv, dsh, dlsi, dloc, dattn, gout, 64)
        ones = torch.ones_like(dv)
        mass = ops.ms_deform_attn_forward(ones, dsh, dlsi, dloc, dattn, 64)   # = sum_s attn * in-range weight mass
        assert torch.allclose(gv.sum().double(), mass.sum().double(), rtol=1e-4)
        # grad_attn of a sample = <gout, bilinear value>; summing attn*grad_attn reproduces sum(out)
        assert torch.allclose((ga * dattn).sum().double(), out.sum().double(), rtol=1e-4)
        # idempotence / determinism of the forward
        assert torch.equal(ops.ms_deform_attn_forward(dv, dsh, dlsi, dloc, dattn, 64), out)


def test_gradcheck_fp64_like_reference(ops):
    """ops/test.py:63-78: gradcheck of the autograd Function in double (generic kernels)."""
    from torch.autograd import gradcheck
    for D in (30, 32, 64, 71):
        value, shapes, lsi, loc, attn = O.make_inputs(1, 2, 2, D, [(6, 4), (3, 2)], 2, 3, dtype=torch.float64)
        dv, dsh, dlsi, dloc, dattn = _cuda(value, shapes, lsi, loc, attn)
        dv.requires_grad_(True); dloc.requires_grad_(True); dattn.requires_grad_(True)
        assert gradcheck(ops.MSDeformAttnFunction.apply, (dv, dsh, dlsi, dloc, dattn, 2))


def test_large_channel_counts(ops):
    """The reference gradchecks D = 1025 / 2048 / 3096 (multi-block / gm kernels); compare to the oracle."""
    for D in (1025, 2048):
        value, shapes, lsi, loc, attn = O.make_inputs(1, 2, 2, D, [(6, 4), (3, 2)], 2, 9, dtype=torch.float64)
        out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 2).cpu()
        assert torch.allclose(out, O.forward_c(value, shapes, lsi, loc, attn, np.float64), rtol=1e-10, atol=1e-14)
        gout = torch.rand(out.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
        gv, gl, ga = ops.ms_deform_attn_backward(*_cuda(value, shapes, lsi, loc, attn, gout), 2)
        cv, cl, ca = O.backward_c(value, shapes, lsi, loc, attn, gout, np.float64)
        assert torch.allclose(gv.cpu(), cv, rtol=1e-9, atol=1e-13)
        assert torch.allclose(gl.cpu(), cl, rtol=1e-9, atol=1e-13)
        assert torch.allclose(ga.cpu(), ca, rtol=1e-9, atol=1e-13)


def test_empty_and_error_behaviour(ops):
    value, shapes, lsi, loc, attn = O.make_inputs(2, 4, 8, 32, PYR, 4, 1)
    dv, dsh, dlsi, dloc, dattn = _cuda(value, shapes, lsi, loc, attn)
    out = ops.ms_deform_attn_forward(dv, dsh, dlsi, dloc[:, :0].contiguous(), dattn[:, :0].contiguous(), 64)
    assert out.shape == (2, 0, 256)
    out = ops.ms_deform_attn_forward(dv[:0].contiguous(), dsh, dlsi, dloc[:0].contiguous(), dattn[:0].contiguous(), 64)
    assert out.shape == (0, 4, 256)
    with pytest.raises(RuntimeError, match="contiguous"):
        ops.ms_deform_attn_forward(dv.transpose(0, 1), dsh, dlsi, dloc, dattn, 64)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        ops.ms_deform_attn_forward(torch.cat([dv, dv[:1]]), dsh, dlsi, torch.cat([dloc, dloc[:1]]),
                                   torch.cat([dattn, dattn[:1]]), 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.ms_deform_attn_forward(dv, dsh.cpu(), dlsi, dloc, dattn, 64)
    # NaN / inf locations are skipped samples, not poison
    bad = dloc.clone()
    bad[0, 0, 0, 0, 0, 0] = float("nan")
    bad[0, 1, 0, 0, 0, 1] = float("inf")
    o = ops.ms_deform_attn_forward(dv, dsh, dlsi, bad, dattn, 64)
    assert torch.isfinite(o).all()


def test_host_buffer_cabi(ops):
    """The reference-facing C ABI with HOST buffers (what a non-torch caller binds)."""
    from vnext_b200 import _cabi
    lib = _cabi.lib()
    value, shapes, lsi, loc, attn = O.make_inputs(2, 100, 8, 32, PYR, 4, 3)
    v, lc, aw = value.numpy(), loc.numpy(), attn.numpy()
    sh, ls = shapes.numpy(), lsi.numpy()
    out = np.empty((2, 100, 256), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.vnext_msda_forward_host(0, p(v), p(sh), p(ls), p(lc), p(aw), 2, 8617, 8, 32, 4, 100, 4, p(out))
    assert rc == 0, lib.vnext_last_error()
    oc = O.forward_c(value, shapes, lsi, loc, attn, np.float32)
    assert torch.allclose(torch.from_numpy(out), oc, rtol=1e-4, atol=1e-7)
    gout = np.random.default_rng(0).random((2, 100, 256), np.float32)
    gv, gl, ga = np.empty_like(v), np.empty_like(lc), np.empty_like(aw)
    rc = lib.vnext_msda_backward_host(0, p(v), p(sh), p(ls), p(lc), p(aw), p(gout), 2, 8617, 8, 32, 4, 100, 4,
                                      p(gv), p(gl), p(ga))
    assert rc == 0, lib.vnext_last_error()
    cv, cl, ca = O.backward_c(value, shapes, lsi, loc, attn, torch.from_numpy(gout), np.float64)
    assert torch.allclose(torch.from_numpy(gv).double(), cv, rtol=1e-3, atol=1e-6)
    assert torch.allclose(torch.from_numpy(gl).double(), cl, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("ref_dim", [2, 4])
@pytest.mark.parametrize("N,Lq,M,L,K,shared", [(3, 300, 8, 4, 4, False), (5, 777, 8, 4, 4, True), (2, 50, 4, 2, 4, False),
                                                (1, 33, 8, 4, 8, False), (2, 19, 8, 1, 4, True)])
def test_fused_frontend_vs_oracle(ops, N, Lq, M, L, K, shared, ref_dim):
    """softmax + sampling-location arithmetic fused into the sampler == the module's ATen pre-ops + oracle core
    (ops/modules/ms_deform_attn.py:98-108)."""
    shapes = [(23, 31), (12, 16), (6, 8), (3, 4)][:L]
    g = torch.Generator().manual_seed(N * 100 + Lq + ref_dim)
    value, shapes_t, lsi, _, _ = O.make_inputs(N, 1, M, 32, shapes, K, 5)
    value = torch.randn(value.shape, generator=g)
    oa = torch.randn(N, Lq, M * L * K * 3, generator=g) * 2
    ref = torch.rand((1 if shared else N), Lq, L, ref_dim, generator=g)
    if ref_dim == 4:
        ref[..., 2:] *= 0.3
    n_off = M * L * K * 2
    off = oa[..., :n_off].reshape(N, Lq, M, L, K, 2)
    attn = torch.softmax(oa[..., n_off:].reshape(N, Lq, M, L * K), -1).view(N, Lq, M, L, K)
    r = ref.expand(N, -1, -1, -1)
    if ref_dim == 2:
        norm = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1).float()
        loc = r[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = r[:, :, None, :, None, :2] + off / K * r[:, :, None, :, None, 2:] * 0.5
    want = O.forward_c(value, shapes_t, lsi, loc, attn, np.float32)
    got = ops.ms_deform_attn_fused(value.cuda(), shapes_t.cuda(), lsi.cuda(), oa.cuda(), ref.cuda(), L, K)
    assert _variant() == "fwd_fused_f32_d32_tma"
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=2e-5)
    # strided projection rows (the op reads them in place out of a wider GEMM output)
    wide = torch.zeros(N, Lq, oa.shape[-1] + 64)
    wide[..., :oa.shape[-1]] = oa
    got2 = ops.ms_deform_attn_fused(value.cuda(), shapes_t.cuda(), lsi.cuda(), wide.cuda()[..., :oa.shape[-1]],
                                    ref.cuda(), L, K)
    assert torch.equal(got2, got)


def test_backward_large_lk32_regression(ops):
    """L*K = 32 with enough pairs to pick the largest tile (shared-memory sizing of the backward tile kernel)."""
    shapes = [(13, 9), (7, 5), (4, 3), (2, 2)]
    value, shapes_t, lsi, loc, attn = O.make_inputs(8, 1000, 8, 32, shapes, 8, 77, loc_range=(-0.1, 1.1))
    gout = torch.rand(8, 1000, 256, generator=torch.Generator().manual_seed(3))
    gv, gl, ga = ops.ms_deform_attn_backward(*_cuda(value, shapes_t, lsi, loc, attn, gout), 64)
    assert _variant() == "bwd_f32_d32_tma"
    idx = torch.arange(0, 1000, 97)
    cv, cl, ca = O.backward_c(value, shapes_t, lsi, loc[:, idx].contiguous(), attn[:, idx].contiguous(),
                              gout[:, idx].contiguous(), np.float64)
    assert torch.allclose(gl.cpu()[:, idx].double(), cl, rtol=1e-3, atol=1e-5)
    assert torch.allclose(ga.cpu()[:, idx].double(), ca, rtol=1e-3, atol=1e-6)
