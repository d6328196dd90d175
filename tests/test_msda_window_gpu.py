"""GPU parity of the self-attention (encoder-shape) kernels behind `vnext_msda_forward_selfattn` — the cache-blocked one
(csrc/msda_cached.cu, the default) and the shared-memory window one (csrc/msda_window.cu, opt-in); every test runs for both:
against the C oracle (restating ms_deform_im2col_cuda.cuh:237-299), against the tile kernel on the same inputs, for
local (windowed), far (global-path) and out-of-range sampling locations, fp32 and bf16 value, both front-ends."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu

SMALL = [(12, 20), (6, 10), (3, 5), (2, 3)]
PYR = [(60, 108), (30, 54), (15, 27), (8, 14)]
PYR720 = [(92, 160), (46, 80), (23, 40), (12, 20)]


KIND = {"auto": "window", "cached": "cached"}


@pytest.fixture(scope="module", params=["cached", "auto"])
def ops(request):
    """The op layer with one of the two self-attention kernels selected for this module run."""
    from vnext_b200 import _cabi, msda
    _cabi.lib()
    prev = msda.WINDOW_MODE
    msda.WINDOW_MODE = request.param
    msda.TEST_MODE = request.param
    yield msda
    msda.WINDOW_MODE = prev


def _variant():
    from vnext_b200 import _cabi
    return _cabi.last_variant()


def encoder_inputs(N, M, pyr, K, seed, sigma_px=1.5, outlier=0.05, dtype=torch.float32):
    """Self-attention inputs: query i = pixel i of the pyramid; loc = pixel centre (mapped to every level) + N(0, sigma)
    pixel offsets, a fraction `outlier` of the samples thrown anywhere in [-0.2, 1.2]^2 (global path / border rules)."""
    g = torch.Generator().manual_seed(seed)
    L = len(pyr)
    S = sum(h * w for h, w in pyr)
    shapes = torch.tensor(pyr, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    refs = []
    for h, w in pyr:
        ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
        refs.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    ref = torch.cat(refs, 0)                                        # [S, 2]
    wh = torch.tensor([[w, h] for h, w in pyr], dtype=torch.float32)  # [L, 2]
    off = torch.randn((N, S, M, L, K, 2), generator=g) * sigma_px / wh[None, None, None, :, None, :]
    loc = ref[None, :, None, None, None, :] + off
    far = torch.rand((N, S, M, L, K), generator=g) < outlier
    loc[far] = torch.rand((int(far.sum()), 2), generator=g) * 1.4 - 0.2
    attn = torch.rand((N, S, M, L, K), generator=g) + 1e-5
    attn = attn / attn.sum((-1, -2), keepdim=True)
    value = (torch.rand((N, S, M, 32), generator=g) * 0.01).to(dtype)
    return value, shapes, lsi, loc.contiguous(), attn.contiguous()


def _cuda(*ts):
    return [t.cuda() for t in ts]


@pytest.mark.parametrize("pyr,N,M,K,sigma,outlier", [
    (SMALL, 2, 8, 4, 1.5, 0.05), (SMALL, 1, 3, 8, 3.0, 0.2), (SMALL[:2], 2, 8, 4, 1.0, 0.0), (SMALL[:1], 1, 8, 8, 2.0, 0.1),
    ([(5, 1), (1, 7), (1, 1), (3, 3)], 2, 2, 4, 1.0, 0.3), (PYR, 1, 8, 4, 1.6, 0.02)])
def test_window_fp32_matches_oracle_and_tile_kernel(ops, pyr, N, M, K, sigma, outlier):
    value, shapes, lsi, loc, attn = encoder_inputs(N, M, pyr, K, 11, sigma, outlier)
    dv = _cuda(value, shapes, lsi, loc, attn)
    out = ops.ms_deform_attn_forward(*dv, 64).cpu()
    assert _variant() == f"fwd_{KIND[ops.TEST_MODE]}_f32"
    oc = O.forward_c(value, shapes, lsi, loc, attn, np.float32)
    assert torch.allclose(out, oc, rtol=1e-4, atol=1e-7)
    ops.WINDOW_MODE = "off"
    try:
        tile = ops.ms_deform_attn_forward(*dv, 64).cpu()
        assert not _variant().startswith(("fwd_window", "fwd_cached"))
    finally:
        ops.WINDOW_MODE = ops.TEST_MODE
    assert torch.allclose(out, tile, rtol=1e-5, atol=1e-8)


def test_window_random_locations_everything_on_the_global_path(ops):
    """No locality at all (uniform locations incl. out of range): the window is only a cache, results stay exact."""
    value, shapes, lsi, loc, attn = O.make_inputs(2, sum(h * w for h, w in SMALL), 8, 32, SMALL, 4, 5, loc_range=(-0.2, 1.2))
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 64).cpu()
    assert _variant() == f"fwd_{KIND[ops.TEST_MODE]}_f32"
    assert torch.allclose(out, O.forward_c(value, shapes, lsi, loc, attn, np.float32), rtol=1e-4, atol=1e-7)


def test_window_nonfinite_locations_are_skipped_like_the_reference(ops):
    value, shapes, lsi, loc, attn = encoder_inputs(1, 8, SMALL, 4, 3)
    loc[0, 5, 0, 0, 0, 0] = float("nan")
    loc[0, 9, 1, 1, 2, 1] = float("inf")
    loc[0, 17, 2, 3, 1, 0] = -1e30
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 64).cpu()
    assert torch.isfinite(out).all()
    assert torch.allclose(out, O.forward_c(value, shapes, lsi, loc, attn, np.float32), rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("pyr,K", [(SMALL, 4), (PYR, 4), (SMALL, 8)])
def test_window_bf16_value(ops, pyr, K):
    value, shapes, lsi, loc, attn = encoder_inputs(2 if pyr is SMALL else 1, 8, pyr, K, 7, dtype=torch.bfloat16)
    out = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 64)
    assert _variant() == f"fwd_{KIND[ops.TEST_MODE]}_bf16" and out.dtype == torch.bfloat16
    ref = O.forward_c(value.float(), shapes, lsi, loc, attn, np.float32)       # exact on the bf16-rounded value
    assert torch.allclose(out.float().cpu(), ref, rtol=8e-3, atol=1e-6)           # one bf16 rounding of the output


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ref_dim", [2, 4])
def test_window_fused_front_end_matches_the_fused_tile_kernel(ops, dtype, ref_dim):
    g = torch.Generator().manual_seed(2)
    N, M, L, K = 2, 8, 4, 4
    pyr = SMALL
    S = sum(h * w for h, w in pyr)
    value, shapes, lsi, loc, _ = encoder_inputs(N, M, pyr, K, 9, dtype=dtype)
    oa = torch.randn((N, S, M * L * K * 3), generator=g)
    oa[..., :M * L * K * 2] *= 2.0                                   # pixel-scale offsets
    ref = loc[:1, :, 0, :, 0, :].contiguous()                        # [1, S, L, 2] around the pixel centres
    if ref_dim == 4:
        ref = torch.cat((ref, torch.full_like(ref, 0.2)), -1).contiguous()
    dv, ds, dl, doa, dref = _cuda(value, shapes, lsi, oa, ref)
    out = ops.ms_deform_attn_fused(dv, ds, dl, doa, dref, L, K)
    if ref_dim == 2:
        assert _variant() == f"fwd_{KIND[ops.TEST_MODE]}_fused_" + ("f32" if dtype == torch.float32 else "bf16")
    else:
        assert _variant().startswith("fwd_fused_")         # box-shaped references (decoder) stay on the tile kernel
    ops.WINDOW_MODE = "off"
    try:
        tile = ops.ms_deform_attn_fused(dv, ds, dl, doa, dref, L, K)
        assert _variant().startswith("fwd_fused_")
    finally:
        ops.WINDOW_MODE = ops.TEST_MODE
    tol = dict(rtol=1e-5, atol=1e-8) if dtype == torch.float32 else dict(rtol=8e-3, atol=1e-6)
    assert torch.allclose(out.float(), tile.float(), **tol)


def test_window_full_size_720p_equals_tile_kernel(ops):
    """BASELINE geometry (736x1280 padded pyramid, T=5 frames as the batch): window kernel == tile kernel on every
    output (the tile kernel is pinned to the oracle / reference goldens in test_msda_gpu.py)."""
    value, shapes, lsi, loc, attn = encoder_inputs(5, 8, PYR720, 4, 1, 1.6, 0.01)
    dv = _cuda(value, shapes, lsi, loc, attn)
    out = ops.ms_deform_attn_forward(*dv, 64)
    assert _variant() == f"fwd_{KIND[ops.TEST_MODE]}_f32"
    ops.WINDOW_MODE = "off"
    try:
        tile = ops.ms_deform_attn_forward(*dv, 64)
    finally:
        ops.WINDOW_MODE = ops.TEST_MODE
    assert torch.allclose(out, tile, rtol=1e-5, atol=1e-8)
    sub = slice(0, 1)                                                 # one frame against the C oracle
    oc = O.forward_c(value[sub], shapes, lsi, loc[sub], attn[sub], np.float32)
    assert torch.allclose(out[sub].cpu(), oc, rtol=1e-4, atol=1e-7)


def test_unordered_levels_fall_back_to_the_tile_kernel(ops):
    """level_start_index that does not tile [0, S) in order: the window kernel does not apply, the op still works."""
    value, shapes, lsi, loc, attn = encoder_inputs(1, 8, SMALL[:2], 4, 3)
    S = value.shape[1]
    lsi2 = torch.tensor([60, 0])                                      # level 1 stored first
    value2 = torch.cat((value[:, 240:], value[:, :240]), 1).contiguous()
    out = ops.ms_deform_attn_forward(*_cuda(value2, shapes, lsi2, loc, attn), 64).cpu()
    assert not _variant().startswith(("fwd_window", "fwd_cached"))
    ref = ops.ms_deform_attn_forward(*_cuda(value, shapes, lsi, loc, attn), 64).cpu()
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-8) and S == 300
