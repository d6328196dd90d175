"""GPU parity of the encoder's "front-end in the GEMM epilogue" path: `vnext_linear_msda_front` (offsets | logits projection
whose epilogue writes sampling locations | softmax weights — MSDeformAttn.forward, ms_deform_attn.py:98-108) and the
cache-blocked sampler's pre-located mode reading those rows (`vnext_msda_forward_selfattn` with reference_points == NULL)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SMALL = [(12, 20), (6, 10), (3, 5), (2, 3)]
PYR = [(60, 108), (30, 54), (15, 27), (8, 14)]
M_, L_, P_ = 8, 4, 4


def _geometry(pyr):
    refs = []
    for h, w in pyr:
        ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
        refs.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    ref = torch.cat(refs, 0)                                               # [S, 2]
    ratios = torch.tensor([[0.97, 0.94], [0.98, 0.95], [1.0, 0.96], [0.93, 1.0]])
    return (ref[:, None, :] * ratios[None]).contiguous()                   # [S, L, 2], like enc_ref (valid ratios differ)


def _inputs(pyr, T, K, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in pyr)
    x = torch.randn(T, S, K, generator=g)
    w = torch.randn(384, K, generator=g) / K ** 0.5
    w[:256] *= 2.0                                                          # pixel-scale offsets
    b2 = torch.randn(S, 384, generator=g) * 0.5
    return x.to(dtype).cuda(), w.to(dtype).cuda(), b2.cuda(), _geometry(pyr).cuda()


def _expected(x, w, b2, ref, pyr):
    T, S, _ = x.shape
    raw = F.linear(x.double(), w.double()) + b2.double()[None]
    off = raw[..., :256].reshape(T, S, M_, L_, P_, 2)
    wh = torch.tensor([[ww, hh] for hh, ww in pyr], dtype=torch.float64, device=x.device)
    loc = ref.double()[None, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    attn = torch.softmax(raw[..., 256:].reshape(T, S, M_, L_ * P_), -1)
    return loc.reshape(T, S, 256), attn.reshape(T, S, 128), raw


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("pyr,T,K", [(SMALL, 3, 256), (PYR, 2, 256), (SMALL, 1, 64)])
def test_epilogue_writes_locations_and_softmax(pyr, T, K, dtype):
    from vnext_b200.linear import linear_msda_front
    x, w, b2, ref = _inputs(pyr, T, K, 3, dtype)
    out = linear_msda_front(x, w, b2, ref, pyr)
    loc, attn, raw = _expected(x, w, b2, ref, pyr)
    assert out.dtype == torch.float32 and out.shape == (T, x.shape[1], 384)
    tol = (2e-3 if dtype == torch.float32 else 2e-5) * raw.abs().max().item()       # TF32 products / exact bf16 products
    min_wh = min(min(h, w_) for h, w_ in pyr)
    assert (out[..., :256].double() - loc).abs().max().item() < tol / min_wh + 1e-6
    assert (out[..., 256:].double() - attn).abs().max().item() < tol
    assert (out[..., 256:].reshape(T, -1, M_, 16).sum(-1) - 1).abs().max().item() < 1e-5


@pytest.mark.parametrize("vdtype", [torch.float32, torch.bfloat16])
def test_located_sampler_equals_fused_sampler_on_the_same_projection(vdtype):
    """linear -> fused sampler  ==  linear_msda_front -> pre-located sampler (same GEMM, same arithmetic, different kernel)."""
    from vnext_b200 import _cabi, msda
    from vnext_b200.linear import linear, linear_msda_front
    pyr, T = PYR, 2
    x, w, b2, ref = _inputs(pyr, T, 256, 5, torch.float32)
    S = x.shape[1]
    value = (torch.rand(T, S, M_, 32, generator=torch.Generator().manual_seed(6)) * 0.01).to(vdtype).cuda()
    shapes = torch.tensor(pyr, dtype=torch.long).cuda()
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    oa = linear(x, w, None, bias2d=b2)
    fused = msda.ms_deform_attn_fused(value, shapes, lsi, oa, ref[None].contiguous(), L_, P_)
    assert _cabi.last_variant().startswith("fwd_cached_fused_")
    la = linear_msda_front(x, w, b2, ref, pyr)
    located = msda.ms_deform_attn_located(value, la, L_, P_, msda.host_geometry(shapes, lsi))
    assert _cabi.last_variant().startswith("fwd_cached_pre_")
    tol = dict(rtol=2e-5, atol=1e-8) if vdtype == torch.float32 else dict(rtol=8e-3, atol=1e-6)
    assert torch.allclose(located.float(), fused.float(), **tol)


def test_located_sampler_against_the_oracle():
    import numpy as np
    from oracle import msda_oracle as O
    from vnext_b200 import msda
    from vnext_b200.linear import linear_msda_front
    pyr, T = SMALL, 2
    x, w, b2, ref = _inputs(pyr, T, 256, 9, torch.float32)
    S = x.shape[1]
    value = torch.rand(T, S, M_, 32, generator=torch.Generator().manual_seed(10)) * 0.01
    shapes = torch.tensor(pyr, dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    la = linear_msda_front(x, w, b2, ref, pyr)
    out = msda.ms_deform_attn_located(value.cuda(), la, L_, P_, msda.host_geometry(shapes, lsi))
    loc = la[..., :256].reshape(T, S, M_, L_, P_, 2).cpu().contiguous()
    attn = la[..., 256:].reshape(T, S, M_, L_, P_).cpu().contiguous()
    want = O.forward_c(value, shapes, lsi, loc, attn, np.float32)
    assert torch.allclose(out.cpu(), want, rtol=1e-4, atol=1e-7)


def test_argument_errors():
    from vnext_b200 import _cabi
    lib = _cabi.lib()
    a = torch.zeros(16, 256, device="cuda")
    rc = lib.vnext_linear_msda_front(0, a.data_ptr(), 256, a.data_ptr(), None, 8, 384, None, None, None, 384, 16, 256, None)
    assert rc != 0 and b"vnext_linear_msda_front" in lib.vnext_last_error()
