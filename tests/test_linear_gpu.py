"""GPU numerics of the tcgen05 linear kernel vs a plain PyTorch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


def _mk(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) * 0.1
    return x.cuda(), w.cuda(), b.cuda()


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (300, 256, 256), (1000, 384, 256), (97800, 256, 256),
                                   (4097, 1024, 256), (513, 256, 1024), (77, 64, 64), (1500, 128, 2048), (5, 169, 256), (2000, 640, 256), (300, 384, 64)])
def test_bias_and_relu(M, N, K):
    from vnext_b200.linear import linear
    torch.backends.cuda.matmul.allow_tf32 = False
    x, w, b = _mk(M, N, K)
    ref = F.linear(x.double(), w.double(), b.double())
    out = linear(x, w, b)
    # TF32 inputs (10-bit mantissa): error ~ 2^-11 * sqrt(K) * |x||w| per term; bound relative to output scale
    tol = 2e-3 * ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() < tol
    out_r = linear(x, w, b, relu=True)
    assert (out_r.double() - ref.clamp(min=0)).abs().max().item() < tol
    out_nb = linear(x, w, None)
    assert (out_nb.double() - F.linear(x.double(), w.double())).abs().max().item() < tol


@pytest.mark.parametrize("M,K", [(128, 256), (333, 1024), (97800, 256)])
def test_residual_layernorm_epilogue(M, K):
    from vnext_b200.linear import linear
    x, w, b = _mk(M, 256, K, 1)
    g = torch.Generator().manual_seed(2)
    res = torch.randn(M, 256, generator=g).cuda()
    gamma = (torch.rand(256, generator=g) + 0.5).cuda()
    beta = (torch.randn(256, generator=g) * 0.1).cuda()
    ref = F.layer_norm(res.double() + F.linear(x.double(), w.double(), b.double()), (256,), gamma.double(),
                       beta.double(), 1e-5)
    out = linear(x, w, b, residual=res, ln_weight=gamma, ln_bias=beta)
    # LayerNorm rescales rows to unit variance, so the TF32 input rounding (2^-11 per operand) shows up as ~1e-3 abs
    assert (out.double() - ref).abs().max().item() < 8e-3
    assert (out.double() - ref).abs().mean().item() < 5e-4


def test_strided_views_and_errors():
    from vnext_b200.linear import linear
    x, w, b = _mk(256, 256, 256, 3)
    big = torch.zeros(256, 512, device="cuda")
    big[:, :256] = x
    out = linear(big[:, :256], w, b)                      # lda = 512
    ref = F.linear(x.double(), w.double(), b.double())
    assert (out.double() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    with pytest.raises(RuntimeError, match="N == 256"):
        linear(x, torch.randn(384, 256, device="cuda"), torch.zeros(384, device="cuda"),
               residual=torch.zeros(256, 384, device="cuda"), ln_weight=torch.ones(384, device="cuda"),
               ln_bias=torch.zeros(384, device="cuda"))


def test_residual_relu_bias2d_and_row_mask():
    from vnext_b200.linear import linear
    M, N, K, S = 1000, 384, 256, 250
    x, w, b = _mk(M, N, K, 5)
    g = torch.Generator().manual_seed(6)
    res = torch.randn(M, N, generator=g).cuda()
    b2 = torch.randn(S, N, generator=g).cuda()
    mask = (torch.rand(S, generator=g) < 0.3).to(torch.uint8).cuda()
    ref = F.linear(x.double(), w.double(), b.double())
    tol = 2e-3 * ref.abs().max().item()
    out = linear(x, w, b, relu=True, residual=res)
    assert (out.double() - (ref + res.double()).clamp(min=0)).abs().max().item() < tol
    out = linear(x, w, None, bias2d=b2)
    want = F.linear(x.double(), w.double()) + b2.double().repeat(M // S, 1)
    assert (out.double() - want).abs().max().item() < tol
    out = linear(x, w, b, row_mask=mask, mask_cols=256)
    want = ref.clone()
    want[mask.bool().repeat(M // S).cpu().cuda(), :256] = 0
    assert (out.double() - want).abs().max().item() < tol
    assert (out[mask.bool().repeat(M // S), :256] == 0).all()


@pytest.mark.parametrize("T,H,W,Cin,Cout", [(2, 23, 40, 64, 64), (1, 8, 16, 64, 128), (3, 46, 80, 256, 256), (5, 184, 320, 128, 128),
                                           (2, 13, 17, 128, 8), (1, 12, 20, 2048, 256), (2, 1, 1, 64, 64), (1, 7, 33, 64, 32)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv3x3_stride2(T, H, W, Cin, Cout, dtype):
    """The stride-in-3x3 convs (res3/res4/res5 openers, extra-level input projection): element-strided TMA boxes."""
    from vnext_b200.linear import conv3x3
    g = torch.Generator().manual_seed(T * 1000 + H)
    x = torch.randn(T, H, W, Cin, generator=g).cuda().to(dtype)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).cuda().to(dtype)
    b = torch.randn(Cout, generator=g).cuda()
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    out = conv3x3(x, wk, b, relu=False, stride=2)
    assert out.shape == ref.shape
    tol = (2e-3 if dtype == torch.float32 else 8e-3) * ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() < tol
    out = conv3x3(x, wk, b, relu=True, stride=2)
    assert (out.double() - ref.clamp(min=0)).abs().max().item() < tol


@pytest.mark.parametrize("S,M", [(200, 1000), (136, 1000), (8, 300), (19560, 3 * 19560)])
def test_periodic_bias2d_staged_by_tma(S, M):
    """period % 8 == 0: the row-periodic 2-D bias tile goes through shared memory as four 8-row TMA boxes (a 32-row tile may
    wrap around the period at any multiple of 8); fp32 and bf16 operands, fp32 and bf16 outputs."""
    from vnext_b200.linear import linear
    N, K = 384, 256
    x, w, _ = _mk(M, N, K, 7)
    b2 = torch.randn(S, N, generator=torch.Generator().manual_seed(8)).cuda()
    rep = b2.double().repeat(M // S + 1, 1)[:M]
    want = F.linear(x.double(), w.double()) + rep
    tol = 2e-3 * want.abs().max().item()
    out = linear(x, w, None, bias2d=b2)
    assert (out.double() - want).abs().max().item() < tol
    out16 = linear(x, w, None, bias2d=b2, out_dtype=torch.bfloat16)
    assert (out16.double() - want).abs().max().item() < 4 * tol
    outb = linear(x.bfloat16(), w.bfloat16(), None, bias2d=b2, out_dtype=torch.float32)
    wantb = F.linear(x.bfloat16().double(), w.bfloat16().double()) + rep
    assert (outb.double() - wantb).abs().max().item() < 2e-5 * want.abs().max().item()


@pytest.mark.parametrize("T,H,W,Cin,Cout", [(2, 23, 40, 64, 64), (1, 8, 16, 32, 128), (3, 46, 80, 256, 256),
                                           (1, 184, 320, 64, 64), (2, 13, 17, 128, 8), (5, 92, 160, 256, 64)])
def test_conv3x3_implicit_gemm(T, H, W, Cin, Cout):
    from vnext_b200.linear import conv3x3
    g = torch.Generator().manual_seed(T * 1000 + H)
    x = torch.randn(T, H, W, Cin, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    out = conv3x3(x, wk, b, relu=False)
    tol = 2e-3 * ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() < tol
    out = conv3x3(x, wk, b, relu=True)
    assert (out.double() - ref.clamp(min=0)).abs().max().item() < tol


# ---------------------------------------------------------------------------------------------- bf16 operands / outputs
def _ref_bf16(x, w, b):
    """fp64 reference on the bf16-ROUNDED operands (what the tensor cores see); accumulation is fp32 in TMEM."""
    return F.linear(x.bfloat16().double(), w.bfloat16().double(), None if b is None else b.double())


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (300, 256, 256), (1000, 384, 256), (97800, 256, 256),
                                   (4097, 1024, 256), (513, 256, 1024), (77, 64, 64), (1500, 128, 2048), (2000, 640, 256),
                                   (300, 384, 64), (130, 256, 136)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_bf16_operands(M, N, K, out_dtype):
    from vnext_b200.linear import linear
    x, w, b = _mk(M, N, K, 11)
    ref = _ref_bf16(x, w, b)
    out = linear(x.bfloat16(), w.bfloat16(), b, out_dtype=out_dtype)
    assert out.dtype == out_dtype
    scale = ref.abs().max().item()
    # fp32 out: only accumulation-order error; bf16 out: one output rounding (2^-9 relative)
    tol = 2e-5 * scale if out_dtype == torch.float32 else 5e-3 * scale
    assert (out.double() - ref).abs().max().item() < tol
    out_r = linear(x.bfloat16(), w.bfloat16(), b, relu=True, out_dtype=out_dtype)
    assert (out_r.double() - ref.clamp(min=0)).abs().max().item() < tol


@pytest.mark.parametrize("M,K", [(128, 256), (333, 1024), (97800, 256)])
def test_bf16_residual_layernorm_epilogue(M, K):
    from vnext_b200.linear import linear
    x, w, b = _mk(M, 256, K, 1)
    g = torch.Generator().manual_seed(2)
    res = torch.randn(M, 256, generator=g).cuda().bfloat16()
    gamma = (torch.rand(256, generator=g) + 0.5).cuda()
    beta = (torch.randn(256, generator=g) * 0.1).cuda()
    ref = F.layer_norm(res.double() + _ref_bf16(x, w, b), (256,), gamma.double(), beta.double(), 1e-5)
    out = linear(x.bfloat16(), w.bfloat16(), b, residual=res, ln_weight=gamma, ln_bias=beta)
    assert out.dtype == torch.bfloat16
    assert (out.double() - ref).abs().max().item() < 3e-2          # one bf16 rounding of O(4) outputs
    assert (out.double() - ref).abs().mean().item() < 3e-3
    out32 = linear(x.bfloat16(), w.bfloat16(), b, residual=res.float(), ln_weight=gamma, ln_bias=beta, out_dtype=torch.float32)
    assert (out32.double() - ref).abs().max().item() < 1e-4


def test_bf16_bias2d_row_mask_residual_relu_and_mixed_fp32_input():
    from vnext_b200.linear import linear
    M, N, K, S = 1000, 384, 256, 250
    x, w, b = _mk(M, N, K, 5)
    g = torch.Generator().manual_seed(6)
    res = torch.randn(M, N, generator=g).cuda().bfloat16()
    b2 = torch.randn(S, N, generator=g).cuda()
    mask = (torch.rand(S, generator=g) < 0.3).to(torch.uint8).cuda()
    ref = _ref_bf16(x, w, b)
    scale = ref.abs().max().item()
    out = linear(x.bfloat16(), w.bfloat16(), b, relu=True, residual=res)
    assert (out.double() - (ref + res.double()).clamp(min=0)).abs().max().item() < 8e-3 * scale
    out = linear(x.bfloat16(), w.bfloat16(), None, bias2d=b2, out_dtype=torch.float32)      # the offsets|logits projection
    want = _ref_bf16(x, w, None) + b2.double().repeat(M // S, 1)
    assert (out.double() - want).abs().max().item() < 2e-5 * scale
    out = linear(x.bfloat16(), w.bfloat16(), b, row_mask=mask, mask_cols=256)               # the value projection
    rows = mask.bool().repeat(M // S)
    assert (out[rows, :256] == 0).all() and (out.double()[~rows] - ref[~rows]).abs().max().item() < 5e-3 * scale
    out = linear(x, w, b, out_dtype=torch.bfloat16)                                          # fp32 (TF32) in, bf16 out
    ref32 = F.linear(x.double(), w.double(), b.double())
    assert out.dtype == torch.bfloat16 and (out.double() - ref32).abs().max().item() < 6e-3 * scale


@pytest.mark.parametrize("T,H,W,Cin,Cout", [(2, 23, 40, 64, 64), (1, 8, 16, 64, 128), (3, 46, 80, 256, 256),
                                           (1, 184, 320, 64, 64), (2, 13, 17, 128, 8), (5, 92, 160, 256, 64)])
def test_conv3x3_bf16(T, H, W, Cin, Cout):
    from vnext_b200.linear import conv3x3
    g = torch.Generator().manual_seed(T * 1000 + H)
    x = torch.randn(T, H, W, Cin, generator=g).cuda().bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).cuda().bfloat16()
    b = torch.randn(Cout, generator=g).cuda()
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    out = conv3x3(x, wk, b, relu=False, out_dtype=torch.float32)
    assert (out.double() - ref).abs().max().item() < 2e-5 * scale
    out = conv3x3(x, wk, b, relu=True)
    assert out.dtype == torch.bfloat16 and (out.double() - ref.clamp(min=0)).abs().max().item() < 5e-3 * scale


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (1000, 1024, 256), (97800, 1024, 256), (513, 256, 1024)])
def test_fp16_output_from_tf32_operands_and_saturation(M, N, K):
    """The tf32 mode's FFN1: fp32 operands (TF32 MMAs), output stored as IEEE fp16 (round to nearest, saturating)."""
    from vnext_b200.linear import linear
    x, w, b = _mk(M, N, K, 11)
    ref = F.linear(x.double(), w.double(), b.double()).clamp(min=0)
    out = linear(x, w, b, relu=True, out_dtype=torch.float16)
    assert out.dtype == torch.float16 and out.shape == (M, N)
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() < 2.6e-3 * scale            # TF32 products (2e-3) + one fp16 rounding (2^-11)
    big = linear(x * 1e5, w, b, out_dtype=torch.float16)                        # |values| far above 65504: saturate, never inf
    assert torch.isfinite(big).all() and big.abs().max().item() == 65504.0
    with pytest.raises(RuntimeError, match="fp16 output cannot carry a residual"):
        linear(x, w, b, residual=torch.zeros(M, N, device="cuda", dtype=torch.float16), out_dtype=torch.float16)


@pytest.mark.parametrize("M,K", [(128, 256), (333, 1024), (97800, 1024)])
def test_fp16_operands_residual_layernorm(M, K):
    """The tf32 mode's FFN2: fp16 hidden x fp16 weight (kind::f16, fp32 accumulate) + fp32 residual + LayerNorm -> fp32;
    exact products of the fp16-rounded operands: only accumulation-order error vs the float64 reference on the same inputs."""
    from vnext_b200.linear import linear
    x, w, b = _mk(M, 256, K, 12)
    x = x.clamp(min=0)                                                         # post-ReLU hidden
    g = torch.Generator().manual_seed(13)
    res = torch.randn(M, 256, generator=g).cuda()
    gamma = (torch.rand(256, generator=g) + 0.5).cuda()
    beta = (torch.randn(256, generator=g) * 0.1).cuda()
    xh, wh = x.half(), w.half()
    lin = F.linear(xh.double(), wh.double(), b.double())
    out = linear(xh, wh, b, out_dtype=torch.float32)
    assert (out.double() - lin).abs().max().item() < 2e-5 * lin.abs().max().item()
    ref = F.layer_norm(res.double() + lin, (256,), gamma.double(), beta.double(), 1e-5)
    out = linear(xh, wh, b, residual=res, ln_weight=gamma, ln_bias=beta, out_dtype=torch.float32)
    assert out.dtype == torch.float32 and (out.double() - ref).abs().max().item() < 1e-4
    # and against the all-fp32 operands the TF32 path would have rounded itself: same 2e-3 class
    full = F.layer_norm(res.double() + F.linear(x.double(), w.double(), b.double()), (256,), gamma.double(), beta.double(), 1e-5)
    assert (out.double() - full).abs().max().item() < 4e-3
