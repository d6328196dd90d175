"""Linear layers on the hand-written tcgen05 GEMM (include/vnext_b200.h: vnext_linear_tf32).

`linear(x, weight, bias, ...)` has F.linear's argument meaning (weight is [out, in]) plus the fused
epilogues the backbone / encoder / decoder layers need: ReLU, residual add, LayerNorm
(deformable_transformer.py:155-169: `src = norm(src + dropout(linear(...)))`; resnet.py:194-210:
`relu(conv3(out) + shortcut)`), a row-periodic 2-D bias and a row mask (see the header).
"""
import torch

from . import _cabi

EPI_RELU, EPI_RESIDUAL, EPI_LAYERNORM, EPI_GELU = 1, 2, 4, 8


def _ptr(t):
    return None if t is None else t.data_ptr()


def _rows(t, n):
    t2 = t.reshape(-1, n)
    return t2 if t2.stride(-1) == 1 else t2.contiguous()


_DT = {torch.float32: _cabi.F32, torch.bfloat16: _cabi.BF16, torch.float16: _cabi.F16}


def linear(x, weight, bias=None, relu=False, gelu=False, residual=None, ln_weight=None, ln_bias=None, ln_eps=1e-5, out=None,
           bias2d=None, row_mask=None, mask_cols=0, name=None, _debug_flags=0, out_dtype=None):
    """x and weight share a dtype (fp32 -> TF32 MMAs, bf16 / fp16 -> kind::f16 MMAs); out_dtype (default: x's) is also the dtype
    the residual is read in (an fp16 output takes no residual); bias / bias2d / LayerNorm parameters are fp32."""
    if not x.is_cuda:
        raise RuntimeError("vnext_b200.linear: CUDA tensors only (no CPU fallback)")
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = _rows(x, K)
    M = x2.shape[0]
    if x.dtype != weight.dtype or x.dtype not in _DT:
        raise RuntimeError(f"vnext_b200.linear: x ({x.dtype}) and weight ({weight.dtype}) must both be fp32, both bf16 or both fp16")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or x.dtype, device=x.device)
    if residual is not None and residual.dtype != out.dtype:
        residual = residual.to(out.dtype)
    out2 = out.view(-1, out.shape[-1])[:, :N] if out.shape[-1] != N else out.view(-1, N)
    flags = (EPI_RELU if relu else 0) | (EPI_GELU if gelu else 0) | _debug_flags
    res2 = None
    if residual is not None:
        res2 = _rows(residual, N)
        flags |= EPI_RESIDUAL
    if ln_weight is not None:
        flags |= EPI_LAYERNORM
    period = 0
    if bias2d is not None:
        period = bias2d.shape[0]
    elif row_mask is not None:
        period = row_mask.shape[0]
    ie, oe = x.element_size(), out.element_size()
    info = {"flops": 2.0 * M * N * K,
            "bytes": float(ie * (M * K + N * K) + oe * M * N * (2 if residual is not None else 1)
                           + (4 * period * N if bias2d is not None else 0))}     # the periodic bias is read from HBM once
    tag = {torch.float32: "tf32", torch.bfloat16: "bf16", torch.float16: "f16"}[x.dtype]
    if out.dtype == torch.float16 and x.dtype != torch.float16:
        tag += "_of16"
    with torch.cuda.device(x.device):
        rc = _cabi.timed_call(
            name or f"linear_{tag}_m{M}_n{N}_k{K}_f{flags}", info, x,
            lambda: _cabi.lib().vnext_linear_tc(
                _DT[x.dtype], _DT[out.dtype], x2.data_ptr(), x2.stride(0), weight.data_ptr(), _ptr(bias), _ptr(bias2d), period,
                0 if bias2d is None else bias2d.stride(0), _ptr(row_mask), int(mask_cols), _ptr(res2),
                0 if res2 is None else res2.stride(0), _ptr(ln_weight), _ptr(ln_bias), float(ln_eps), out2.data_ptr(),
                out2.stride(0), M, N, K, flags, torch.cuda.current_stream(x.device).cuda_stream))
    _cabi.check(rc, "vnext_linear_tc")
    return out2.view(*x.shape[:-1], N) if out2.is_contiguous() else out2


def conv3x3(x, weight_khwc, bias, relu=True, out_dtype=None, stride=1):
    """x [T,H,W,Cin] contiguous NHWC; weight_khwc [Cout, 9*Cin] (= conv weight.permute(0,2,3,1)); stride 1 or 2, padding 1.
    x / weight fp32 (TF32 MMAs) or bf16; output dtype defaults to x's."""
    T, H, W, Cin = x.shape
    Cout = weight_khwc.shape[0]
    assert x.is_contiguous() and weight_khwc.is_contiguous() and x.dtype == weight_khwc.dtype
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty((T, Ho, Wo, Cout), dtype=out_dtype or x.dtype, device=x.device)
    info = {"flops": 2.0 * T * Ho * Wo * Cout * 9 * Cin,
            "bytes": float(x.element_size() * (x.numel() + weight_khwc.numel()) + out.element_size() * out.numel())}
    tag = "tf32" if x.dtype == torch.float32 else "bf16"
    with torch.cuda.device(x.device):
        rc = _cabi.timed_call(f"conv3x3_{tag}_c{Cin}_n{Cout}" + ("_s2" if stride == 2 else ""), info, x,
                              lambda: _cabi.lib().vnext_conv3x3_strided_tc(
            _DT[x.dtype], _DT[out.dtype], x.data_ptr(), weight_khwc.data_ptr(), _ptr(bias), out.data_ptr(), T, H, W, Cin,
            Cout, int(stride), int(relu), torch.cuda.current_stream(x.device).cuda_stream))
    _cabi.check(rc, "vnext_conv3x3_strided_tc")
    return out


def linear_msda_front(x, weight, bias2d, reference_points, level_hw):
    """The encoder's offsets | logits projection with the sampler front-end in the epilogue (vnext_linear_msda_front):
    x [..., K] (fp32 or bf16, weight same dtype [384, K]); bias2d [S, 384] fp32 (row-periodic: (pos W^T + b) per pixel);
    reference_points [S, 4, 2] fp32; level_hw = 4 x (H, W).  -> [..., 384] fp32: sampling locations | softmax weights."""
    import ctypes
    K = x.shape[-1]
    x2 = _rows(x, K)
    M = x2.shape[0]
    assert weight.shape == (384, K) and x.dtype == weight.dtype and bias2d.shape[1] == 384 and len(level_hw) == 4
    ref = reference_points.reshape(-1, 8)
    assert ref.is_contiguous() and ref.shape[0] == bias2d.shape[0] and ref.dtype == torch.float32
    out = torch.empty((M, 384), dtype=torch.float32, device=x.device)
    wh = (ctypes.c_float * 8)(*[float(v) for hw in level_hw for v in (hw[1], hw[0])])
    ie = x.element_size()
    info = {"flops": 2.0 * M * 384 * K, "bytes": float(ie * (M * K + 384 * K) + 4 * M * 384 + 4 * bias2d.shape[0] * 384 +
                                                       32 * bias2d.shape[0])}   # x, W, out, periodic bias, reference points
    tag = "tf32" if x.dtype == torch.float32 else "bf16"
    with torch.cuda.device(x.device):
        rc = _cabi.timed_call(f"linear_{tag}_m{M}_n384_k{K}_msda_front", info, x, lambda: _cabi.lib().vnext_linear_msda_front(
            _DT[x.dtype], x2.data_ptr(), x2.stride(0), weight.data_ptr(), bias2d.data_ptr(), bias2d.shape[0], bias2d.stride(0),
            ref.data_ptr(), ctypes.cast(wh, ctypes.c_void_p), out.data_ptr(), 384, M, K,
            torch.cuda.current_stream(x.device).cuda_stream))
    _cabi.check(rc, "vnext_linear_msda_front")
    return out.view(*x.shape[:-1], 384)
