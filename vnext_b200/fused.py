"""Python face of the single-pass NHWC kernels (include/vnext_b200.h: vnext_bias_act, vnext_stem_pool,
vnext_groupnorm_nhwc, vnext_preprocess).  CUDA only."""
import ctypes
import os

import torch

from . import _cabi


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


_DT = {torch.float32: _cabi.F32, torch.bfloat16: _cabi.BF16}


def bias_act(x, bias, relu=True, residual=None, out=None):
    """x: [..., C] contiguous (NHWC); returns act(x + bias (+ residual)); in place when out is x."""
    C = x.shape[-1]
    assert x.is_contiguous() and (residual is None or residual.is_contiguous())
    out = torch.empty_like(x) if out is None else out
    rows = x.numel() // C
    assert out.dtype == x.dtype and (residual is None or residual.dtype == x.dtype)
    info = {"bytes": float(x.element_size()) * x.numel() * (3 if residual is not None else 2)}
    rc = _cabi.timed_call("bias_act", info, x, lambda: _cabi.lib().vnext_bias_act_dt(
        _DT[x.dtype], x.data_ptr(), bias.data_ptr(), None if residual is None else residual.data_ptr(), out.data_ptr(), rows, C,
        int(relu), _st(x)))
    _cabi.check(rc, "vnext_bias_act")
    return out


def stem_pool(x, bias, out_dtype=None):
    """x: [N, H, W, C] conv output (NHWC contiguous) -> maxpool3x3s2p1(relu(x + bias)) [N, Ho, Wo, C]."""
    N, H, W, C = x.shape
    assert x.is_contiguous()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty((N, Ho, Wo, C), dtype=out_dtype or x.dtype, device=x.device)
    info = {"bytes": float(x.element_size() * x.numel() + out.element_size() * out.numel())}
    rc = _cabi.timed_call("stem_pool", info, x, lambda: _cabi.lib().vnext_stem_pool_dt(
        _DT[x.dtype], _DT[out.dtype], x.data_ptr(), bias.data_ptr(), out.data_ptr(), N, H, W, C, _st(x)))
    _cabi.check(rc, "vnext_stem_pool")
    return out


def groupnorm_nhwc(x, gamma, beta, groups=32, eps=1e-5, out=None, stats_ws=None):
    """x: [N, HW, C] contiguous; out (optional): a [N, HW, C] view whose batch stride may exceed HW*C."""
    N, HW, C = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    assert out.stride(2) == 1 and out.stride(1) == C
    if stats_ws is None:
        stats_ws = torch.empty(_cabi.lib().vnext_groupnorm_workspace_floats(N, HW, groups), dtype=torch.float32, device=x.device)
    info = {"bytes": float(x.element_size() * x.numel() * 2 + out.element_size() * x.numel())}
    rc = _cabi.timed_call("groupnorm_nhwc", info, x, lambda: _cabi.lib().vnext_groupnorm_nhwc_dt(
        _DT[x.dtype], _DT[out.dtype], x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), out.stride(0),
        stats_ws.data_ptr(), N, HW, C, groups, float(eps), _st(x)))
    _cabi.check(rc, "vnext_groupnorm_nhwc")
    return out


def preprocess(frames, mean, std, Hp, Wp, out=None, out_dtype=torch.float32):
    """frames [T,3,H,W] uint8 or float32 on device -> normalised zero-padded NHWC [T,Hp,Wp,3] (fp32 or bf16)."""
    T, _, H, W = frames.shape
    assert frames.is_contiguous() and frames.dtype in (torch.uint8, torch.float32)
    if out is None:
        out = torch.empty((T, Hp, Wp, 3), dtype=out_dtype, device=frames.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    info = {"bytes": float(frames.numel() * frames.element_size() + out.numel() * out.element_size())}
    rc = _cabi.timed_call("preprocess", info, frames, lambda: _cabi.lib().vnext_preprocess_dt(
        int(frames.dtype == torch.uint8), _DT[out.dtype], frames.data_ptr(), out.data_ptr(), T, H, W, Hp, Wp,
        ctypes.cast(m, ctypes.c_void_p), ctypes.cast(s, ctypes.c_void_p), _st(frames)))
    _cabi.check(rc, "vnext_preprocess")
    return out


def stem_conv7x7(frames, weight_packed, mean, std, Hp, Wp, out_dtype=torch.float32, bias=None, relu=False):
    """frames [T,3,H,W] uint8 / float32 (0..255) on the device -> conv7x7/s2/p3 of the normalised, zero-padded frames:
    [T, Hp/2, Wp/2, Cout] channels-last (tensor-core implicit GEMM, csrc/gemm_tcgen05.cu stem_conv_device).
    weight_packed [Cout, 7*8*4] fp32 (see pack_stem_weight)."""
    T, _, H, W = frames.shape
    Cout = weight_packed.shape[0]
    assert frames.is_contiguous() and frames.dtype in (torch.uint8, torch.float32) and weight_packed.is_contiguous()
    out = torch.empty((T, Hp // 2, Wp // 2, Cout), dtype=out_dtype, device=frames.device)
    scratch = torch.empty((T, Hp + 6, Wp + 6, 4), dtype=torch.float32, device=frames.device)
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    sd = (ctypes.c_float * 3)(*[float(v) for v in std])
    info = {"bytes": float(frames.numel() * frames.element_size() + 2 * scratch.numel() * 4 + out.numel() * out.element_size()),
            "flops": 2.0 * out.numel() * 147}
    rc = _cabi.timed_call("stem_conv7x7", info, frames, lambda: _cabi.lib().vnext_stem_conv7x7(
        int(frames.dtype == torch.uint8), _DT[out.dtype], frames.data_ptr(), weight_packed.data_ptr(), _ptr_or_none(bias),
        out.data_ptr(), scratch.data_ptr(), T, H, W, Hp, Wp, Cout, ctypes.cast(m, ctypes.c_void_p), ctypes.cast(sd, ctypes.c_void_p),
        int(relu), _st(frames)))
    _cabi.check(rc, "vnext_stem_conv7x7")
    return out


def pack_stem_weight(w):
    """[Cout, 3, 7, 7] conv weight -> [Cout, 7 * 8 * 4] (ky, kx padded 7 -> 8, channel padded 3 -> 4 with zeros), fp32."""
    Cout = w.shape[0]
    p = torch.zeros((Cout, 7, 8, 4), dtype=torch.float32, device=w.device)
    p[:, :, :7, :3] = w.float().permute(0, 2, 3, 1)
    return p.reshape(Cout, 7 * 8 * 4).contiguous()


def _ptr_or_none(t):
    return None if t is None else t.data_ptr()


def layernorm(x, gamma, beta, eps=1e-5, out=None, out_dtype=None):
    """Row LayerNorm over the last dimension (contiguous rows); x / out fp32 or bf16, statistics fp32."""
    C = x.shape[-1]
    assert x.is_contiguous()
    out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device) if out is None else out
    rows = x.numel() // C
    info = {"bytes": float(x.element_size() * x.numel() + out.element_size() * x.numel())}
    rc = _cabi.timed_call("layernorm", info, x, lambda: _cabi.lib().vnext_layernorm_dt(
        _DT[x.dtype], _DT[out.dtype], x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), rows, C, float(eps), _st(x)))
    _cabi.check(rc, "vnext_layernorm")
    return out


def swin_window_attention(qkv, qkv_bias, table, heads, window, shift, tensor_cores=False):
    """qkv [B,H,W,3C] contiguous -> attention output [B,H,W,C] (input of the block's proj Linear), same dtype as qkv.
    tensor_cores: tcgen05 TF32 kernel (csrc/swin_attention_tc.cu; fp32 or bf16 storage) instead of the exact-fp32 SIMT kernel."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    assert qkv.is_contiguous() and table.is_contiguous()
    out = torch.empty((B, H, W, C), dtype=qkv.dtype, device=qkv.device)
    n = window * window
    e = qkv.element_size()
    info = {"bytes": float(e) * (qkv.numel() + out.numel()), "flops": 4.0 * B * (-(-H // window)) * (-(-W // window)) * heads * n * n * 32}
    tc = tensor_cores and window >= 4
    if qkv.dtype == torch.bfloat16 and os.environ.get("VNEXT_SWIN_BF16_KERNEL", "mma") == "mma":
        # bf16 storage: the register-resident mma.sync kernel (csrc/swin_attention_mma.cu)
        tt = getattr(table, "_vnext_t", None)                     # [heads, (2ws-1)^2]: transposed once per table tensor
        if tt is None or tt[0] != table._version:
            tt = (table._version, table.t().contiguous())
            table._vnext_t = tt
        rc = _cabi.timed_call("swin_window_attention_bf16", info, qkv, lambda: _cabi.lib().vnext_swin_window_attention_bf16(
            qkv.data_ptr(), qkv_bias.data_ptr(), tt[1].data_ptr(), out.data_ptr(), B, H, W, C, heads, window, shift, _st(qkv)))
        _cabi.check(rc, "vnext_swin_window_attention_bf16")
        return out
    if not tc and qkv.dtype != torch.float32:
        raise RuntimeError("swin_window_attention: bf16 storage needs a tensor-core kernel")
    if tc:
        call = lambda: _cabi.lib().vnext_swin_window_attention_tc_dt(
            _DT[qkv.dtype], qkv.data_ptr(), qkv_bias.data_ptr(), table.data_ptr(), out.data_ptr(), B, H, W, C, heads, window, shift, _st(qkv))
    else:
        call = lambda: _cabi.lib().vnext_swin_window_attention(
            qkv.data_ptr(), qkv_bias.data_ptr(), table.data_ptr(), out.data_ptr(), B, H, W, C, heads, window, shift, _st(qkv))
    rc = _cabi.timed_call("swin_window_attention_tc" if tc else "swin_window_attention", info, qkv, call)
    _cabi.check(rc, "vnext_swin_window_attention")
    return out


def time_pool_ln(out_in, t2, logit, gamma, beta, eps=1e-5):
    """out_in [1,nq,C], t2 [T,nq,C], logit [T,nq,1] -> LayerNorm(out_in + sum_t softmax_t(logit) t2) [1,nq,C] (fp32)."""
    T, nq, C = t2.shape
    assert out_in.is_contiguous() and t2.is_contiguous() and logit.is_contiguous() and t2.dtype == torch.float32
    out = torch.empty_like(out_in)
    rc = _cabi.timed_call("time_pool_ln", {"bytes": 4.0 * (t2.numel() + 2 * out.numel())}, t2, lambda: _cabi.lib().vnext_time_pool_ln(
        out_in.data_ptr(), t2.data_ptr(), logit.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), T, nq, C,
        float(eps), _st(t2)))
    _cabi.check(rc, "vnext_time_pool_ln")
    return out


def box_refine(delta, ref):
    """delta [..., 4], ref [..., 2|4] (same leading shape) -> sigmoid(delta + pad(inverse_sigmoid(ref))) [..., 4]."""
    delta, ref = delta.contiguous(), ref.contiguous()
    out = torch.empty_like(delta)
    rows = delta.numel() // 4
    rc = _cabi.timed_call("box_refine", {"bytes": 4.0 * (2 * delta.numel() + ref.numel())}, delta, lambda: _cabi.lib().vnext_box_refine(
        delta.data_ptr(), ref.data_ptr(), out.data_ptr(), rows, ref.shape[-1], _st(delta)))
    _cabi.check(rc, "vnext_box_refine")
    return out


def ref_points_px(ref, scale_w, scale_h):
    """ref [..., 2|4] in [0, 1] -> sigmoid(inverse_sigmoid(ref[..., :2])) * (scale_w, scale_h)  [..., 2]."""
    ref = ref.contiguous()
    out = torch.empty(ref.shape[:-1] + (2,), dtype=torch.float32, device=ref.device)
    rows = out.numel() // 2
    rc = _cabi.timed_call("ref_points_px", {"bytes": 4.0 * (ref.numel() + out.numel())}, ref, lambda: _cabi.lib().vnext_ref_points_px(
        ref.data_ptr(), out.data_ptr(), rows, ref.shape[-1], float(scale_w), float(scale_h), _st(ref)))
    _cabi.check(rc, "vnext_ref_points_px")
    return out


def small_mha(qkv, heads):
    """qkv [B, n, 3*heads*32] fp32 (rows may be strided) -> softmax(q k^T / sqrt(32)) v, heads merged: [B, n, heads*32]."""
    B, n, C3 = qkv.shape
    C = heads * 32
    assert C3 >= 3 * C and qkv.dtype == torch.float32 and qkv.stride(2) == 1 and qkv.stride(0) == n * qkv.stride(1)
    out = torch.empty((B, n, C), dtype=torch.float32, device=qkv.device)
    info = {"bytes": 4.0 * (qkv.numel() + out.numel()), "flops": 4.0 * B * heads * n * n * 32}
    rc = _cabi.timed_call("mha_small", info, qkv, lambda: _cabi.lib().vnext_mha_small(
        qkv.data_ptr(), qkv.stride(1), out.data_ptr(), C, B, n, heads, _st(qkv)))
    _cabi.check(rc, "vnext_mha_small")
    return out


def dynamic_mask_head(feats_nhwc, params, ref_px, stride=8):
    """feats_nhwc [T,H,W,8]; params [n,169] or [T,n,169]; ref_px [T,n,2] -> stride-8 logits [n,T,H,W]."""
    T, H, W, C = feats_nhwc.shape
    assert C == 8 and feats_nhwc.is_contiguous()
    per_frame = params.dim() == 3
    n = params.shape[-2]
    params, ref_px = params.contiguous(), ref_px.contiguous()
    out = torch.empty((n, T, H, W), dtype=torch.float32, device=feats_nhwc.device)
    info = {"bytes": 4.0 * (out.numel() + feats_nhwc.numel()), "flops": 304.0 * out.numel()}
    rc = _cabi.timed_call("dynamic_mask_head", info, feats_nhwc, lambda: _cabi.lib().vnext_dynamic_mask_head(
        feats_nhwc.data_ptr(), params.data_ptr(), int(per_frame), ref_px.data_ptr(), out.data_ptr(), T, n, H, W, stride,
        _st(feats_nhwc)))
    _cabi.check(rc, "vnext_dynamic_mask_head")
    return out


def mask_finalize(logits8, frame_hw, out_hw, mask_stride=4):
    """logits8 [n,T,H8,W8] -> bool masks [n,T,OH,OW] (upsample x2 aligned, x mask_stride bilinear, sigmoid, crop, resize, >0.5)."""
    n, T, H8, W8 = logits8.shape
    assert logits8.is_contiguous()
    out = torch.empty((n, T, out_hw[0], out_hw[1]), dtype=torch.uint8, device=logits8.device)
    info = {"bytes": float(out.numel() + 4 * logits8.numel())}
    rc = _cabi.timed_call("mask_finalize", info, logits8, lambda: _cabi.lib().vnext_mask_finalize(
        logits8.data_ptr(), out.data_ptr(), n * T, H8, W8, frame_hw[0], frame_hw[1], out_hw[0], out_hw[1], mask_stride,
        _st(logits8)))
    _cabi.check(rc, "vnext_mask_finalize")
    return out.view(torch.bool)
