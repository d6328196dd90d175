"""Multi-scale deformable attention operator — host side.

Mirrors, name for name, the reference's operator layer:
  * pybind module functions ``ms_deform_attn_forward`` / ``ms_deform_attn_backward``
    (projects/IDOL/idol/models/ops/src/vision.cpp:13-16, src/ms_deform_attn.h:20-61,
     host checks of src/cuda/ms_deform_attn_cuda.cu:28-52, 93-117)
  * ``MSDeformAttnFunction`` (functions/ms_deform_attn_func.py:21-39)
on top of the C ABI (include/vnext_b200.h).  Differences, all supersets:
  * bf16 ``value`` is accepted (sampling_locations / attention_weights are used in fp32);
  * the whole batch is one launch (``im2col_step`` is validated exactly as the reference does and
    otherwise ignored: chunking has no numerical effect);
  * a leading frame axis can be folded into the batch by the caller (see modules.ms_deform_attn).
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _cabi

_DTYPES = {torch.float32: _cabi.F32, torch.float64: _cabi.F64, torch.bfloat16: _cabi.BF16}


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    named = (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight))
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")          # ms_deform_attn.h:38
    for name, t in named:
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
    for name, t in named:
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
    _require(value.dtype in _DTYPES, f"ms_deform_attn: unsupported dtype {value.dtype}")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes and level_start_index must be int64 tensors")
    _require(value.dim() == 4 and sampling_loc.dim() == 6 and attn_weight.dim() == 5,
             "expected value [N,S,M,D], sampling_loc [N,Lq,M,L,P,2], attn_weight [N,Lq,M,L,P]")
    batch = value.size(0)
    step = min(batch, int(im2col_step))
    _require(batch == 0 or (step > 0 and batch % step == 0),
             f"batch({batch}) must divide im2col_step({step})")   # ms_deform_attn_cuda.cu:52
    N, S, M, D = value.shape
    L = spatial_shapes.size(0)
    Lq, P = sampling_loc.size(1), sampling_loc.size(4)
    _require(tuple(sampling_loc.shape) == (N, Lq, M, L, P, 2) and tuple(attn_weight.shape) == (N, Lq, M, L, P),
             "sampling_loc / attn_weight shapes do not match value / spatial_shapes")
    return N, S, M, D, L, Lq, P


# ---------------------------------------------------------------- self-attention (encoder shape) kernels
# When the queries are the pixel grid (Lq == S: encoder self-attention) two more kernels exist besides the tile kernel,
# both exact for ANY sampling locations (include/vnext_b200.h: vnext_msda_forward_selfattn):
#   "cached": cache-blocked (csrc/msda_cached.cu; DEFAULT) — 2-D query tiles x one head per work unit, 32 warps/SM, the L1
#             keeps what neighbouring queries share (the tile kernel re-fetches it from L2: 5 GB of L2 sectors per 720p
#             launch, 1 GB here).  Never slower than the tile kernel over profiles/r02_msda_encoder_kernels.jsonl (8 shapes
#             incl. uniformly random locations): 352 vs 387 us fp32, 295 vs 309 us bf16 value at 720p x 5 frames;
#   "auto":   shared-memory windows staged by TMA (csrc/msda_window.cu).  Removes the L2 traffic too, but at 16 warps/SM it
#             is bound by shared-memory wavefronts / issue slots: 433 us (bf16 value) / 600 us (fp32) at 720p against
#             311 / 387 us for the tile kernel (profiles/r02_msda_window_*): kept as an opt-in;
#   "off":    always the tile kernel.
WINDOW_MODE = "cached"
WINDOW_MARGIN = {torch.float32: 4, torch.bfloat16: 6}      # pixels around the tile footprint staged per level


def host_geometry(spatial_shapes, level_start_index):
    """HOST copies (numpy int64) of the two geometry tensors.  The copy is kept ON the shapes tensor object itself
    (an attribute, validated against both tensors' in-place version counters and the identity of the index tensor), so a
    recycled device address can never serve another tensor's geometry; one device sync per new tensor (the engines pass
    their plan's host copies instead and never hit this)."""
    tag = (spatial_shapes._version, id(level_start_index), level_start_index._version)
    hit = getattr(spatial_shapes, "_vnext_host_geom", None)
    if hit is None or hit[0] != tag:
        if spatial_shapes.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("ms_deform_attn: spatial_shapes must be known on the host before CUDA-graph capture "
                               "(call msda.host_geometry(shapes, lsi) once outside the capture)")
        hit = (tag, spatial_shapes.cpu().numpy().astype("int64").copy(), level_start_index.cpu().numpy().astype("int64").copy(),
               level_start_index)          # keeps the index tensor alive so its id() stays unique
        spatial_shapes._vnext_host_geom = hit
    return hit[1], hit[2]


def attach_host_geometry(spatial_shapes, level_start_index, level_hw):
    """A caller that built the two geometry tensors from host integers registers them here: no device sync later."""
    import numpy as np
    hs = np.asarray(level_hw, dtype=np.int64).reshape(-1, 2).copy()
    starts = np.concatenate(([0], np.cumsum(hs[:, 0] * hs[:, 1])[:-1])).astype(np.int64)
    tag = (spatial_shapes._version, id(level_start_index), level_start_index._version)
    spatial_shapes._vnext_host_geom = (tag, hs, starts, level_start_index)


def _selfattn_margin(value):
    return -1 if WINDOW_MODE == "cached" else WINDOW_MARGIN[value.dtype]


def window_supported(value, S, Lq, L, P):
    return (WINDOW_MODE != "off" and Lq == S and value.dtype in WINDOW_MARGIN and value.shape[-1] == 32 and 1 <= L <= 4
            and L * P in (8, 16, 32) and value.numel() < 2 ** 31)


def _pyramid_in_order(hs, hl, S):
    """The window kernel reads a query's pixel from its index: levels must tile [0, S) back to back."""
    start = 0
    for (h, w), s0 in zip(hs.tolist(), hl.tolist()):
        if h <= 0 or w <= 0 or s0 != start:
            return False
        start += h * w
    return start == S


def _loc_dtype(value):
    return torch.float64 if value.dtype == torch.float64 else torch.float32


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    N, S, M, D, L, Lq, P = _check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                         im2col_step)
    ldt = _loc_dtype(value)
    if sampling_loc.dtype != ldt:
        sampling_loc = sampling_loc.to(ldt)
    if attn_weight.dtype != ldt:
        attn_weight = attn_weight.to(ldt)
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    e = value.element_size()
    info = {"bytes": float(N) * (e * S * M * D + sampling_loc.element_size() * 3 * Lq * M * L * P + e * Lq * M * D),
            "N": N, "Lq": Lq, "S": S}
    hs = hl = None
    if window_supported(value, S, Lq, L, P):
        hs, hl = host_geometry(spatial_shapes, level_start_index)
    if hs is not None and _pyramid_in_order(hs, hl, S):
        with torch.cuda.device(value.device):
            rc = _cabi.timed_call(
                "msda_cached_enc" if WINDOW_MODE == "cached" else "msda_win_enc", info, value,
                lambda: _cabi.lib().vnext_msda_forward_selfattn(
                    _DTYPES[value.dtype], value.data_ptr(), hs.ctypes.data, hl.ctypes.data, sampling_loc.data_ptr(),
                    attn_weight.data_ptr(), None, 0, None, 0, 0, N, S, M, D, L, P, _selfattn_margin(value),
                    out.data_ptr(), _stream(value)))
        _cabi.check(rc, "ms_deform_attn_forward (window)")
        return out
    with torch.cuda.device(value.device):
        rc = _cabi.timed_call(
            "msda_fwd_enc" if Lq == S else "msda_fwd_dec", info, value,
            lambda: _cabi.lib().vnext_msda_forward(
                _DTYPES[value.dtype], value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr(),
                _stream(value)))
    _cabi.check(rc, "ms_deform_attn_forward")
    return out


def ms_deform_attn_fused(value, spatial_shapes, level_start_index, offsets_logits, reference_points, n_levels, n_points,
                         host_geom=None, margin=None):
    """Fused module core: softmax + sampling locations + sampling (include/vnext_b200.h: vnext_msda_forward_fused).
    value [N,S,M,32]; offsets_logits [N,Lq,M*L*P*3] (rows may be strided); reference_points [N|1, Lq, L, 2|4].
    Lq == S (encoder self-attention) runs the shared-memory window kernel (vnext_msda_forward_selfattn); `host_geom` =
    (shapes, level_start_index) as host int64 numpy arrays (else they are fetched once from the device and cached)."""
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    N, S, M, D = value.shape
    Lq = offsets_logits.shape[1]
    L, P = n_levels, n_points
    _require(value.is_contiguous() and offsets_logits.stride(-1) == 1 and reference_points.is_contiguous(),
             "value / reference_points must be contiguous and offsets_logits row-contiguous")
    _require(offsets_logits.dtype == torch.float32 and reference_points.dtype == torch.float32,
             "offsets_logits and reference_points must be float32")
    ol = offsets_logits.reshape(N * Lq, offsets_logits.shape[-1])
    ref_shared = int(reference_points.shape[0] == 1 and N > 1) if reference_points.dim() == 4 else 0
    _require(reference_points.shape[-3:-1] == (Lq, L) and reference_points.shape[0] in (1, N),
             "reference_points must be [N|1, Lq, L, 2|4]")
    out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
    e = value.element_size()
    info = {"bytes": float(N) * (e * S * M * D + 4 * 3 * Lq * M * L * P + e * Lq * M * D), "N": N, "Lq": Lq, "S": S}
    hs = hl = None
    if window_supported(value, S, Lq, L, P) and reference_points.shape[-1] == 2:
        hs, hl = host_geom if host_geom is not None else host_geometry(spatial_shapes, level_start_index)
    if hs is not None and _pyramid_in_order(hs, hl, S):
        mg = _selfattn_margin(value) if margin is None else margin
        with torch.cuda.device(value.device):
            rc = _cabi.timed_call(
                "msda_cached_fused_enc" if WINDOW_MODE == "cached" else "msda_win_fused_enc", info, value,
                lambda: _cabi.lib().vnext_msda_forward_selfattn(
                    _DTYPES[value.dtype], value.data_ptr(), hs.ctypes.data, hl.ctypes.data, None, None, ol.data_ptr(),
                    ol.stride(0), reference_points.data_ptr(), reference_points.shape[-1],
                    1 if reference_points.shape[0] == 1 else 0, N, S, M, D, L, P, mg, out.data_ptr(), _stream(value)))
        _cabi.check(rc, "ms_deform_attn_fused (window)")
        return out
    with torch.cuda.device(value.device):
        rc = _cabi.timed_call(
            "msda_fused_enc" if Lq == S else "msda_fused_dec", info, value,
            lambda: _cabi.lib().vnext_msda_forward_fused(
                _DTYPES[value.dtype], value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                ol.data_ptr(), ol.stride(0), reference_points.data_ptr(), reference_points.shape[-1],
                1 if reference_points.shape[0] == 1 else 0, N, S, M, D, L, Lq, P, out.data_ptr(), _stream(value)))
    _cabi.check(rc, "ms_deform_attn_fused")
    return out


def ms_deform_attn_located(value, loc_attn, n_levels, n_points, host_geom):
    """Self-attention sampling from rows that already hold [sampling locations (M*L*P*2) | normalised weights (M*L*P)] — the
    output of linear.linear_msda_front — through the cache-blocked kernel (no front-end arithmetic in the sampler).
    value [N,S,M,32]; loc_attn [N,S,M*L*P*3] fp32; host_geom = (shapes, level_start_index) host int64 arrays."""
    N, S, M, D = value.shape
    L, P = n_levels, n_points
    _require(value.is_cuda and value.is_contiguous() and loc_attn.stride(-1) == 1 and loc_attn.dtype == torch.float32 and
             loc_attn.shape[1] == S and D == 32 and value.dtype in WINDOW_MARGIN, "ms_deform_attn_located: bad arguments")
    hs, hl = host_geom
    _require(_pyramid_in_order(hs, hl, S), "ms_deform_attn_located: levels must tile [0, S) in order")
    la = loc_attn.reshape(N * S, loc_attn.shape[-1])
    out = torch.empty((N, S, M * D), dtype=value.dtype, device=value.device)
    e = value.element_size()
    info = {"bytes": float(N) * (e * S * M * D + 4 * 3 * S * M * L * P + e * S * M * D), "N": N, "Lq": S, "S": S}
    with torch.cuda.device(value.device):
        rc = _cabi.timed_call("msda_cached_located_enc", info, value, lambda: _cabi.lib().vnext_msda_forward_selfattn(
            _DTYPES[value.dtype], value.data_ptr(), hs.ctypes.data, hl.ctypes.data, None, None, la.data_ptr(), la.stride(0),
            None, 0, 0, N, S, M, D, L, P, -1, out.data_ptr(), _stream(value)))
    _cabi.check(rc, "ms_deform_attn_located")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    N, S, M, D, L, Lq, P = _check_common(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                         im2col_step)
    _require(grad_output.is_contiguous(), "grad_output tensor has to be contiguous")
    _require(grad_output.is_cuda, "grad_output must be a CUDA tensor")
    _require(grad_output.dtype == value.dtype and grad_output.numel() == N * Lq * M * D,
             "grad_output must match the forward output")
    ldt = _loc_dtype(value)
    loc_in, attn_in = sampling_loc.to(ldt), attn_weight.to(ldt)
    grad_value = torch.empty_like(value)
    grad_loc = torch.empty(sampling_loc.shape, dtype=ldt, device=value.device)
    grad_attn = torch.empty(attn_weight.shape, dtype=ldt, device=value.device)
    ws = None
    ws_ptr = None
    if value.dtype == torch.bfloat16:
        ws = torch.empty(value.shape, dtype=torch.float32, device=value.device)
        ws_ptr = ws.data_ptr()
    e = value.element_size()
    info = {"bytes": float(N) * (e * (2 * S * M * D + Lq * M * D) + loc_in.element_size() * 6 * Lq * M * L * P),
            "N": N, "Lq": Lq, "S": S}                  # SURVEY §8d bwd_bytes
    with torch.cuda.device(value.device):
        rc = _cabi.timed_call(
            "msda_bwd_enc" if Lq == S else "msda_bwd_dec", info, value,
            lambda: _cabi.lib().vnext_msda_backward(
                _DTYPES[value.dtype], value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                loc_in.data_ptr(), attn_in.data_ptr(), grad_output.data_ptr(), N, S, M, D, L, Lq, P,
                grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(), ws_ptr, _stream(value)))
    _cabi.check(rc, "ms_deform_attn_backward")
    if grad_loc.dtype != sampling_loc.dtype:
        grad_loc = grad_loc.to(sampling_loc.dtype)
    if grad_attn.dtype != attn_weight.dtype:
        grad_attn = grad_attn.to(attn_weight.dtype)
    return [grad_value, grad_loc, grad_attn]


class MSDeformAttnFunction(Function):
    """Same call signature and gradient contract as functions/ms_deform_attn_func.py:21-39."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                        attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        grad_value, grad_loc, grad_attn = ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_output,
                                                                  ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step=64):
    return MSDeformAttnFunction.apply(value, spatial_shapes, level_start_index, sampling_locations,
                                      attention_weights, im2col_step)
