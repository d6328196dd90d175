"""Linear layers on the hand-written tcgen05 GEMM (include/vnext_b200.h: vnext_linear_tf32).

`linear(x, weight, bias, ...)` has F.linear's argument meaning (weight is [out, in]) plus the fused
epilogues the encoder/decoder layers need: ReLU, or residual-add + LayerNorm
(deformable_transformer.py:155-169: `src = norm(src + dropout(linear(...)))`).
"""
import torch

from . import _cabi

EPI_BIAS, EPI_BIAS_RELU, EPI_BIAS_RES_LN = 0, 1, 2


def _ptr(t):
    return None if t is None else t.data_ptr()


def linear(x, weight, bias=None, relu=False, residual=None, ln_weight=None, ln_bias=None, ln_eps=1e-5, out=None):
    if not x.is_cuda:
        raise RuntimeError("vnext_b200.linear: CUDA tensors only (no CPU fallback)")
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    M = x2.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    mode = EPI_BIAS_RELU if relu else EPI_BIAS
    res2 = None
    if residual is not None:
        res2 = residual.reshape(-1, N)
        if res2.stride(-1) != 1:
            res2 = res2.contiguous()
        mode = EPI_BIAS_RES_LN
    info = {"flops": 2.0 * M * N * K, "bytes": 4.0 * (M * K + N * K + M * N * (2 if residual is not None else 1))}
    with torch.cuda.device(x.device):
        rc = _cabi.timed_call(
            f"linear_tf32_n{N}_k{K}_e{mode}", info, x,
            lambda: _cabi.lib().vnext_linear_tf32(
                x2.data_ptr(), x2.stride(0), weight.data_ptr(), _ptr(bias), _ptr(res2),
                0 if res2 is None else res2.stride(0), _ptr(ln_weight), _ptr(ln_bias), float(ln_eps), out.data_ptr(),
                out.stride(0), M, N, K, mode, torch.cuda.current_stream(x.device).cuda_stream))
    _cabi.check(rc, "vnext_linear_tf32")
    return out.view(*x.shape[:-1], N)
