"""ctypes binding of include/vnext_b200.h.  Fails loudly when the CUDA library is absent."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvnext_b200.so")

F32, F64, BF16, F16 = 0, 1, 2, 3

_lib = None

_c_int, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/vnext_b200.h one to one
SIGNATURES = {
    "vnext_last_error": (ctypes.c_char_p, []),
    "vnext_version": (ctypes.c_char_p, []),
    "vnext_launch_count": (ctypes.c_ulonglong, []),
    "vnext_msda_last_variant": (ctypes.c_char_p, []),
    "vnext_msda_forward": (_c_int, [_c_int] + [_c_void_p] * 5 + [_c_int] * 7 + [_c_void_p, _c_void_p]),
    "vnext_msda_forward_fused": (_c_int, [_c_int] + [_c_void_p] * 4 + [_c_int, _c_void_p, _c_int, _c_int] + [_c_int] * 7
                                 + [_c_void_p, _c_void_p]),
    "vnext_msda_forward_selfattn": (_c_int, [_c_int] + [_c_void_p] * 6 + [_c_int, _c_void_p, _c_int, _c_int] + [_c_int] * 7
                                    + [_c_void_p, _c_void_p]),
    "vnext_msda_backward": (_c_int, [_c_int] + [_c_void_p] * 6 + [_c_int] * 7 + [_c_void_p] * 5),
    "vnext_msda_backward_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "vnext_linear_tf32": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int,
                                   _c_void_p, _c_int, _c_void_p, _c_void_p, ctypes.c_float, _c_void_p, _c_int,
                                   _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "vnext_linear_tc": (_c_int, [_c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p,
                                 _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, ctypes.c_float, _c_void_p, _c_int,
                                 _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "vnext_rle_workspace_bytes": (ctypes.c_longlong, [_c_int, _c_int, _c_int]),
    "vnext_rle_encode": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, ctypes.c_longlong, _c_void_p, _c_void_p,
                                  ctypes.c_longlong, _c_void_p]),
    "vnext_linear_msda_front": (_c_int, [_c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p,
                                         _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "vnext_conv3x3_tc": (_c_int, [_c_int, _c_int] + [_c_void_p] * 4 + [_c_int] * 6 + [_c_void_p]),
    "vnext_conv3x3_strided_tc": (_c_int, [_c_int, _c_int] + [_c_void_p] * 4 + [_c_int] * 7 + [_c_void_p]),
    "vnext_time_pool_ln": (_c_int, [_c_void_p] * 6 + [_c_int] * 3 + [ctypes.c_float, _c_void_p]),
    "vnext_ref_points_px": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, ctypes.c_float, ctypes.c_float, _c_void_p]),
    "vnext_box_refine": (_c_int, [_c_void_p] * 3 + [_c_int] * 2 + [_c_void_p]),
    "vnext_stem_conv7x7": (_c_int, [_c_int, _c_int] + [_c_void_p] * 5 + [_c_int] * 6 + [_c_void_p, _c_void_p, _c_int, _c_void_p]),
    "vnext_mha_small": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "vnext_conv3x3_tf32": (_c_int, [_c_void_p] * 4 + [_c_int] * 6 + [_c_void_p]),
    "vnext_bias_act": (_c_int, [_c_void_p] * 4 + [ctypes.c_longlong, _c_int, _c_int, _c_void_p]),
    "vnext_stem_pool": (_c_int, [_c_void_p] * 3 + [_c_int] * 4 + [_c_void_p]),
    "vnext_groupnorm_workspace_floats": (_c_size_t, [_c_int] * 3),
    "vnext_groupnorm_nhwc": (_c_int, [_c_void_p] * 4 + [ctypes.c_longlong, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                                      ctypes.c_float, _c_void_p]),
    "vnext_preprocess": (_c_int, [_c_int, _c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p, _c_void_p, _c_void_p]),
    "vnext_bias_act_dt": (_c_int, [_c_int] + [_c_void_p] * 4 + [ctypes.c_longlong, _c_int, _c_int, _c_void_p]),
    "vnext_stem_pool_dt": (_c_int, [_c_int, _c_int] + [_c_void_p] * 3 + [_c_int] * 4 + [_c_void_p]),
    "vnext_groupnorm_nhwc_dt": (_c_int, [_c_int, _c_int] + [_c_void_p] * 4 + [ctypes.c_longlong, _c_void_p, _c_int, _c_int,
                                                                  _c_int, _c_int, ctypes.c_float, _c_void_p]),
    "vnext_preprocess_dt": (_c_int, [_c_int, _c_int, _c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p, _c_void_p, _c_void_p]),
    "vnext_swin_window_attention": (_c_int, [_c_void_p] * 4 + [_c_int] * 7 + [_c_void_p]),
    "vnext_swin_window_attention_tc": (_c_int, [_c_void_p] * 4 + [_c_int] * 7 + [_c_void_p]),
    "vnext_layernorm": (_c_int, [_c_void_p] * 4 + [ctypes.c_longlong, _c_int, ctypes.c_float, _c_void_p]),
    "vnext_layernorm_dt": (_c_int, [_c_int, _c_int] + [_c_void_p] * 4 + [ctypes.c_longlong, _c_int, ctypes.c_float, _c_void_p]),
    "vnext_swin_window_attention_bf16": (_c_int, [_c_void_p] * 4 + [_c_int] * 7 + [_c_void_p]),
    "vnext_swin_window_attention_tc_dt": (_c_int, [_c_int] + [_c_void_p] * 4 + [_c_int] * 7 + [_c_void_p]),
    "vnext_adamw": (_c_int, [_c_void_p] * 4 + [ctypes.c_longlong] + [ctypes.c_float] * 7 + [_c_void_p, _c_void_p]),
    "vnext_dynamic_mask_head": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p] + [_c_int] * 5 + [_c_void_p]),
    "vnext_mask_finalize": (_c_int, [_c_void_p, _c_void_p] + [_c_int] * 8 + [_c_void_p]),
    "vnext_msda_forward_host": (_c_int, [_c_int] + [_c_void_p] * 5 + [_c_int] * 7 + [_c_void_p]),
    "vnext_msda_backward_host": (_c_int, [_c_int] + [_c_void_p] * 6 + [_c_int] * 7 + [_c_void_p] * 3),
}


def lib():
    """Load (once) and return the ctypes handle; RuntimeError if it was never built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"vnext_b200: CUDA library not found at {LIB_PATH}; build it with `make` or "
                "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc, who):
    if rc != 0:
        msg = lib().vnext_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{who}: {msg} (code {rc})")


def launch_count():
    return int(lib().vnext_launch_count())


def last_variant():
    return lib().vnext_msda_last_variant().decode()


# ------------------------------------------------------------------ optional per-kernel CUDA-event timing
class KernelTimer:
    """When active, every C-ABI kernel call site brackets its launch with CUDA events on the launching
    stream.  bench.py uses it on a few eager (non-graph) steps to get the live per-kernel durations the
    roofline line is computed from; it is off in the timed region."""
    active = None

    def __init__(self):
        self.records = []          # (name, info dict, start_event, end_event)

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *exc):
        KernelTimer.active = None

    def summary(self):
        import torch
        torch.cuda.synchronize()
        agg = {}
        for name, info, a, b in self.records:
            d = agg.setdefault(name, {"calls": 0, "us": 0.0, "bytes": 0.0, "flops": 0.0})
            d["calls"] += 1
            d["us"] += a.elapsed_time(b) * 1e3
            d["bytes"] += info.get("bytes", 0.0)
            d["flops"] += info.get("flops", 0.0)
        return agg


def timed_call(name, info, stream_tensor, fn):
    """Run fn() (which enqueues exactly one library call) with optional event bracketing."""
    kt = KernelTimer.active
    if kt is None:
        return fn()
    import torch
    st = torch.cuda.current_stream(stream_tensor.device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    rc = fn()
    b.record(st)
    kt.records.append((name, info, a, b))
    return rc
