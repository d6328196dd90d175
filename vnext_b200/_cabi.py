"""ctypes binding of include/vnext_b200.h.  Fails loudly when the CUDA library is absent."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvnext_b200.so")

F32, F64, BF16 = 0, 1, 2

_lib = None

_c_int, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/vnext_b200.h one to one
SIGNATURES = {
    "vnext_last_error": (ctypes.c_char_p, []),
    "vnext_version": (ctypes.c_char_p, []),
    "vnext_launch_count": (ctypes.c_ulonglong, []),
    "vnext_msda_last_variant": (ctypes.c_char_p, []),
    "vnext_msda_forward": (_c_int, [_c_int] + [_c_void_p] * 5 + [_c_int] * 7 + [_c_void_p, _c_void_p]),
    "vnext_msda_backward": (_c_int, [_c_int] + [_c_void_p] * 6 + [_c_int] * 7 + [_c_void_p] * 5),
    "vnext_msda_backward_workspace_bytes": (_c_size_t, [_c_int] * 5),
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
