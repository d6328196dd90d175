"""CPU-only: the C-ABI library loads and exports every symbol include/vnext_b200.h declares."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vnext_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vnext_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from vnext_b200 import _cabi
    lib = _cabi.lib()
    names = _declared_symbols()
    assert "vnext_msda_forward" in names and "vnext_msda_backward" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vnext_b200.h but not exported"
    assert set(names) == set(_cabi.SIGNATURES), "ctypes signatures out of sync with the header"
    assert b"sm_100a" in lib.vnext_version()


def test_argument_errors_without_gpu():
    from vnext_b200 import _cabi
    lib = _cabi.lib()
    rc = lib.vnext_msda_forward(7, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, None, None)
    assert rc == 1 and b"dtype" in lib.vnext_last_error()
    rc = lib.vnext_msda_forward(0, None, None, None, None, None, 1, 1, 0, 1, 1, 1, 1, None, None)
    assert rc == 1 and b"positive" in lib.vnext_last_error()
    rc = lib.vnext_msda_forward(0, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, None, None)
    assert rc == 1 and b"null" in lib.vnext_last_error()
    assert lib.vnext_msda_backward_workspace_bytes(2, 2, 10, 8, 32) == 2 * 10 * 8 * 32 * 4
    assert lib.vnext_msda_backward_workspace_bytes(0, 2, 10, 8, 32) == 0


def test_op_layer_rejects_cpu_tensors_like_reference():
    import MultiScaleDeformableAttention as MSDA
    from vnext_b200.msda import MSDeformAttnFunction
    v = torch.zeros(1, 4, 1, 2)
    sh = torch.tensor([[2, 2]])
    lsi = torch.tensor([0])
    loc = torch.zeros(1, 1, 1, 1, 1, 2)
    aw = torch.ones(1, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_forward(v, sh, lsi, loc, aw, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDeformAttnFunction.apply(v, sh, lsi, loc, aw, 64)
