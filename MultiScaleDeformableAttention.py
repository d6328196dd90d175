"""Drop-in for the reference's compiled extension module of the same name.

The reference does `import MultiScaleDeformableAttention as MSDA` and calls
`MSDA.ms_deform_attn_forward/backward` (projects/IDOL/idol/models/ops/functions/
ms_deform_attn_func.py:18-37; pybind names in src/vision.cpp:13-16).  Putting this repo root on
PYTHONPATH makes that import resolve to the B200 kernels.
"""
from vnext_b200.msda import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401
